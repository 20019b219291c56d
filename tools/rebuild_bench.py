#!/usr/bin/env python
"""EnableRebuild's in-degree repair (NeighborhoodGraph::RebuildGraph, NeighborhoodGraph.h:404-456) at scale: build an
index with the torch set-up builder (rows of 32 = 2 x 16 candidates), run sptag_b200_rebuild_graph on the device and the
unmodified reference's RebuildGraph (oracle/_ref, one thread: its only deterministic order) + the oracle's restatement on
the same rows; report times and parity.

    python tools/rebuild_bench.py --num-vectors 1000000 --dim 128
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-vectors", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--metric", default="L2", choices=["L2", "Cosine"])
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import reflib
    from sptag_b200 import B200Index
    args = argparse.Namespace(algo="bkt", metric=a.metric, n=a.num_vectors, dim=a.dim, data="lowrank",
                              rank_dim=32 if a.dim >= 512 else 16, seed=a.seed, tpt_above=2500000, quantizer="none",
                              raw_type="float", builder="gpu", cache=os.environ.get("SPTAG_B200_CACHE", "/tmp/sptag_b200_cache"))
    folder = bench.ensure_index(args, 0, torch.device("cuda", 0))
    idx = B200Index.load(folder)
    wide = idx.get_graph()
    n2 = wide.shape[1] // 2
    t0 = time.time()
    rows = idx.rebuild_graph()
    dev_s = time.time() - t0
    g = wide.copy()
    g[g < -1] = -1
    t0 = time.time()
    exp = reflib.oracle_rebuild_graph(g, n2)[:, :n2]
    ora_s = time.time() - t0
    out = {"tool": "rebuild_bench", "nodes": int(wide.shape[0]), "row_width": int(wide.shape[1]), "neighborhood": int(n2),
           "device_seconds": round(dev_s, 3), "device_nodes_per_s": round(wide.shape[0] / dev_s),
           "oracle_seconds_1_thread": round(ora_s, 3), "rows_identical_to_oracle": bool(np.array_equal(rows, exp)),
           "rows_changed": int((rows != wide[:, :n2]).any(axis=1).sum())}
    if not a.skip_reference:
        r = reflib.RefIndex.load(folder)
        t0 = time.time()
        ref_rows = r.rebuild_graph(g, n2)[:, :n2]
        out["reference_seconds_1_thread"] = round(time.time() - t0, 3)   # incl. its GraphAccuracyEstimation log line
        out["rows_identical_to_reference"] = bool(np.array_equal(rows, ref_rows))
    idx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
