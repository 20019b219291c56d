#!/usr/bin/env python
"""Device graph refinement at scale: build an index with the torch set-up builder, load it into
the product library, run sptag_b200_refine_graph passes (RefineSearchIndex + RebuildNeighbors on the device, installed
in place) and report recall@k / kernel QPS per MaxCheck before and after each pass, plus the pass throughput.

    python tools/refine_bench.py --num-vectors 10000000 --dim 128 --metric L2 --algo kdt --cef 128 --mcr 2048
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-vectors", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--metric", default="L2", choices=["L2", "Cosine"])
    ap.add_argument("--algo", default="bkt", choices=["bkt", "kdt"])
    ap.add_argument("--data", default="lowrank", choices=["lowrank", "iid"])
    ap.add_argument("--rank-dim", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--nq", type=int, default=2000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--maxchecks", default="512,2048,8192")
    ap.add_argument("--cef", type=int, default=128)
    ap.add_argument("--mcr", type=int, default=2048, help="MaxCheckForRefineGraph")
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--tpt-above", type=int, default=2500000)
    ap.add_argument("--cand", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="also run the reference's RefineSearchIndex + RebuildNeighbors (oracle/_ref) on the first N "
                         "nodes of the same index: parity check at scale + CPU nodes/s")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    a.raw_type = "float"
    a.n = a.num_vectors
    if a.rank_dim <= 0:
        a.rank_dim = 32 if a.dim >= 512 else 16

    import numpy as np
    import torch
    import bench
    from tools import gpu_index_builder as B
    from sptag_b200 import B200Index, capi

    dev = torch.device("cuda", 0)
    log = bench.log
    x = bench.gen_data(a, a.n, a.seed + 1000, dev)
    q = bench.gen_data(a, a.nq, a.seed + 7, dev)
    truth = B.exact_topk(x, q, a.k, a.metric)
    t0 = time.time()
    torch.backends.cuda.matmul.allow_tf32 = True
    nodes, starts, graph = B.build_index(x, a.metric, seed=a.seed, log=log, algo=a.algo.upper(), tpt_above=a.tpt_above,
                                         cand=a.cand)
    torch.backends.cuda.matmul.allow_tf32 = False
    build_s = time.time() - t0
    xh = x.cpu().numpy()
    qh = q.cpu().numpy()
    del x, q
    torch.cuda.empty_cache()
    idx = B200Index.create(algo=capi.ALGO_KDT if a.algo == "kdt" else capi.ALGO_BKT, value_type=capi.VT_FLOAT,
                           metric=capi.METRIC_COSINE if a.metric == "Cosine" else capi.METRIC_L2, vectors=xh,
                           graph=graph, tree_starts=starts, tree_nodes=nodes)
    idx.set_param("MaxCheckForRefineGraph", a.mcr)

    def measure(tag):
        rows = []
        for mc in [int(v) for v in a.maxchecks.split(",")]:
            idx.set_param("MaxCheck", mc)
            idx.search(qh, a.k)
            ids, _ = idx.search(qh, a.k)
            ms = idx.last_kernel_ms()
            rec = bench.recall_at_k(ids, truth, a.k)
            rows.append({"max_check": mc, "recall": round(float(rec), 5), "kernel_qps": round(a.nq / ms * 1000.0, 1)})
            log("%s MaxCheck %5d: recall@%d %.4f, kernel %.0f q/s" % (tag, mc, a.k, rec, a.nq / ms * 1000.0))
        return rows

    report = {"n": a.n, "dim": a.dim, "metric": a.metric, "algo": a.algo, "data": a.data, "cef": a.cef,
              "max_check_refine": a.mcr, "graph": "partition-tree candidates" if a.n > a.tpt_above else "brute-force kNN",
              "build_seconds": round(build_s, 1), "stages": [{"stage": "built", "curve": measure("built")}]}
    if a.cpu_sample > 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import reflib
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            B.save_index_folder(tmp, xh, graph, nodes, starts, a.metric, algo=a.algo.upper())
            r = reflib.RefIndex.load(tmp)
            r.set_param("MaxCheckForRefineGraph", a.mcr)
            ns = min(a.cpu_sample, a.n)
            threads = os.cpu_count() or 1
            r.refine_nodes(0, min(ns, 64), a.cef, graph.shape[1], 1.0, threads=threads)  # warm-up
            t = time.time()
            rows_r, ids_r, d_r = r.refine_nodes(0, ns, a.cef, graph.shape[1], 1.0, threads=threads)
            cpu_s = time.time() - t
            rows_g, ids_g, d_g = idx.refine_graph(a.cef, first=0, num=ns, want_results=True)
            same = bool(np.array_equal(rows_r, rows_g) and np.array_equal(ids_r, ids_g)
                        and np.array_equal(d_r.view(np.int32), d_g.view(np.int32)))
            report["parity_vs_reference"] = {"nodes": ns, "rows_ids_dists_bit_exact": same}
            report["cpu_reference"] = {"nodes_per_second": round(ns / cpu_s, 1), "threads": threads, "sample_nodes": ns,
                                       "kind": "reference (oracle/_ref RefineSearchIndex + RebuildNeighbors)"}
            log("reference refine on %d nodes: %.0f nodes/s on %d threads; device rows/ids/dists bit-exact: %s"
                % (ns, ns / cpu_s, threads, same))
            del r
    for p in range(a.passes):
        g0 = idx.get_graph() if p == 0 else None
        t = time.time()
        idx.refine_graph(a.cef, install=True, want_rows=False)
        dt = time.time() - t
        g1 = idx.get_graph()
        deg = float((g1 >= 0).sum(1).mean())
        s_ms = int(idx.get_param("B200.LastRefineSearchUs")) / 1000.0
        r_ms = int(idx.get_param("B200.LastRefineRebuildUs")) / 1000.0
        log("refine pass %d: %.1fs (%.0f nodes/s; search kernel %.0f ms, RebuildNeighbors kernel %.0f ms), mean out-degree %.2f"
            % (p + 1, dt, a.n / dt, s_ms, r_ms, deg))
        st = {"stage": "refine pass %d" % (p + 1), "seconds": round(dt, 2), "nodes_per_second": round(a.n / dt, 1),
              "mean_out_degree": round(deg, 2), "search_kernel_ms": round(s_ms, 1),
              "rebuild_kernel_ms": round(r_ms, 1), "curve": measure("pass %d" % (p + 1))}
        if g0 is not None:
            st["rows_changed_frac"] = round(float((g0 != g1).any(1).mean()), 4)
        report["stages"].append(st)
    idx.close()
    line = json.dumps(report)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
