"""Tuning sweep over the B200.* knobs on the bench workload (dev tool; prints one line per config)."""
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:] if not a.startswith("--grid=")]
    grid_arg = [a for a in sys.argv if a.startswith("--grid=")]
    args = bench.parse_args()
    from sptag_b200 import B200Index, capi
    dev = torch.device("cuda", 0)
    folder = bench.ensure_index(args, 0, dev)
    files = bench.load_folder_arrays(folder)
    idx = B200Index.create(algo=capi.ALGO_BKT, value_type=capi.VT_FLOAT, metric=files.metric, vectors=files.vectors,
                           graph=files.graph, tree_starts=files.tree_starts, tree_nodes=files.nodes, device=0)
    idx.set_param("MaxCheck", args.maxcheck)
    d_q = bench.gen_data(args, args.nq, args.seed + 7, dev)
    d_ids = torch.empty((args.nq, args.k), dtype=torch.int32, device=dev)
    d_d = torch.empty((args.nq, args.k), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ref_ids = None
    configs = json.loads(os.environ.get("SWEEP", "[]"))
    for cfg in configs:
        try:
            for k, v in cfg.items():
                idx.set_param(k, v)
            for _ in range(2):
                idx.search_device(d_q.data_ptr(), args.nq, args.k, d_ids.data_ptr(), d_d.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                idx.search_device(d_q.data_ptr(), args.nq, args.k, d_ids.data_ptr(), d_d.data_ptr(), 0, stream)
                ms.append(idx.last_kernel_ms())
            ids = d_ids.cpu().numpy()
            if ref_ids is None:
                ref_ids = ids
            same = bool((ids == ref_ids).all())
            print("SWEEP %s ms=%.2f qps=%.0f same_ids=%s" % (json.dumps(cfg), min(ms), args.nq / min(ms) * 1000, same),
                  flush=True)
        except Exception as e:  # noqa: BLE001
            print("SWEEP %s FAILED %s" % (json.dumps(cfg), e), flush=True)


if __name__ == "__main__":
    main()
