#!/usr/bin/env python
"""bench.py -- QPS of the batched search hot path (BASELINE.json metric) on N B200s of one node.

A "step" is one pass of the hot path over one batch of synthetic queries:
    VectorIndex::SearchIndex(batch of 10 000 queries, k = 10)  over a device-resident BKT + RNG index.

Workload at N=1 = BASELINE.json configs[1]: SPTAG-BKT, 1M x 768 float32, cosine, batch 10k queries
(`--n/--dim/--metric/--nq` override it for experiments; the JSON always names what ran).

JSON line keys (bench contract): metric/value/unit = whole-job QPS with inputs resident in HBM;
`e2e` = the same metric through the C-ABI entry point with HOST buffers (H2D of the queries and D2H of
the results inside the timed region); `roofline` = achieved algorithmic GB/s of the search kernel
(sum over queries of D_q*row_bytes + E_q*degree*4 + Tn_q*12) / CUDA-event time vs
MEASURED_PEAKS.json; `cpu_baseline` = the reference's own CPU search (oracle/_ref, all host threads)
on a bounded sample of the same batch on the same index files.

--impl reference times the UNMODIFIED reference CPU implementation (oracle/_ref/libsptag_ref.so,
VectorIndex::LoadIndex + SearchIndex(batch)) on the same index folder.  The reference/oracle is only
ever executed in that leg and in the cpu_baseline leg -- never on the product path.

Multi-GPU (`--gpus N` under torchrun, one rank per GPU):
  --mode replica (default): every rank holds the index and searches its OWN 10k-query batch; no
        data-path collective (queries are the independent units); weak scaling.
  --mode shard: vector-partition sharding: rank r holds shard r (ids offset), every rank
        searches the SAME batch, results are exchanged with one NCCL all-gather and merged on the GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BUILDER_VERSION = 3


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="replica", choices=["replica", "shard"])
    # (--num-vectors: under `python -m torch.distributed.run` a bare --n is swallowed by torchrun's own
    #  abbreviation matching (--nnodes / --nproc-per-node), so multi-GPU launches must use the long name)
    ap.add_argument("--n", "--num-vectors", dest="n", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--metric", default="Cosine", choices=["Cosine", "L2"])
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--maxcheck", type=int, default=8192)
    ap.add_argument("--data", default="lowrank", choices=["lowrank", "iid"])
    ap.add_argument("--rank-dim", type=int, default=32, help="latent rank of the low-rank synthetic set")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cache", default=os.environ.get("SPTAG_B200_CACHE", "/tmp/sptag_b200_cache"))
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="debug: do not sample nvidia-smi during the timed region")
    ap.add_argument("--param", action="append", default=[], help="Name=Value passed to sptag_b200_set_param")
    ap.add_argument("--tpt-above", type=int, default=2500000,
                    help="builder: above this many vectors use partition-tree kNN candidates instead of brute force")
    ap.add_argument("--algo", default="bkt", choices=["bkt", "kdt"], help="space-partition tree of the index")
    ap.add_argument("--quantizer", default="none", choices=["none", "pq", "opq"],
                    help="index holds uint8 PQ codes (BASELINE config 4 shape: --quantizer opq --raw-type int8 --dim 100 --pq-m 50)")
    ap.add_argument("--pq-m", type=int, default=50, help="number of PQ sub-vectors")
    ap.add_argument("--raw-type", default="float", choices=["float", "int8"], help="element type of raw vectors/queries")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic data + index folder (built on the GPU once per box, cached under --cache)
# ---------------------------------------------------------------------------------------------
def gen_data(args, n, seed, device):
    """Synthetic vectors [n, dim] (float32 values; int8-valued when --raw-type int8).  Generated in 10M-row pieces so
    that 100M-point sets never need more than the result plus one piece of temporaries."""
    import torch
    piece = 10000000
    if n > piece:
        out = torch.empty((n, args.dim), dtype=torch.float32, device=device)
        for i, s in enumerate(range(0, n, piece)):
            e = min(n, s + piece)
            out[s:e] = gen_data(args, e - s, seed * 1000003 + i + 1, device)
        return out
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if args.data == "iid":
        x = torch.randn((n, args.dim), generator=g, device=device, dtype=torch.float32)
    else:
        # BASELINE.md "low-rank synthetic": x = z.A + 0.1*eps, A entries N(0,1)/sqrt(r)
        ga = torch.Generator(device=device)
        ga.manual_seed(args.seed)  # the mixing matrix is shared by base vectors, shards and queries
        r = args.rank_dim
        A = torch.randn((r, args.dim), generator=ga, device=device, dtype=torch.float32) / (r ** 0.5)
        z = torch.randn((n, r), generator=g, device=device, dtype=torch.float32)
        x = z @ A
        x += 0.1 * torch.randn((n, args.dim), generator=g, device=device, dtype=torch.float32)
    if args.raw_type == "int8":
        # SPACEV-style int8 raw vectors: clamp(round(32 x), -127, 127), kept as float values here
        x = torch.clamp(torch.round(32.0 * x), -127, 127)
    if args.metric == "Cosine":
        # the reference normalises base vectors at build time (BKTIndex.cpp:749-756) and expects
        # unit-norm queries from the caller
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-30)
    return x.contiguous()


def index_folder(args, shard):
    key = "%s_%s_%dx%d_%s_r%d_s%d_shard%d_v%d" % (args.algo, args.metric, args.n, args.dim, args.data, args.rank_dim,
                                                   args.seed, shard, BUILDER_VERSION)
    if args.tpt_above != 2500000:
        key += "_tpt%d" % args.tpt_above
    if args.quantizer != "none":
        key += "_%s%d_%s" % (args.quantizer, args.pq_m, args.raw_type)
    elif args.raw_type != "float":
        key += "_" + args.raw_type
    return os.path.join(args.cache, key)


def ensure_index(args, shard, device):
    """Build (GPU, torch) and save the reference-format folder unless it is cached. Returns folder."""
    import numpy as np
    import torch
    from tools import gpu_index_builder as B
    folder = index_folder(args, shard)
    done = os.path.join(folder, "DONE")
    if os.path.exists(done):
        return folder
    t0 = time.time()
    torch.backends.cuda.matmul.allow_tf32 = True
    x = gen_data(args, args.n, args.seed + 1000 * (shard + 1), device)
    nodes, starts, graph = B.build_index(x, args.metric, seed=args.seed + shard, log=log, algo=args.algo.upper(),
                                         tpt_above=args.tpt_above)
    if args.quantizer != "none":
        cb, rot = B.train_quantizer_gpu(x, args.pq_m, opq=(args.quantizer == "opq"), seed=args.seed)
        codes = B.encode_gpu(x, cb, rot)
        blob = B.quantizer_blob(cb, rot, 0 if args.raw_type == "int8" else 3)
        B.save_index_folder(folder, codes.cpu().numpy(), graph, nodes, starts, args.metric, quantizer=blob)
    elif args.raw_type == "int8":   # unquantized int8 rows (DistanceUtils int8 variants), e.g. SPACEV / PerfTest.cpp shape
        B.save_index_folder(folder, x.cpu().numpy().astype(np.int8), graph, nodes, starts, args.metric,
                            algo=args.algo.upper(), value_type="Int8")
    else:
        B.save_index_folder(folder, x.cpu().numpy(), graph, nodes, starts, args.metric, algo=args.algo.upper())
    torch.backends.cuda.matmul.allow_tf32 = False
    with open(done, "w") as f:
        f.write("ok\n")
    del x
    torch.cuda.empty_cache()
    log("index shard %d built and saved in %.1fs -> %s" % (shard, time.time() - t0, folder))
    return folder


class _InMemoryIndex:
    """What bench needs from reflib.IndexFiles, for an index that is built and handed to the device without touching
    the disk (shard mode on many GPUs: eight 8-GB folders would not fit the box's scratch disk)."""

    def __init__(self, vectors, graph, nodes, tree_starts, metric_name):
        self.vectors, self.graph, self.nodes, self.tree_starts = vectors, graph, nodes, tree_starts
        self.value_type = 3
        self.metric = {"L2": 0, "Cosine": 1}[metric_name]
        self.n, self.dim = vectors.shape
        self.degree = graph.shape[1]
        self.quantizer = None


def build_in_memory(args, shard, device):
    import numpy as np
    import torch
    from tools import gpu_index_builder as B
    t0 = time.time()
    torch.backends.cuda.matmul.allow_tf32 = True
    x = gen_data(args, args.n, args.seed + 1000 * (shard + 1), device)
    nodes, starts, graph = B.build_index(x, args.metric, seed=args.seed + shard, log=log, algo=args.algo.upper(),
                                         tpt_above=args.tpt_above)
    torch.backends.cuda.matmul.allow_tf32 = False
    files = _InMemoryIndex(x.cpu().numpy(), graph, nodes, starts, args.metric)
    del x
    torch.cuda.empty_cache()
    log("index shard %d built in memory in %.1fs" % (shard, time.time() - t0))
    return files


def load_folder_arrays(folder):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    return reflib.IndexFiles(folder)


# ---------------------------------------------------------------------------------------------
# clocks sampler (profiling recipe's nvidia-smi line)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference / cpu baseline leg (the ONLY place the oracle is executed by bench.py)
# ---------------------------------------------------------------------------------------------
def cpu_search_leg(folder, queries_np, k, maxcheck, threads, sample, repeats=1, each=False):
    """Times the reference's CPU SearchIndex(batch) on `sample` queries. Returns dict + ids for a parity check."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    q = np.ascontiguousarray(queries_np[:sample])
    if reflib.have_ref():
        kind = "reference"
        idx = reflib.RefIndex.load(folder)
        idx.set_param("MaxCheck", maxcheck)
        # quantized indexes: the per-query overload on raw queries, as IndexSearcher does
        run = idx.search_each if each else idx.search
        run(q[:min(256, sample)], k, threads=threads)  # creates the per-thread work spaces
        best = None
        for _ in range(repeats):
            ids, dists, sec = run(q, k, threads=threads)
            best = sec if best is None else min(best, sec)
        isa = reflib.ref().ref_isa()
    else:
        kind = "port"
        files = reflib.IndexFiles(folder)
        o = reflib.OracleIndex(files)
        o.max_check = maxcheck
        best = None
        for _ in range(repeats):
            t = time.time()
            ids, dists, _ = o.search(q, k, threads=threads, want_stats=False)
            sec = time.time() - t
            best = sec if best is None else min(best, sec)
        isa = 512
    return {"value": sample / best, "unit": "queries/s", "cores": threads, "kind": kind,
            "sample": "%d of the %d-query batch, MaxCheck %d, %d OpenMP threads, ISA %d, best of %d"
                      % (sample, queries_np.shape[0], maxcheck, threads, isa, repeats),
            "seconds": best}, ids, dists


def best_cpu_threads(folder, queries_np, k, maxcheck, each=False):
    """The reference gets every host thread it can use; on SMT boxes one thread per physical core is
    sometimes faster for this DRAM-bound loop, so probe both and keep the faster."""
    n = os.cpu_count() or 1
    cands = sorted({n, max(1, n // 2)}, reverse=True)
    best_t, best_v = n, -1.0
    for t in cands:
        r, _, _ = cpu_search_leg(folder, queries_np, k, maxcheck, t, min(queries_np.shape[0], 1024), each=each)
        if r["value"] > best_v:
            best_t, best_v = t, r["value"]
    return best_t


def recall_at_k(ids, truth, k):
    import numpy as np
    hit = 0
    for i in range(ids.shape[0]):
        hit += len(set(ids[i, :k].tolist()) & set(truth[i, :k].tolist()))
    return hit / float(ids.shape[0] * k)


# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world

    quantized = args.quantizer != "none"
    if quantized or args.raw_type == "int8":
        # the reference's quantizers have no cosine distance (PQQuantizer.h:130-136); int8 cosine needs base-127
        # normalised rows, which this synthetic generator does not produce
        args.metric = "L2"
    workload = "SPTAG-%s, %dx%d %s %s, batch %d queries, k=%d, MaxCheck=%d, %s synthetic" % (
        args.algo.upper(), args.n, args.dim, "float32" if args.raw_type == "float" else "int8", args.metric.lower(), args.nq, args.k,
        args.maxcheck, args.data)
    if quantized:
        workload += ", %s uint8 codes M=%d Ks=256 (SDC)" % (args.quantizer.upper(), args.pq_m)
    config = {"workload": workload, "index": "%s+RNG(degree 32)" % args.algo.upper(), "n": args.n, "dim": args.dim,
              "metric": args.metric, "batch": args.nq, "k": args.k, "max_check": args.maxcheck,
              "parallelism": ("%s x%d" % (args.mode, args.gpus)) if args.gpus > 1 else "single GPU",
              "l2_policy": "index (%.2f GB of vector rows) and per-step traffic are larger than L2; no explicit flush"
                           % (args.n * (args.pq_m if quantized else args.dim * 4) / 1e9)}

    # ------------------------------ reference arm ------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        import numpy as np
        import torch
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else None
        if dev is None and not os.path.exists(os.path.join(index_folder(args, 0), "DONE")):
            print(json.dumps({"impl": "reference", "unavailable": "index folder not cached and no GPU to build it"}))
            return 0
        folder = ensure_index(args, 0, dev) if dev is not None else index_folder(args, 0)
        q = gen_data(args, args.nq, args.seed + 7, dev if dev is not None else "cpu").cpu().numpy()
        if args.raw_type == "int8":
            q = q.astype(np.int8)
        threads = best_cpu_threads(folder, q, args.k, args.maxcheck, each=quantized)
        # bounded sample per step: probe the speed, then size a step to ~3 s of CPU work
        probe, _, _ = cpu_search_leg(folder, q, args.k, args.maxcheck, threads, min(args.nq, 512), each=quantized)
        sample = int(max(256, min(args.nq, probe["value"] * 3.0)))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import reflib
        idx = reflib.RefIndex.load(folder) if reflib.have_ref() else None
        secs = []
        if idx is not None:
            idx.set_param("MaxCheck", args.maxcheck)
            kind = "reference"
            for s in range(args.warmup + args.steps):
                _, _, sec = (idx.search_each if quantized else idx.search)(q[:sample], args.k, threads=threads)
                if s >= args.warmup:
                    secs.append(sec)
        else:
            kind = "port"
            o = reflib.OracleIndex(reflib.IndexFiles(folder))
            o.max_check = args.maxcheck
            for s in range(args.warmup + args.steps):
                t = time.time()
                o.search(q[:sample], args.k, threads=threads, want_stats=False)
                if s >= args.warmup:
                    secs.append(time.time() - t)
        total = sum(secs)
        qps = sample * len(secs) / total
        line = {"impl": "reference", "metric": "queries_per_second", "value": qps, "unit": "queries/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * total / len(secs), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind,
                                 "sample": "%d of the %d-query batch per step, MaxCheck %d, %d OpenMP threads"
                                           % (sample, args.nq, args.maxcheck, threads)},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------ B200 arm ------------------------------
    import numpy as np
    import torch
    import __graft_entry__
    if not os.path.exists(__graft_entry__.LIB):
        __graft_entry__.build_cuda()
    from sptag_b200 import B200Index, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))

    # ---- set-up (untimed): index folder(s), device-resident index, queries, ground truth ----
    shard = rank if args.mode == "shard" else 0
    if world > 1 and args.mode == "replica":
        if rank == 0:
            ensure_index(args, 0, dev)
        dist.barrier()
    if world > 1 and args.mode == "shard" and not quantized and args.raw_type == "float":
        folder = None
        files = build_in_memory(args, shard, dev)   # no reference leg in this mode: nothing needs the folder
    else:
        folder = ensure_index(args, shard, dev)
        files = load_folder_arrays(folder)
    id_offset = shard * args.n if args.mode == "shard" else 0
    t0 = time.time()
    idx = B200Index.create(algo=capi.ALGO_KDT if args.algo == "kdt" else capi.ALGO_BKT, value_type=files.value_type, metric=files.metric, vectors=files.vectors,
                           graph=files.graph, tree_starts=files.tree_starts, tree_nodes=files.nodes,
                           device=local_rank, id_offset=id_offset)
    if quantized:
        idx.set_quantizer(files.quantizer.blob())
    idx.set_param("MaxCheck", args.maxcheck)
    for kv in args.param:
        nm, v = kv.split("=", 1)
        idx.set_param(nm, v)
    log("index uploaded to HBM in %.1fs" % (time.time() - t0))

    qseed = args.seed + 7 + (rank if args.mode == "replica" else 0)
    d_q_f32 = gen_data(args, args.nq, qseed, dev)
    qdtype = torch.int8 if args.raw_type == "int8" else torch.float32
    d_q = d_q_f32.to(qdtype).contiguous()
    h_q = torch.empty((args.nq, args.dim), dtype=qdtype, pin_memory=True)
    h_q.copy_(d_q)
    d_ids = torch.empty((args.nq, args.k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((args.nq, args.k), dtype=torch.float32, device=dev)
    d_stats = torch.zeros((args.nq, capi.STATS_PER_QUERY), dtype=torch.int32, device=dev)
    h_ids = torch.empty((args.nq, args.k), dtype=torch.int32, pin_memory=True)
    h_dists = torch.empty((args.nq, args.k), dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream

    gathered_ids = gathered_d = m_ids = m_d = None
    if args.mode == "shard" and world > 1:
        gathered_ids = torch.empty((world * args.nq, args.k), dtype=torch.int32, device=dev)
        gathered_d = torch.empty((world * args.nq, args.k), dtype=torch.float32, device=dev)
        m_ids = torch.empty_like(d_ids)
        m_d = torch.empty_like(d_dists)

    def step_device(with_stats=False, local_only=False):
        idx.search_device(d_q.data_ptr(), args.nq, args.k, d_ids.data_ptr(), d_dists.data_ptr(),
                          d_stats.data_ptr() if with_stats else 0, stream)
        if gathered_ids is not None and not local_only:
            dist.all_gather_into_tensor(gathered_ids, d_ids)
            dist.all_gather_into_tensor(gathered_d, d_dists)
            capi.merge_topk(local_rank, gathered_ids.data_ptr(), gathered_d.data_ptr(), world, args.nq, args.k,
                            m_ids.data_ptr(), m_d.data_ptr(), stream)

    def step_e2e():
        # the call a user makes: host query buffer in, host results out (H2D + kernel + D2H, blocking)
        idx.search(h_q.numpy(), args.k, out_ids=h_ids.numpy(), out_dists=h_dists.numpy())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # one untimed pass with counters -> algorithmic bytes of one launch
    step_device(with_stats=True)
    torch.cuda.synchronize()
    st = d_stats.cpu().numpy().astype(np.int64)
    row_bytes = files.vectors.shape[1] * files.vectors.itemsize  # dim*4, or M code bytes when quantized
    alg_bytes = int((st[:, capi.ST_NDIST] * row_bytes + st[:, capi.ST_NEXPAND] * files.degree * 4
                     + st[:, capi.ST_NTREE] * (16 if args.algo == "kdt" else 12)).sum())
    res_ids = (m_ids if m_ids is not None else d_ids).cpu().numpy()
    shard_merge_check = None
    if gathered_ids is not None:
        from sptag_b200 import sharded
        nchk = min(args.nq, 512)
        gi = gathered_ids.view(world, args.nq, args.k)[:, :nchk].cpu().numpy()
        gd = gathered_d.view(world, args.nq, args.k)[:, :nchk].cpu().numpy()
        e_ids, e_d = sharded.merge_topk_host(gi, gd, args.k)
        shard_merge_check = {"queries": nchk, "identical": bool((e_ids == res_ids[:nchk]).all()
                                                                and (e_d == m_d[:nchk].cpu().numpy()).all()),
                             "ids_from_other_shards": int((res_ids[:nchk] // args.n != rank).sum())}

    # recall@10 against exact search (untimed)
    recall = None
    if args.mode == "replica" or world == 1:
        from tools import gpu_index_builder as B
        if quantized:  # ground truth on the raw vectors (regenerated: the folder only holds codes)
            x_dev = gen_data(args, args.n, args.seed + 1000 * (shard + 1), dev)
        else:
            x_dev = torch.from_numpy(files.vectors).to(dev).float()   # int8 rows: exact truth on their float values
        truth = B.exact_topk(x_dev, d_q_f32, args.k, args.metric)
        del x_dev
        torch.cuda.empty_cache()
        recall = recall_at_k(res_ids, truth, args.k)
        log("recall@%d = %.4f, mean D_q = %.0f, E_q = %.0f, Tn_q = %.0f" % (
            args.k, recall, st[:, capi.ST_NDIST].mean(), st[:, capi.ST_NEXPAND].mean(), st[:, capi.ST_NTREE].mean()))

    # ---- timed region 1: inputs resident in HBM (value, roofline) ----
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()
        time.sleep(0.3)
    for _ in range(args.warmup):
        step_device()
    sync_all()
    launches0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tm0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    sync_all()
    tm1 = time.time()
    launches = capi.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    queries_per_step = args.nq * (world if args.mode == "replica" else 1)
    value = queries_per_step / (ms_step / 1000.0)

    # ---- timed region 2: end to end through the C-ABI with host buffers ----
    for _ in range(args.warmup):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = queries_per_step * args.steps / e2e_s
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(tm0, tm1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel of a step ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # kernel duration: the step on this stream is memset(4 B) + the search kernel; time the kernel alone
    # with the library's own CUDA events (recorded on the launching stream around the launch)
    kms = []
    for _ in range(5):
        step_device(local_only=True)  # rank 0 only from here on: no collectives
        kms.append(idx.last_kernel_ms())
    kernel_ms = float(np.mean(kms))
    achieved = alg_bytes / (kernel_ms / 1000.0) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": ("search_kernel<PQ,L2,BKT>" if quantized else "search_kernel<int8,L2,%s>" % args.algo.upper()
                           if args.raw_type == "int8" else
                           "search_kernel<%d,%s,%s>" % (args.dim if args.dim in (128, 768) else 0, args.metric, args.algo.upper())), "kernel_ms": kernel_ms,
                "kernel_ms_samples": [round(v, 3) for v in kms],
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "share_of_step": kernel_ms / ms_step}

    # ---- cpu baseline on a bounded sample (rank 0, N=1 only) ----
    cpu_baseline = None
    parity = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        qn = h_q.numpy()
        threads = best_cpu_threads(folder, qn, args.k, args.maxcheck, each=quantized)
        probe, _, _ = cpu_search_leg(folder, qn, args.k, args.maxcheck, threads, min(args.nq, 512), each=quantized)
        sample = args.cpu_sample or int(max(512, min(args.nq, probe["value"] * 5.0)))
        cpu_baseline, cpu_ids, cpu_d = cpu_search_leg(folder, qn, args.k, args.maxcheck, threads, sample, repeats=3,
                                                      each=quantized)
        same = int((cpu_ids == res_ids[:sample]).all(axis=1).sum())
        parity = {"queries_compared": sample, "identical_id_lists": same}
        cpu_baseline.pop("seconds", None)

    line = {"metric": "queries_per_second", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if not quantized else "u8 codes, f32 SDC sums", "data": "synthetic", "config": config,
            "recall_at_10": recall, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": args.nq * args.dim * h_q.element_size(),
                    "d2h_bytes_per_step": args.nq * args.k * 8},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_vs_reference": parity, "shard_merge_check": shard_merge_check}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
