"""CPU test of the N>1 host logic with a world_size-2 gloo group: replica-mode query partition and
shard-mode search -> all-gather -> merge, with the oracle as the per-shard search engine and the
host merge as the merge (on the GPU box the same plumbing runs the CUDA search and merge kernels)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import reflib
    import test_oracle_pin as pin
    from sptag_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # two committed golden indexes play the two vector partitions
    names = ["bkt_l2_2k_16", "kdt_l2_2k_16"]
    g = np.load(os.path.join(HERE, "golden", names[rank] + ".npz"))
    files = reflib.IndexFiles.__new__(reflib.IndexFiles)
    pin._files_from_npz(files, g)
    sizes = [2000, 2000]
    offset = int(sharded.shard_id_offsets(sizes)[rank])
    queries = np.load(os.path.join(HERE, "golden", "bkt_l2_2k_16.npz"))["queries"]

    def local_search(q, k):
        o = reflib.OracleIndex(files)
        o.max_check = 512
        ids, d, _ = o.search(q, k, threads=1, want_stats=False)
        ids = np.where(ids >= 0, ids + offset, ids)
        return torch.from_numpy(ids), torch.from_numpy(d)

    def merge(g_ids, g_d, k):
        return sharded.merge_topk_host(g_ids.numpy(), g_d.numpy(), k)

    s = sharded.ShardedSearch(dist, local_search, merge, world)
    ids, d = s.search(queries, 10)
    lids, ld = local_search(queries, 10)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ids=ids, d=d, lids=lids.numpy(), ld=ld.numpy())
    b, e = sharded.partition_queries(queries.shape[0], world, rank)
    np.save(os.path.join(out_dir, "part%d.npy" % rank), np.array([b, e]))
    dist.destroy_process_group()


def test_shard_mode_allgather_merge_and_query_partition(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # every rank ends with the same merged lists
    assert np.array_equal(r0["ids"], r1["ids"]) and np.array_equal(r0["d"], r1["d"])
    # merged = top-10 by (dist, id) of the union of the two shards' lists
    from sptag_b200 import sharded
    exp_ids, exp_d = sharded.merge_topk_host(np.stack([r0["lids"], r1["lids"]]), np.stack([r0["ld"], r1["ld"]]), 10)
    assert np.array_equal(r0["ids"], exp_ids) and np.array_equal(r0["d"], exp_d)
    assert (r0["ids"] >= 2000).any() and (r0["ids"] < 2000).any()   # both shards contribute
    assert (np.diff(r0["d"], axis=1) >= 0).all()
    # replica-mode partition covers the query stream exactly once
    p0, p1 = np.load(tmp_path / "part0.npy"), np.load(tmp_path / "part1.npy")
    assert p0[0] == 0 and p0[1] == p1[0] and p1[1] == 64


def test_partition_and_offsets():
    from sptag_b200 import sharded
    cover = []
    for r in range(8):
        b, e = sharded.partition_queries(10003, 8, r)
        cover += list(range(b, e))
    assert cover == list(range(10003))
    assert sharded.shard_id_offsets([5, 7, 9]).tolist() == [0, 5, 12]
