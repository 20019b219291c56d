"""Vector-partition shards of one process (sptag_b200_group_*): every shard is an independent index with
its id offset (on its own GPU when the box has several), a group search runs all of them and merges with the
(Dist, VID) comparator of QueryResultSet.h:17-26 by reading the shards' lists through peer memory.  Expected result:
the oracle searched shard by shard, lists merged on the host -- the reference-side analogue is the Aggregator
(AggregatorService.cpp:215-412)."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("names,k,mc", [
    (["bkt2_l2_6k_32", "kdt2_l2_6k_32", "bkt_l2_20k_32", "bkt_l2_deleted_6k_32"], 10, 1024),
    (["bkt_l2_10k_128", "bkt_l2_10k_128"], 16, 512),     # the same partition twice: ties on distance, ordered by id
    (["bkt_cos_3k_768"], 10, 2048),
])
def test_group_search_equals_per_shard_oracle_merge(names, k, mc):
    import torch
    from sptag_b200 import B200Index, capi, sharded
    ndev = torch.cuda.device_count()
    folders = [data_folder(n) for n in names]
    q = np.load(os.path.join(folders[0], "queries.npy"))[:80]
    shards, lists_i, lists_d, offset = [], [], [], 0
    try:
        for i, folder in enumerate(folders):
            files = reflib.IndexFiles(folder)
            idx = B200Index.load(folder, device=i % ndev, id_offset=offset)
            idx.set_param("MaxCheck", mc)
            shards.append(idx)
            o = reflib.OracleIndex(files)
            o.max_check = mc
            ids_o, d_o, _ = o.search(q, k)
            lists_i.append(np.where(ids_o >= 0, ids_o + offset, ids_o))
            lists_d.append(d_o)
            offset += files.n
        group = capi.B200ShardGroup(shards)
        try:
            for _ in range(2):   # second call: buffers reused
                ids, dists = group.search(q, k)
                e_ids, e_d = sharded.merge_topk_host(np.stack(lists_i), np.stack(lists_d), k)
                assert np.array_equal(ids, e_ids)
                assert np.array_equal(dists.view(np.int32), e_d.view(np.int32))
        finally:
            group.close()
    finally:
        for s in shards:
            s.close()
