"""Host-side plumbing for the multi-GPU forms of the search path (one process per GPU).

* replica mode -- every rank holds the whole index and searches its own slice of the query stream:
  queries are the independent units, no data-path collective (`partition_queries`).
* shard mode   -- vector-partition sharding (the reference's Aggregator deployment,
  AnnService/src/Aggregator/AggregatorService.cpp:209-412, does the same over TCP and merely
  concatenates): rank r holds an independent index over its slice of the vectors (ids offset by
  `id_offset`), every rank searches the SAME batch, the per-shard top-k lists are exchanged with ONE
  all-gather and merged with the comparator of QueryResultSet.h:17-26 (`ShardedSearch`).

torch.distributed is only plumbing here (NCCL on GPUs; gloo in the CPU tests, where the local search
and the merge are injected callables).
"""
import numpy as np


def partition_queries(num_queries, world_size, rank):
    """Contiguous slice [begin, end) of the query stream owned by `rank` (replica mode)."""
    base, rem = divmod(num_queries, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_id_offsets(shard_sizes):
    """Global id offset of every shard for contiguous vector partitions."""
    return np.concatenate([[0], np.cumsum(np.asarray(shard_sizes, dtype=np.int64))[:-1]]).astype(np.int64)


def merge_topk_host(ids, dists, k):
    """Reference merge on the host (numpy): ids/dists are [num_lists, nq, k'], each list ascending by
    (dist, id) with -1/MaxDist padding.  Used by the CPU tests as the checker of the merge kernel and
    as the injected merge in the gloo test."""
    num_lists, nq, kk = ids.shape
    flat_ids = np.transpose(ids, (1, 0, 2)).reshape(nq, num_lists * kk)
    flat_d = np.transpose(dists, (1, 0, 2)).reshape(nq, num_lists * kk)
    out_ids = np.full((nq, k), -1, np.int32)
    out_d = np.full((nq, k), np.float32(np.finfo(np.float32).max) / np.float32(10), np.float32)
    for q in range(nq):
        valid = flat_ids[q] >= 0
        vi, vd = flat_ids[q][valid], flat_d[q][valid]
        order = np.lexsort((vi, vd))[:k]
        out_ids[q, :order.size] = vi[order]
        out_d[q, :order.size] = vd[order]
    return out_ids, out_d


class ShardedSearch:
    """search -> all-gather -> merge over a process group (the path bench.py's shard legs time).

    local_search(queries, k) -> (ids, dists) tensors on the group's device, ids already global
    merge(gathered_ids, gathered_dists, k) -> (ids, dists)
    on_exchange_start / on_exchange_end: optional callables invoked right before the all-gather and right after the
    merge (bench.py records CUDA events there to split search time from exchange + merge time).
    """

    def __init__(self, dist, local_search, merge, world_size, on_exchange_start=None, on_exchange_end=None):
        self.dist = dist
        self.local_search = local_search
        self.merge = merge
        self.world_size = world_size
        self.on_exchange_start = on_exchange_start
        self.on_exchange_end = on_exchange_end
        self._g_ids = self._g_d = None

    def search(self, queries, k):
        import torch
        ids, dists = self.local_search(queries, k)
        nq = ids.shape[0]
        if self.on_exchange_start:
            self.on_exchange_start()
        # concatenation along dim 0 == the [num_lists][nq][k] layout the merge kernel expects; the gather buffers are
        # kept between calls (800 KB per rank at nq = 10k, k = 10)
        shape_i = (self.world_size * nq,) + tuple(ids.shape[1:])
        if self._g_ids is None or tuple(self._g_ids.shape) != shape_i or self._g_ids.device != ids.device:
            self._g_ids = torch.empty(shape_i, dtype=ids.dtype, device=ids.device)
            self._g_d = torch.empty(shape_i, dtype=dists.dtype, device=dists.device)
        self.dist.all_gather_into_tensor(self._g_ids, ids.contiguous())
        self.dist.all_gather_into_tensor(self._g_d, dists.contiguous())
        shape = (self.world_size, nq) + tuple(ids.shape[1:])
        out = self.merge(self._g_ids.view(shape), self._g_d.view(shape), k)
        if self.on_exchange_end:
            self.on_exchange_end()
        return out

    def gathered(self):
        """The per-shard lists of the last search: ([world, nq, k] ids, dists)."""
        nq = self._g_ids.shape[0] // self.world_size
        shape = (self.world_size, nq) + tuple(self._g_ids.shape[1:])
        return self._g_ids.view(shape), self._g_d.view(shape)
