"""GPU parity for the rest of SURVEY.md 8 row f3: VectorIndex::GetIterator / ResultIterator::Next on the device
(sptag_b200_iterator_open / _next / _close) against the oracle's restatement (ora_iter_*), which is pinned to the
reference's own ResultIterator in tests/test_oracle_pin.py: per call the result count, ids, distances (bit-exact) and
the RelaxedMono flag."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu


def _run(name, mc, nq, schedule, knobs=()):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:nq]
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheck", mc)
        for k, v in knobs:
            idx.set_param(k, v)
        o = reflib.OracleIndex(files)
        o.max_check = mc
        oits = [o.iterator(qq) for qq in q]
        its = idx.iterators(q)
        for step, b in enumerate(schedule):
            counts, ids, dists, relaxed = its.next(b)
            for i, oi in enumerate(oits):
                c, io, do, ro = oi.next(b)
                tag = (name, mc, step, b, i)
                assert counts[i] == c, tag
                assert np.array_equal(ids[i], io), tag
                assert np.array_equal(dists[i].view(np.int32), do.view(np.int32)), tag
                assert bool(relaxed[i]) == ro, tag
        its.close()
        for oi in oits:
            oi.close()
    finally:
        idx.close()


def test_iterator_known_answer_on_gpu():
    # Test/src/IterativeScanTest.cpp: line data, MaxCheck 5, query 0, two batches of 5 -> ids 0..9, RelaxedMono true
    from sptag_b200 import B200Index
    idx = B200Index.load(data_folder("algo_line_bkt"))
    try:
        idx.set_param("MaxCheck", 5)
        its = idx.iterators(np.zeros((1, 10), np.float32))
        got = []
        for _ in range(2):
            counts, ids, dists, relaxed = its.next(5)
            assert counts[0] == 5 and relaxed[0]
            got += ids[0].tolist()
        assert got == list(range(10))
        its.close()
    finally:
        idx.close()


@pytest.mark.parametrize("name,mc", [("bkt_l2_20k_32", 8192), ("bkt_l2_20k_32", 64), ("bkt_cos_10k_128", 1024),
                                     ("bkt_l2_3k_30", 512), ("bkt2_l2_6k_32", 256), ("bkt_cos_3k_768", 2048),
                                     ("bkt_i8_cos_6k_64", 512), ("bkt_u8_l2_6k_128", 512), ("bkt_i16_l2_4k_27", 300)])
def test_iterator_bit_exact(name, mc):
    # growing requests are capped by the previous count (the reference's ResultIterator never grows a batch)
    _run(name, mc, 24, [10, 10, 5, 7, 10, 3, 10, 1, 4])


def test_iterator_duplicate_groups_and_exhaustion():
    # 2360 vectors with exact duplicates: group members enter NGQueue with their own distance; batches of 1000 run
    # the index dry (count < batch, then 0 forever)
    _run("bkt_l2_dups", 256, 16, [32, 32, 16, 32, 8])
    _run("bkt_l2_dups", 256, 6, [1000, 1000, 1000, 1000, 1000])


def test_iterator_long_scan_and_large_batches():
    _run("bkt_l2_10k_128", 1024, 8, [50] * 40)
    _run("bkt_l2_10k_128", 8192, 4, [333, 333, 333, 100])


def test_iterator_small_queue_heads():
    # tiny shared-memory heads: the queues live almost entirely in the iterator's HBM arenas between and during calls
    _run("bkt_l2_20k_32", 1024, 12, [10, 10, 10, 25, 10], knobs=(("B200.NGCacheEntries", 8), ("B200.SPTCacheEntries", 8)))


def test_iterator_rejected_where_the_reference_rejects():
    from sptag_b200 import B200Index, capi
    idx = B200Index.load(data_folder("kdt_l2_10k_64"))
    try:
        with pytest.raises(capi.SptagB200Error):
            idx.iterators(np.zeros((1, 64), np.float32))     # "ITERATIVE NOT SUPPORT FOR KDT"
    finally:
        idx.close()


class _DeviceIteratorBatch:
    def __init__(self, files, q, max_check):
        from sptag_b200 import B200Index, capi
        self.idx = B200Index.create(algo=capi.ALGO_BKT, value_type=files.value_type, metric=files.metric,
                                    vectors=files.vectors, graph=files.graph, tree_starts=files.tree_starts,
                                    tree_nodes=files.nodes)
        self.idx.set_param("MaxCheck", max_check)
        self.its = self.idx.iterators(q)

    def next(self, b):
        return self.its.next(b)

    def close(self):
        self.its.close()
        self.idx.close()


def _golden_cases():
    import test_oracle_pin as pin
    return pin.iterator_golden_cases()


@pytest.mark.parametrize("name", _golden_cases())
def test_iterator_matches_reference_golden_outputs(name):
    """The device against the committed outputs of the reference's own ResultIterator (no oracle in between)."""
    import test_oracle_pin as pin
    pin.check_iterator_golden(name, _DeviceIteratorBatch)
