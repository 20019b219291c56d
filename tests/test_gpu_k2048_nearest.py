"""K and CEF+1 up to 2048 (the 64-register m_Results variants needed by the reference's default RefineGraph schedule:
CEF x CEFScale + 1 = 2001 results, NeighborhoodGraph.h:459-470, and by MaxCheck > 16384), p_searchDeleted = true, and
SearchIndexIterativeFromNeareast (BKTIndex.cpp:543-595).  First run on a B200 in round 2 (20 / 20 green,
gpurun_out/r02_unverified.log); the oracle side of every case is pinned to the reference
(tests/test_oracle_pin.py::test_large_k_and_budget_bit_exact_vs_reference)."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,k,mc", [("bkt_l2_10k_128", 2048, 8192), ("bkt_l2_10k_128", 1500, 2048),
                                       ("bkt_l2_10k_128", 10, 20000), ("bkt_i8_l2_5k_100", 2048, 8192),
                                       ("bkt_cos_3k_768", 1100, 8192), ("bkt_i16_l2_5k_64", 1025, 4096)])
def test_search_k_up_to_2048(name, k, mc):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:60]
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheck", mc)
        if mc > 8192:
            idx.set_param("MaxCheckForRefineGraph", mc)
        ids, dists = idx.search(q, k)
        o = reflib.OracleIndex(files)
        o.max_check = mc
        if mc > 8192:
            o.max_check_refine = mc
        ids_o, d_o, _ = o.search(q, k)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["bkt_cos_3k_768", "bkt_l2_20k_32", "bkt_i16_l2_4k_27"])
def test_refine_with_the_reference_default_first_pass_cef(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheckForRefineGraph", 8192)
        rows, ids, dists = idx.refine_graph(2000, first=10, num=200, want_results=True)
        o = reflib.OracleIndex(files)
        o.max_check_refine = 8192
        rows_o, ids_o, d_o = o.refine_nodes(10, 200, 2000, files.degree, 1.0)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
        assert np.array_equal(rows, rows_o)
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["bkt_l2_deleted_6k_32", "bkt_cos_deleted_5k_64"])
def test_search_deleted_parameter(name):
    """Handle parameter "SearchDeleted" = p_searchDeleted of SearchIndex / GetIterator (host-side switch: the kernel
    simply gets no tombstone map); sticky until reset, refine ignores it."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:100]
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheck", 1024)
        o = reflib.OracleIndex(files)
        o.max_check = 1024
        ids0, d0 = idx.search(q, 10)
        idx.set_param("SearchDeleted", 1)
        assert idx.get_param("SearchDeleted") == "1"
        o.search_deleted = True
        ids, dists = idx.search(q, 10)
        ids_o, d_o, _ = o.search(q, 10)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
        assert (files.deleted[ids[ids >= 0]] == 1).any()
        its = idx.iterators(q[:8])                  # sampled at open
        idx.set_param("SearchDeleted", 0)
        oits = [o.iterator(qq) for qq in q[:8]]
        for b in (10, 10, 10):
            counts, iids, idists, relaxed = its.next(b)
            for i, oi in enumerate(oits):
                c, io, do, ro = oi.next(b)
                assert counts[i] == c and np.array_equal(iids[i], io) and bool(relaxed[i]) == ro
                assert np.array_equal(idists[i].view(np.int32), do.view(np.int32))
        its.close()
        for oi in oits:
            oi.close()
        ids1, d1 = idx.search(q, 10)                # back to the default: identical to the first search
        assert np.array_equal(ids0, ids1) and np.array_equal(d0.view(np.int32), d1.view(np.int32))
    finally:
        idx.close()


@pytest.mark.parametrize("name,mc,k", [("bkt_l2_20k_32", 8192, 10), ("bkt_l2_20k_32", 128, 32), ("bkt_cos_10k_128", 1024, 64),
                                       ("bkt_l2_dups", 256, 16), ("bkt_l2_3k_30", 512, 8), ("bkt_i8_cos_6k_64", 512, 10),
                                       ("bkt_l2_deleted_6k_32", 512, 10), ("bkt_i16_l2_4k_27", 300, 5),
                                       ("bkt_cos_3k_768", 2048, 100)])
def test_iterative_from_nearest(name, mc, k):
    """sptag_b200_iterator_next_from_nearest (SearchIndexIterativeFromNeareast, the SPANN head-index call): first call
    = the k nearest + re-seeding, later calls = the next k -- against the oracle, pinned to the reference on CPU."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:12]
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheck", mc)
        o = reflib.OracleIndex(files)
        o.max_check = mc
        oits = [o.iterator(qq) for qq in q]
        its = idx.iterators(q)
        for rd in range(10):
            found, ids, dists = its.next_from_nearest(k)
            for i, oi in enumerate(oits):
                ok, io, do = oi.next_from_nearest(k)
                assert bool(found[i]) == ok, (name, rd, i)
                assert np.array_equal(ids[i], io), (name, rd, i)
                assert np.array_equal(dists[i].view(np.int32), do.view(np.int32)), (name, rd, i)
        its.close()
        for oi in oits:
            oi.close()
    finally:
        idx.close()
