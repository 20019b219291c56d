"""B200.SimdWidth 8 / 4: the reference picks its DistanceUtils variant by cpuid (DistanceUtils.h:118-163); on a host
without AVX-512 the AVX / SSE variants round differently (8 or 4 accumulators instead of 16).  The device reproduces
those summation trees too (float, int8, uint8): search ids / distance bits / counters, the stand-alone distance kernel
and a refine pass against the oracle's restatement of the same trees (pinned to the compiled reference's
ComputeL2Distance_AVX / _SSE etc. in tests/test_oracle_pin.py)."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("width", [8, 4])
@pytest.mark.parametrize("name", ["bkt_l2_10k_128", "bkt_cos_3k_768", "bkt_l2_3k_30", "bkt_l2_5k_100", "kdt_l2_10k_64",
                                  "bkt_i8_l2_5k_100", "bkt_i8_cos_6k_64", "bkt_u8_l2_6k_128"])
def test_search_with_avx_and_sse_trees(name, width):
    from sptag_b200 import B200Index, capi
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:80]
    idx = B200Index.load(folder)
    try:
        idx.set_param("B200.SimdWidth", width)
        for mc in (2048, 256):
            idx.set_param("MaxCheck", mc)
            ids, dists, stats = idx.search(q, 10, want_stats=True)
            o = reflib.OracleIndex(files, simd_width=width)
            o.max_check = mc
            ids_o, d_o, st_o = o.search(q, 10)
            assert np.array_equal(ids, ids_o), (name, width, mc)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32)), (name, width, mc)
            assert np.array_equal(stats[:, capi.ST_CHECKED], st_o[:, reflib.ST_CHECKED])
    finally:
        idx.close()


@pytest.mark.parametrize("width", [8, 4])
def test_distance_kernel_all_dims_avx_sse(width):
    """Every tail combination of the float trees through sptag_b200_distance_batch."""
    from sptag_b200 import B200Index, capi
    rng = np.random.default_rng(7)
    for dim in (1, 3, 4, 7, 8, 15, 16, 17, 31, 33, 100, 128, 131):
        n = 64
        x = rng.standard_normal((n, dim)).astype(np.float32)
        qv = rng.standard_normal((4, dim)).astype(np.float32)
        graph = np.full((n, 4), -1, np.int32)
        nodes = np.array([[n, 1, 2], [0, -1, -1], [-1, -1, -1]], np.int32)
        idx = B200Index.create(algo=capi.ALGO_BKT, value_type=capi.VT_FLOAT, metric=capi.METRIC_L2, vectors=x, graph=graph,
                               tree_starts=np.array([0], np.int32), tree_nodes=nodes)
        try:
            idx.set_param("B200.SimdWidth", width)
            ids = np.tile(np.arange(n, dtype=np.int32), (4, 1))
            out = idx.distance_batch(qv, ids)
            for qi in range(4):
                for v in range(n):
                    e = reflib.ora().ora_distance(0, reflib.VT_FLOAT, width, qv[qi].ctypes.data, x[v].ctypes.data, dim)
                    assert np.float32(e).view(np.int32) == out[qi, v].view(np.int32), (dim, width)
        finally:
            idx.close()


def test_refine_pass_with_avx_tree():
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_5k_100")
    files = reflib.IndexFiles(folder)
    idx = B200Index.load(folder)
    try:
        idx.set_param("B200.SimdWidth", 8)
        idx.set_param("MaxCheckForRefineGraph", 512)
        rows, ids, dists = idx.refine_graph(40, first=100, num=300, want_results=True)
        o = reflib.OracleIndex(files, simd_width=8)
        o.max_check_refine = 512
        rows_o, ids_o, d_o = o.refine_nodes(100, 300, 40, files.degree, 1.0)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
        assert np.array_equal(rows, rows_o)
    finally:
        idx.close()


def test_unsupported_width_combinations_are_refused():
    from sptag_b200 import B200Index, capi
    idx = B200Index.load(data_folder("bkt_i16_l2_5k_64"))
    try:
        idx.set_param("B200.SimdWidth", 8)
        q = np.load(os.path.join(data_folder("bkt_i16_l2_5k_64"), "queries.npy"))[:4]
        with pytest.raises(capi.SptagB200Error) as e:
            idx.search(q, 5)
        assert e.value.code == 0x13
    finally:
        idx.close()
