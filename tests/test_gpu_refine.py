"""GPU parity for SURVEY.md 8 row f2 (graph refinement on the device): sptag_b200_refine_graph -- one
NeighborhoodGraph::RefineNode pass = RefineSearchIndex + RelativeNeighborhoodGraph::RebuildNeighbors -- against the
oracle's restatement (oracle/sptag_oracle.c ora_refine_nodes, pinned to the reference's own RefineSearchIndex +
RebuildNeighbors in tests/test_oracle_pin.py): new graph rows, refine-search ids and distances all bit-exact."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu

# (index, CEF, MaxCheckForRefineGraph, nodes)
CASES = [
    ("algo_line_bkt", 1000, 8192, 400),
    ("bkt_l2_dups", 20, 256, 600),          # duplicate groups: searchDuplicated = false adds one member per group
    ("bkt_l2_20k_32", 100, 2048, 500),
    ("bkt_cos_3k_768", 1000, 8192, 300),    # the reference's default CEF: K = 1001 result heap in HBM
    ("bkt_l2_3k_30", 64, 1024, 400),        # dim % 4 != 0: scalar tails in the node-to-node distances
    ("bkt2_l2_6k_32", 40, 512, 400),
    ("kdt_l2_10k_64", 64, 1024, 400),
    ("bkt_i8_cos_6k_64", 50, 512, 400),
    ("bkt_u8_l2_6k_128", 50, 512, 400),
    ("bkt_i8_l2_5k_100", 31, 300, 400),
    ("bkt_i16_l2_4k_27", 40, 512, 400),
    ("bkt_i16_cos_5k_40", 40, 512, 400),
    ("kdt_i16_l2_5k_32", 40, 512, 400),
    # quantized indexes (NeighborhoodGraph.h:538-543): reconstruct the node's code row, quantize the reconstruction again,
    # search with that, RebuildNeighbors through the quantizer's SDC distance
    ("bkt_pq_6k_32", 64, 1024, 400),
    ("bkt_opq_6k_48", 100, 2048, 400),
    ("bkt_opq_i8_8k_100", 40, 512, 400),
    ("bkt_opq_i8_8k_100", 1000, 8192, 200),   # K = 1001: the result set in HBM
]


@pytest.mark.parametrize("name,cef,mcr,num", CASES)
def test_refine_pass_bit_exact(name, cef, mcr, num):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    num = min(num, files.n)
    first = max(0, files.n // 2 - num // 2)
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheckForRefineGraph", mcr)
        rows, ids, dists = idx.refine_graph(cef, first=first, num=num, want_results=True)
        o = reflib.OracleIndex(files)
        o.max_check_refine = mcr
        rows_o, ids_o, d_o = o.refine_nodes(first, num, cef, files.degree, 1.0)
        assert np.array_equal(ids, ids_o), name
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32)), name
        assert np.array_equal(rows, rows_o), name
        # the public search path is untouched by a pass that was not installed
        assert np.array_equal(idx.get_graph(), files.graph)
    finally:
        idx.close()


@pytest.mark.parametrize("neighborhood,factor", [(16, 1.0), (32, 1.5), (48, 0.8)])
def test_refine_neighbourhood_and_rng_factor(neighborhood, factor):
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_10k_128")
    files = reflib.IndexFiles(folder)
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheckForRefineGraph", 1024)
        rows = idx.refine_graph(100, first=100, num=300, neighborhood=neighborhood, rng_factor=factor)
        o = reflib.OracleIndex(files)
        o.max_check_refine = 1024
        rows_o, _, _ = o.refine_nodes(100, 300, 100, neighborhood, factor)
        assert np.array_equal(rows, rows_o)
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["bkt_l2_dups", "bkt_l2_5k_100"])
def test_refine_install_then_search(name):
    """A full pass installed on the device: the graph becomes the oracle's pass output (+ the duplicate-group
    back-pointers BuildGraph re-attaches, NeighborhoodGraph.h:395-401), and searching the refined index equals the
    oracle searching the same refined graph."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheckForRefineGraph", 512)
        rows = idx.refine_graph(48, install=True)
        o = reflib.OracleIndex(files)
        o.max_check_refine = 512
        rows_o, _, _ = o.refine_nodes(0, files.n, 48, files.degree, 1.0)
        assert np.array_equal(rows, rows_o)
        expect = rows_o.copy()
        back = files.graph[:, -1] < -1
        expect[back, -1] = files.graph[back, -1]
        assert np.array_equal(idx.get_graph(), expect)
        old_graph = files.graph
        files.graph = np.ascontiguousarray(expect)
        try:
            o2 = reflib.OracleIndex(files)
            for mc in (2048, 128):
                o2.max_check = mc
                idx.set_param("MaxCheck", mc)
                ids, dists = idx.search(q, 10)
                ids_o, d_o, _ = o2.search(q, 10)
                assert np.array_equal(ids, ids_o)
                assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
        finally:
            files.graph = old_graph
    finally:
        idx.close()


@pytest.mark.parametrize("name,iters,cef,cef_scale,nscale,mcr", [
    ("bkt_l2_5k_100", 2, 1000, 2.0, 2.0, 2048),   # the reference's defaults: first pass CEF x CEFScale = 2000 on 64-wide rows
    ("bkt_l2_dups", 3, 24, 2.0, 2.0, 256),        # duplicate back-pointers carried across the width changes
    ("kdt_l2_10k_64", 2, 40, 1.5, 2.0, 512),
])
def test_refine_schedule_matches_pass_by_pass_oracle(name, iters, cef, cef_scale, nscale, mcr):
    """NeighborhoodGraph::RefineGraph (NeighborhoodGraph.h:460-492): RefineIterations - 1 passes with CEF x CEFScale on
    rows NeighborhoodScale times wider, then one pass with CEF on NeighborhoodSize rows.  Each device pass reads the
    previous pass's graph; the oracle restates the same frozen-graph passes one after the other."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:60]
    idx = B200Index.load(folder)
    old_graph, old_degree = files.graph, files.degree
    try:
        idx.set_param("MaxCheckForRefineGraph", mcr)
        idx.refine_schedule(iters, cef, cef_scale, old_degree, nscale, 1.0)
        wide = int(old_degree * nscale)
        narrow = int(wide / nscale)
        back = old_graph[:, -1] < -1
        for it in range(iters):
            last = it == iters - 1
            width = narrow if last else wide
            o = reflib.OracleIndex(files)
            o.max_check_refine = mcr
            rows, _, _ = o.refine_nodes(0, files.n, cef if last else int(cef * cef_scale), width, 1.0)
            rows[back, -1] = old_graph[back, -1]
            files.graph = np.ascontiguousarray(rows)
            files.degree = width
        assert idx.graph_degree == narrow
        assert np.array_equal(idx.get_graph(), files.graph)
        o2 = reflib.OracleIndex(files)
        o2.max_check = 1024
        idx.set_param("MaxCheck", 1024)
        ids, dists = idx.search(q, 10)
        ids_o, d_o, _ = o2.search(q, 10)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
    finally:
        files.graph, files.degree = old_graph, old_degree
        idx.close()


@pytest.mark.parametrize("name,widen_to", [("bkt_l2_20k_32", 0), ("bkt_l2_dups", 0), ("kdt_l2_10k_64", 0),
                                           ("bkt_l2_5k_100", 48), ("bkt_u8_l2_6k_128", 20)])
def test_rebuild_graph_matches_oracle(name, widen_to):
    """EnableRebuild's in-degree repair (NeighborhoodGraph::RebuildGraph, NeighborhoodGraph.h:404-456) on the device against
    the oracle's restatement of its single-thread order (pinned to the reference itself in tests/test_oracle_pin.py); on
    the index's own 32-wide rows (N = 16) and on rows a refine pass installed with another width; then installed and
    searched."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:100]
    idx = B200Index.load(folder)
    old_graph, old_degree = files.graph, files.degree
    try:
        if widen_to:
            idx.set_param("MaxCheckForRefineGraph", 512)
            idx.refine_graph(60, neighborhood=widen_to, install=True)
        wide = idx.get_graph()
        n2 = wide.shape[1] // 2
        g = wide.copy()
        g[g < -1] = -1                      # back-pointers are not neighbours; BuildGraph attaches them afterwards
        expect = reflib.oracle_rebuild_graph(g, n2)[:, :n2]
        rows = idx.rebuild_graph(install=True)
        assert np.array_equal(rows, expect), name
        back = wide[:, -1] < -1
        expect = expect.copy()
        expect[back, -1] = wide[back, -1]
        assert idx.graph_degree == n2
        assert np.array_equal(idx.get_graph(), expect)
        files.graph, files.degree = np.ascontiguousarray(expect), n2
        o = reflib.OracleIndex(files)
        for mc in (1024, 64):
            o.max_check = mc
            idx.set_param("MaxCheck", mc)
            ids, dists = idx.search(q, 10)
            ids_o, d_o, _ = o.search(q, 10)
            assert np.array_equal(ids, ids_o)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
    finally:
        files.graph, files.degree = old_graph, old_degree
        idx.close()


def test_refine_argument_errors():
    from sptag_b200 import B200Index, capi
    idx = B200Index.load(data_folder("bkt_l2_3k_30"))
    try:
        with pytest.raises(capi.SptagB200Error):
            idx.refine_graph(2048)                       # K = CEF + 1 > 2048
        with pytest.raises(capi.SptagB200Error):
            idx.refine_graph(10, first=2990, num=100)    # past the end
        with pytest.raises(capi.SptagB200Error):
            idx.refine_graph(10, first=0, num=10, install=True)   # install needs a full pass
    finally:
        idx.close()
    q = B200Index.load(data_folder("bkt_pq_6k_32"))
    try:
        q.set_param("EnableADC", 1)   # the reference's RebuildNeighbors is undefined with ADC on (reads a code row as a table)
        with pytest.raises(capi.SptagB200Error):
            q.refine_graph(10)
    finally:
        q.close()
