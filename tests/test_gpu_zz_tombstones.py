"""GPU parity on reference-written tombstones (deletes.bin read by sptag_b200_load): indexes the
reference built, deleted ~30 % of with VectorIndex::DeleteIndex (including true nearest neighbours of the first
queries) and saved.  Search, one refine step and iterator scans against the oracle, whose tombstone handling is pinned
to the reference on the same folders (tests/test_oracle_pin.py::test_tombstones_bit_exact_vs_reference).
Also closes the builder-side loop: a device-refined graph written as graph.bin and searched by the unmodified reference."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu

SETS = ["bkt_l2_deleted_6k_32", "bkt_cos_deleted_5k_64"]


@pytest.mark.parametrize("name", SETS)
def test_search_with_reference_written_tombstones(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    assert files.num_deleted > 1000
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        o = reflib.OracleIndex(files)
        for mc in (8192, 512, 64):
            idx.set_param("MaxCheck", mc)
            o.max_check = mc
            ids, dists = idx.search(q, 10)
            ids_o, d_o, _ = o.search(q, 10)
            assert np.array_equal(ids, ids_o), (name, mc)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32)), (name, mc)
            assert not (files.deleted[ids[ids >= 0]] == 1).any()
    finally:
        idx.close()


@pytest.mark.parametrize("name", SETS)
def test_refine_with_tombstones(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheckForRefineGraph", 512)
        rows, ids, dists = idx.refine_graph(48, first=100, num=400, want_results=True)
        o = reflib.OracleIndex(files)
        o.max_check_refine = 512
        rows_o, ids_o, d_o = o.refine_nodes(100, 400, 48, files.degree, 1.0)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
        assert np.array_equal(rows, rows_o)
    finally:
        idx.close()


@pytest.mark.parametrize("name", SETS)
def test_iterator_with_tombstones(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:16]
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheck", 256)
        o = reflib.OracleIndex(files)
        o.max_check = 256
        oits = [o.iterator(qq) for qq in q]
        its = idx.iterators(q)
        for b in [10, 10, 5, 10, 10, 10]:
            counts, ids, dists, relaxed = its.next(b)
            for i, oi in enumerate(oits):
                c, io, do, ro = oi.next(b)
                assert counts[i] == c and bool(relaxed[i]) == ro, (name, i, b)
                assert np.array_equal(ids[i], io), (name, i, b)
                assert np.array_equal(dists[i].view(np.int32), do.view(np.int32)), (name, i, b)
        its.close()
        for oi in oits:
            oi.close()
    finally:
        idx.close()


def test_device_refined_graph_round_trips_through_the_reference(tmp_path):
    """Builder-side loop closed: refine + install on the device, write graph.bin (NeighborhoodGraph::SaveGraph format)
    next to the other files of the folder, and let the UNMODIFIED REFERENCE load that folder (oracle/_ref travels to
    the GPU box): its SearchIndex on the device-built graph must equal the device's own search on it."""
    import shutil
    from sptag_b200 import B200Index
    if not reflib.have_ref():
        pytest.skip("oracle/_ref not built")
    folder = data_folder("bkt_l2_5k_100")
    q = np.load(os.path.join(folder, "queries.npy"))
    out = str(tmp_path / "refined")
    shutil.copytree(folder, out)
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheckForRefineGraph", 1024)
        idx.refine_graph(64, install=True, want_rows=False)
        g = idx.save_graph(os.path.join(out, "graph.bin"))
        files = reflib.IndexFiles(out)
        assert np.array_equal(files.graph, g)
        r = reflib.RefIndex.load(out)
        for mc in (2048, 256):
            idx.set_param("MaxCheck", mc)
            r.set_param("MaxCheck", mc)
            ids, dists = idx.search(q, 10)
            ids_r, d_r, _ = r.search(q, 10, threads=4)
            assert np.array_equal(ids, ids_r), mc
            assert np.array_equal(dists.view(np.int32), d_r.view(np.int32)), mc
    finally:
        idx.close()
