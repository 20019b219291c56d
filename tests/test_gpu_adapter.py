"""GPU test: the C++ host-side mirror of VectorIndex/QueryResult/BasicResult (vector_index_adapter.hpp)
returns exactly what the oracle returns, when driven the way the reference's callers drive VectorIndex."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_adapter_matches_oracle():
    import __graft_entry__
    exe = __graft_entry__.build_adapter_test()
    folder = data_folder("bkt_cos_10k_128")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:64]
    k, mc = 10, 1024
    with tempfile.TemporaryDirectory() as tmp:
        qf, of = os.path.join(tmp, "q.f32"), os.path.join(tmp, "out.bin")
        q.astype(np.float32).tofile(qf)
        subprocess.check_call([exe, folder, qf, str(q.shape[0]), str(q.shape[1]), str(k), str(mc), of])
        raw = np.fromfile(of, dtype=np.int32).reshape(q.shape[0], k, 2)
    ids = raw[:, :, 0]
    dist_bits = raw[:, :, 1]
    o = reflib.OracleIndex(files)
    o.max_check = mc
    ids_o, d_o, _ = o.search(q, k)
    assert np.array_equal(ids, ids_o)
    assert np.array_equal(dist_bits, d_o.view(np.int32))


def test_cpp_adapter_refine_calls_match_oracle():
    """RefineSearchIndex / one RefineGraph pass through the C++ mirror (row f2) against the oracle."""
    import __graft_entry__
    exe = __graft_entry__.build_adapter_test()
    folder = data_folder("bkt_l2_3k_30")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:8]
    cef, nodes = 40, 500
    with tempfile.TemporaryDirectory() as tmp:
        qf, of, rf = os.path.join(tmp, "q.f32"), os.path.join(tmp, "out.bin"), os.path.join(tmp, "refine.bin")
        q.astype(np.float32).tofile(qf)
        subprocess.check_call([exe, folder, qf, str(q.shape[0]), str(q.shape[1]), "10", "1024", of, rf, str(cef),
                               str(nodes)])
        raw = np.fromfile(rf, dtype=np.int32)
    rows = raw[:nodes * files.degree].reshape(nodes, files.degree)
    pairs = raw[nodes * files.degree:nodes * files.degree + 2 * (cef + 1)].reshape(cef + 1, 2)
    triples = raw[nodes * files.degree + 2 * (cef + 1):].reshape(10, 3)
    o = reflib.OracleIndex(files)   # MaxCheckForRefineGraph from the index's ini, like the adapter's loader
    rows_o, ids_o, d_o = o.refine_nodes(0, nodes, cef, files.degree, 1.0)
    assert np.array_equal(rows, rows_o)
    assert np.array_equal(pairs[:, 0], ids_o[0])
    assert np.array_equal(pairs[:, 1], d_o[0].view(np.int32))
    # GetIterator / ResultIterator::Next(5) x 2 on the first query (MaxCheck 1024 as set by the driver)
    o.max_check = 1024
    oi = o.iterator(q[0])
    for rnd in range(2):
        c, ids_i, d_i, relaxed = oi.next(5)
        assert c == 5
        assert np.array_equal(triples[5 * rnd:5 * rnd + 5, 0], ids_i)
        assert np.array_equal(triples[5 * rnd:5 * rnd + 5, 1], d_i.view(np.int32))
        assert (triples[5 * rnd:5 * rnd + 5, 2] != 0).all() == relaxed
    oi.close()
