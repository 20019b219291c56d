"""CPU tests of the bench's set-up utility (tools/gpu_index_builder.py, run here with torch on the CPU): whatever it
writes must be a structurally valid SPTAG index that the oracle (and, where present, the unmodified reference)
can load and search -- the bench's parity check is only meaningful on such files."""
import os
import sys

import numpy as np
import pytest

import reflib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch = pytest.importorskip("torch")
from tools import gpu_index_builder as B  # noqa: E402


def _bkt_is_valid(nodes, n):
    # every data point is exactly one node; root centerid = N; a sentinel ends the array; children contiguous
    assert nodes.shape == (n + 2, 3)
    assert nodes[0, 0] == n and tuple(nodes[-1]) == (-1, -1, -1)
    assert sorted(nodes[1:-1, 0].tolist()) == list(range(n))
    internal = nodes[:-1, 1] > 0
    cs, ce = nodes[:-1, 1][internal], nodes[:-1, 2][internal]
    assert (ce > cs).all() and (ce - cs <= 32).all() and ce.max() == n + 1
    # every non-root node is the child of exactly one parent
    covered = np.zeros(n + 1, np.int32)
    for a, b in zip(cs.tolist(), ce.tolist()):
        covered[a:b] += 1
    assert (covered[1:] == 1).all()


def _graph_is_valid(graph, n):
    assert graph.shape == (n, 32)
    g = graph.astype(np.int64)
    assert ((g >= -1) & (g < n)).all()
    rows = np.arange(n)[:, None]
    assert not (g == rows).any()                      # no self loops
    valid = g >= 0
    assert (valid[:, :-1] | ~valid[:, 1:]).all()        # -1 padding only at the end of a row


@pytest.mark.parametrize("builder", ["kmeans", "balanced"])
def test_bkt_builders_produce_valid_trees(builder):
    x = torch.from_numpy(reflib.gen_lowrank(20011, 24, 6, 3))
    nodes, starts = (B.build_bkt(x) if builder == "kmeans" else B.build_bkt_balanced(x))
    assert starts.tolist() == [0]
    _bkt_is_valid(nodes, x.shape[0])


def test_kdt_builder_produces_a_valid_tree():
    x = torch.from_numpy(reflib.gen_lowrank(9001, 16, 5, 4))
    nodes, starts = B.build_kdt(x)
    n = x.shape[0]
    leaves = np.concatenate([nodes[:, 0][nodes[:, 0] < 0], nodes[:, 1][nodes[:, 1] < 0]])
    assert sorted((-leaves - 1).tolist()) == list(range(n))        # every point is exactly one leaf
    used = (nodes[:, 0] != 0) | (nodes[:, 1] != 0)
    assert used.sum() == n - 1                                      # a binary tree over n leaves
    assert ((nodes[used, 2] >= 0) & (nodes[used, 2] < x.shape[1])).all()


@pytest.mark.parametrize("graph_kind", ["brute", "tpt"])
def test_built_folder_is_searchable(oracle_lib, tmp_path, graph_kind):
    x = torch.from_numpy(reflib.gen_lowrank(12007, 32, 8, 5))
    q = reflib.gen_lowrank(100, 32, 8, 6)
    nodes, starts, graph = B.build_index(x, "L2", tpt_above=(10 ** 9 if graph_kind == "brute" else 1000), tpt_trees=5)
    _bkt_is_valid(nodes, x.shape[0])
    _graph_is_valid(graph, x.shape[0])
    folder = str(tmp_path / "idx")
    B.save_index_folder(folder, x.numpy(), graph, nodes, starts, "L2")
    files = reflib.IndexFiles(folder)
    o = reflib.OracleIndex(files)
    o.max_check = 2048
    ids, dists, _ = o.search(q, 10)
    truth = B.exact_topk(x, torch.from_numpy(q), 10, "L2")
    recall = np.mean([len(set(truth[i]) & set(ids[i])) / 10 for i in range(q.shape[0])])
    assert recall > 0.9
    if reflib.have_ref():                                           # and the unmodified reference agrees bit for bit
        r = reflib.RefIndex.load(folder)
        r.set_param("MaxCheck", 2048)
        ids_r, d_r, _ = r.search(q, 10, threads=2)
        assert np.array_equal(ids_r, ids) and np.array_equal(d_r.view(np.int32), dists.view(np.int32))


def test_quantized_folder_round_trips(oracle_lib, tmp_path):
    x = torch.from_numpy(np.clip(np.round(32 * reflib.gen_lowrank(6000, 20, 6, 7)), -127, 127).astype(np.float32))
    cb, rot = B.train_quantizer_gpu(x, 10, iters=2, sample=3000)
    codes = B.encode_gpu(x, cb, rot)
    nodes, starts, graph = B.build_index(x, "L2")
    folder = str(tmp_path / "qidx")
    B.save_index_folder(folder, codes.numpy(), graph, nodes, starts, "L2", quantizer=B.quantizer_blob(cb, rot, 0))
    files = reflib.IndexFiles(folder)
    assert files.value_type == reflib.VT_UINT8 and files.quantizer is not None and files.quantizer.qtype == reflib.Q_OPQ
    q = np.clip(np.round(32 * reflib.gen_lowrank(50, 20, 6, 8)), -127, 127).astype(np.int8)
    ids, _, _ = reflib.OracleIndex(files).search(q, 10)
    assert (ids >= 0).all()
