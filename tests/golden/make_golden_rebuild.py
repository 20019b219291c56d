"""Golden fixtures for EnableRebuild's in-degree repair -- TEST INFRASTRUCTURE.

For every committed search fixture tests/golden/<name>.npz (an index built by the unmodified reference) this writes
tests/golden/rebuild/<name>.npz with the UNMODIFIED REFERENCE's NeighborhoodGraph::RebuildGraph (NeighborhoodGraph.h:404-456,
run with one thread -- the only order in which its result is a function of the input; oracle/ref_shim.cpp
ref_rebuild_graph) applied to the fixture's own graph rows read as 2 x N candidates, and to a copy with a skewed in-degree
and holes.  /root/reference is not needed to USE the fixtures.
Run (where oracle/_ref exists):  python tests/golden/make_golden_rebuild.py
"""
import glob
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import reflib  # noqa: E402


def inputs(graph, n):
    nbh = graph.shape[1] // 2
    g0 = graph.copy()
    g0[g0 < 0] = -1                      # duplicate back-pointers are attached after RebuildGraph (:395-401)
    rng = np.random.default_rng(20260921)
    g1 = g0.copy()
    g1[:, nbh // 2:nbh] = rng.integers(0, max(2, n // 40), size=(n, nbh - nbh // 2))
    g1[rng.random(g1.shape) < 0.05] = -1
    return nbh, g0, g1


def make(name):
    from tools.gpu_index_builder import save_index_folder
    g = np.load(os.path.join(HERE, name + ".npz"))
    params = dict(zip(g["param_names"].tolist(), g["param_values"].tolist()))
    with tempfile.TemporaryDirectory() as tmp:
        save_index_folder(tmp, g["vectors"], g["graph"], g["nodes"], g["tree_starts"], params["DistCalcMethod"],
                          algo=params["IndexAlgoType"], value_type=params["ValueType"])
        files = reflib.IndexFiles(tmp)
        r = reflib.RefIndex.load(tmp)
        nbh, g0, g1 = inputs(files.graph, files.n)
        out0 = r.rebuild_graph(g0, nbh)
        out1 = r.rebuild_graph(g1, nbh)
    os.makedirs(os.path.join(HERE, "rebuild"), exist_ok=True)
    out = os.path.join(HERE, "rebuild", name + ".npz")
    np.savez_compressed(out, neighborhood=np.int32(nbh), own_rows=out0[:, :nbh].astype(np.int32),
                        skewed_in=g1.astype(np.int32), skewed_rows=out1[:, :nbh].astype(np.int32))
    print("golden rebuild", name, out0.shape, int((out0[:, :nbh] != g0[:, :nbh]).any(axis=1).sum()), "rows changed,",
          os.path.getsize(out), "bytes")


if __name__ == "__main__":
    names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz")))
    for nm in names:
        make(nm)
