"""Golden fixtures for the graph-refinement row -- TEST INFRASTRUCTURE.

For every committed search fixture tests/golden/<name>.npz (an index built by the unmodified reference) this writes
tests/golden/refine/<name>.npz with the UNMODIFIED REFERENCE's outputs of one RefineNode step per node on that index:
VectorIndex::RefineSearchIndex result lists (ids, distances) and RelativeNeighborhoodGraph::RebuildNeighbors rows
(oracle/ref_shim.cpp ref_refine_nodes).  /root/reference is not needed to USE the fixtures.
Run (where oracle/_ref exists):  python tests/golden/make_golden_refine.py
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

# name -> (CEF, MaxCheckForRefineGraph, neighbourhood size, RNGFactor, nodes refined: the first `num`)
SETTINGS = {
    "algo_line_bkt": (1000, 8192, 32, 1.0, 200),
    "bkt_l2_2k_16": (64, 1024, 32, 1.0, 600),
    "bkt_cos_1500_20": (100, 2048, 32, 1.0, 400),
    "bkt_l2_dups_1k_12": (24, 256, 32, 1.0, 1030),
    "kdt_l2_2k_16": (48, 512, 16, 1.25, 600),
}


def make(name):
    from tools.gpu_index_builder import save_index_folder
    g = np.load(os.path.join(HERE, name + ".npz"))
    params = dict(zip(g["param_names"].tolist(), g["param_values"].tolist()))
    cef, mcr, nbh, factor, num = SETTINGS[name]
    with tempfile.TemporaryDirectory() as tmp:
        save_index_folder(tmp, g["vectors"], g["graph"], g["nodes"], g["tree_starts"], params["DistCalcMethod"],
                          algo=params["IndexAlgoType"], value_type=params["ValueType"])
        files = reflib.IndexFiles(tmp)
        assert np.array_equal(files.graph, g["graph"]) and np.array_equal(files.nodes, g["nodes"])
        r = reflib.RefIndex.load(tmp)
        r.set_param("MaxCheckForRefineGraph", mcr)
        rows, ids, dists = r.refine_nodes(0, min(num, files.n), cef, nbh, factor, threads=8)
    out = os.path.join(HERE, "refine", name + ".npz")
    np.savez_compressed(out, cef=np.int32(cef), max_check_refine=np.int32(mcr), neighborhood=np.int32(nbh),
                        rng_factor=np.float32(factor), rows=rows, res_ids=ids.astype(np.int32),
                        res_dists=dists, ref_isa=np.int32(reflib.ref().ref_isa()))
    print("golden refine", name, rows.shape, ids.shape, os.path.getsize(out))


if __name__ == "__main__":
    names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz")))
    for nm in names:
        make(nm)
