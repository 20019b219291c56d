"""Golden fixtures for the iterator row -- TEST INFRASTRUCTURE.

For every committed search fixture tests/golden/<name>.npz with a BKT index this writes tests/golden/iterator/<name>.npz
holding the UNMODIFIED REFERENCE's ResultIterator outputs on that index (oracle/ref_shim.cpp ref_iter_*): for the first
queries, a schedule of Next(batch) calls -> per call the result count, ids, distances and RelaxedMono.
/root/reference is not needed to USE the fixtures.  Run (where oracle/_ref exists):
    python tests/golden/make_golden_iterator.py
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

SCHEDULE = [8, 8, 5, 8, 3, 2, 6]   # the 4th and 7th requests are capped by the previous count (reference behaviour)
MAX_CHECK = {"algo_line_bkt": 5, "bkt_l2_2k_16": 256, "bkt_cos_1500_20": 1024, "bkt_l2_dups_1k_12": 64}
NQ = 16


def make(name):
    from tools.gpu_index_builder import save_index_folder
    g = np.load(os.path.join(HERE, name + ".npz"))
    params = dict(zip(g["param_names"].tolist(), g["param_values"].tolist()))
    if params["IndexAlgoType"] != "BKT":
        return
    mc = MAX_CHECK[name]
    q = np.ascontiguousarray(g["queries"][:NQ])
    width = max(SCHEDULE)
    with tempfile.TemporaryDirectory() as tmp:
        save_index_folder(tmp, g["vectors"], g["graph"], g["nodes"], g["tree_starts"], params["DistCalcMethod"],
                          algo="BKT", value_type=params["ValueType"])
        r = reflib.RefIndex.load(tmp)
        r.set_param("MaxCheck", mc)
        counts = np.zeros((q.shape[0], len(SCHEDULE)), np.int32)
        relaxed = np.zeros((q.shape[0], len(SCHEDULE)), np.uint8)
        ids = np.full((q.shape[0], len(SCHEDULE), width), -1, np.int32)
        dists = np.full((q.shape[0], len(SCHEDULE), width), np.float32(np.finfo(np.float32).max / np.float32(10)), np.float32)
        for i in range(q.shape[0]):
            it = r.iterator(q[i])
            for s, b in enumerate(SCHEDULE):
                c, a, d, rm = it.next(b)
                counts[i, s], relaxed[i, s] = c, rm
                ids[i, s, :b], dists[i, s, :b] = a, d
            it.close()
    out = os.path.join(HERE, "iterator", name + ".npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, max_check=np.int32(mc), schedule=np.array(SCHEDULE, np.int32), nq=np.int32(q.shape[0]),
                        counts=counts, relaxed=relaxed, ids=ids, dists=dists, ref_isa=np.int32(reflib.ref().ref_isa()))
    print("golden iterator", name, counts.sum(), os.path.getsize(out))


if __name__ == "__main__":
    names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz")))
    for nm in names:
        make(nm)
