"""Generate the committed golden fixtures (TEST INFRASTRUCTURE).

Each tests/golden/<name>.npz holds a tiny index BUILT BY THE UNMODIFIED REFERENCE
(oracle/_ref/libsptag_ref.so -> VectorIndex::BuildIndex/SaveIndex), flattened to arrays, plus the
reference's own outputs on it: ids / distances from VectorIndex::SearchIndex and the WorkSpace
counters (m_iNumberOfCheckedLeaves, m_iNumberOfTreeCheckedLeaves, NGQueue.size(), SPTQueue.size())
for several MaxCheck values.  /root/reference is not needed to USE the fixtures.
Run (where oracle/_ref exists):  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib  # noqa: E402


def line_data():
    return (np.arange(2000, dtype=np.float32)[:, None] * np.ones((1, 10), np.float32)).copy()


def dup_data():
    base = reflib.gen_iid(700, 12, 21)
    reps = np.concatenate([base, base[:150], base[:150], base[40:70]])
    return reps[np.random.default_rng(22).permutation(reps.shape[0])].copy()


CASES = {
    # name: (algo, metric, data, queries, k, max_checks)
    "algo_line_bkt": ("BKT", "L2", line_data, lambda: np.array([[0] * 10, [2] * 10, [4] * 10], np.float32), 3,
                      [8192, 64]),
    "bkt_l2_2k_16": ("BKT", "L2", lambda: reflib.gen_iid(2000, 16, 31), lambda: reflib.gen_iid(64, 16, 32), 10,
                     [8192, 512, 64]),
    "bkt_cos_1500_20": ("BKT", "Cosine", lambda: reflib.gen_lowrank(1500, 20, 6, 33),
                        lambda: reflib.normalize_rows(reflib.gen_lowrank(64, 20, 6, 34)), 10, [8192, 256]),
    "bkt_l2_dups_1k_12": ("BKT", "L2", dup_data, lambda: reflib.gen_iid(64, 12, 23), 10, [8192, 128]),
    "kdt_l2_2k_16": ("KDT", "L2", lambda: reflib.gen_iid(2000, 16, 35), lambda: reflib.gen_iid(64, 16, 36), 10,
                     [8192, 512, 64]),
}


def make(name):
    algo, metric, gen_data, gen_q, k, max_checks = CASES[name]
    data = np.ascontiguousarray(gen_data())
    q = np.ascontiguousarray(gen_q())
    with tempfile.TemporaryDirectory() as tmp:
        idx = reflib.RefIndex.build(algo, data, metric, threads=8)
        idx.save(tmp)
        files = reflib.IndexFiles(tmp)
        r = reflib.RefIndex.load(tmp)
        assert r.enable_stats() == 0
        ref_ids, ref_dists, ref_stats = [], [], []
        for mc in max_checks:
            r.set_param("MaxCheck", mc)
            ids = np.empty((q.shape[0], k), np.int32)
            dists = np.empty((q.shape[0], k), np.float32)
            stats = np.empty((q.shape[0], 4), np.int32)
            for i in range(q.shape[0]):
                ids[i], dists[i], stats[i] = r.search_one_stats(q[i], k)
            ref_ids.append(ids)
            ref_dists.append(dists)
            ref_stats.append(stats)
        names = sorted(files.params)
        out = dict(param_names=np.array(names), param_values=np.array([files.params[n] for n in names]),
                   vectors=files.vectors, graph=files.graph, tree_starts=files.tree_starts, nodes=files.nodes,
                   queries=q, k=np.int32(k), max_checks=np.array(max_checks, np.int32),
                   ref_ids=np.stack(ref_ids), ref_dists=np.stack(ref_dists), ref_stats=np.stack(ref_stats),
                   ref_isa=np.int32(reflib.ref().ref_isa()))
        if files.deleted is not None:
            out["deleted"] = files.deleted
            out["num_deleted"] = np.int32(files.num_deleted)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("golden", name, data.shape, os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        make(nm)
