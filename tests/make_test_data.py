"""Generate the small reference index folders used by the parity tests (TEST INFRASTRUCTURE).

Every index is BUILT AND SAVED BY THE UNMODIFIED REFERENCE (oracle/_ref/libsptag_ref.so:
VectorIndex::BuildIndex + SaveIndex), so the files are exactly what the reference's own
LoadIndex consumes.  Output goes to tests/_data/<name>/ (git-ignored, travels to the GPU box
with the gpurun snapshot).  Run:  python tests/make_test_data.py [name ...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reflib  # noqa: E402


def _line_data():
    # Test/src/AlgoTest.cpp:163-201: n=2000, dim=10, vec[i] = (i, ..., i)
    n, m = 2000, 10
    return (np.arange(n, dtype=np.float32)[:, None] * np.ones((1, m), np.float32)).copy()


def _dup_data():
    # many exact duplicates -> BKT duplicate groups (BKTree.h:598-609) and graph back-pointers
    base = reflib.gen_iid(1500, 24, 11)
    reps = np.concatenate([base, base[:400], base[:400], base[100:160]])
    rng = np.random.default_rng(12)
    return reps[rng.permutation(reps.shape[0])].copy()


# name -> (algo, metric, data generator, queries generator, build params)
SPECS = {
    "algo_line_bkt": ("BKT", "L2", _line_data, lambda: np.array([[0] * 10, [2] * 10, [4] * 10], np.float32), ""),
    "bkt_l2_20k_32": ("BKT", "L2", lambda: reflib.gen_iid(20000, 32, 1), lambda: reflib.gen_iid(500, 32, 2), ""),
    "bkt_cos_10k_128": ("BKT", "Cosine", lambda: reflib.gen_lowrank(10000, 128, 16, 3),
                        lambda: reflib.normalize_rows(reflib.gen_lowrank(300, 128, 16, 4)), ""),
    "bkt_l2_10k_128": ("BKT", "L2", lambda: reflib.gen_iid(10000, 128, 5), lambda: reflib.gen_iid(300, 128, 6), ""),
    "bkt_l2_5k_100": ("BKT", "L2", lambda: reflib.gen_iid(5000, 100, 7), lambda: reflib.gen_iid(200, 100, 8), ""),
    "bkt_l2_3k_30": ("BKT", "L2", lambda: reflib.gen_iid(3000, 30, 17), lambda: reflib.gen_iid(200, 30, 18), ""),
    "bkt_cos_3k_768": ("BKT", "Cosine", lambda: reflib.gen_lowrank(3000, 768, 32, 9),
                       lambda: reflib.normalize_rows(reflib.gen_lowrank(100, 768, 32, 10)), ""),
    "bkt_l2_dups": ("BKT", "L2", _dup_data, lambda: reflib.gen_iid(200, 24, 13), ""),
    # integer element types (DistanceUtils int8 / uint8 variants; PerfTest.cpp uses int8 cosine)
    "bkt_i8_cos_6k_64": ("BKT", "Cosine", lambda: _int8_lowrank(6000, 64, 10, 61), lambda: _norm_i8(_int8_lowrank(200, 64, 10, 62)), ""),
    "bkt_u8_l2_6k_128": ("BKT", "L2", lambda: _uint8_lowrank(6000, 128, 12, 63), lambda: _uint8_lowrank(200, 128, 12, 64), ""),
    "bkt_i8_l2_5k_100": ("BKT", "L2", lambda: _int8_lowrank(5000, 100, 12, 65), lambda: _int8_lowrank(200, 100, 12, 66), ""),
    "kdt_i8_l2_6k_32": ("KDT", "L2", lambda: _int8_lowrank(6000, 32, 8, 67), lambda: _int8_lowrank(200, 32, 8, 68), ""),
    # int16 rows (DistanceUtils int16 variants; squares of differences exceed 2^24, every rounding step shows)
    "bkt_i16_l2_5k_64": ("BKT", "L2", lambda: _int16_lowrank(5000, 64, 10, 81), lambda: _int16_lowrank(200, 64, 10, 82), ""),
    "bkt_i16_cos_5k_40": ("BKT", "Cosine", lambda: _int16_lowrank(5000, 40, 8, 83), lambda: _norm_i16(_int16_lowrank(200, 40, 8, 84)), ""),
    "bkt_i16_l2_4k_27": ("BKT", "L2", lambda: _int16_lowrank(4000, 27, 8, 85), lambda: _int16_lowrank(200, 27, 8, 86), ""),
    "kdt_i16_l2_5k_32": ("KDT", "L2", lambda: _int16_lowrank(5000, 32, 8, 87), lambda: _int16_lowrank(200, 32, 8, 88), ""),
    # more than one space-partition tree (BKTNumber / KDTNumber; the reference's ReconstructIndexSimilarityTest uses KDTNumber=2)
    "bkt2_l2_6k_32": ("BKT", "L2", lambda: reflib.gen_iid(6000, 32, 71), lambda: reflib.gen_iid(200, 32, 72), "BKTNumber=2"),
    "kdt2_l2_6k_32": ("KDT", "L2", lambda: reflib.gen_iid(6000, 32, 73), lambda: reflib.gen_iid(200, 32, 74), "KDTNumber=2"),
    "kdt_l2_10k_64": ("KDT", "L2", lambda: reflib.gen_iid(10000, 64, 14), lambda: reflib.gen_iid(300, 64, 15), ""),
    # 512-byte rows under a KD-tree: the fast-path KDT kernel (search_kernel<128, ..., KDT>)
    "kdt_l2_8k_128": ("KDT", "L2", lambda: reflib.gen_lowrank(8000, 128, 16, 91), lambda: reflib.gen_lowrank(300, 128, 16, 92), ""),
}


def _uint8_lowrank(n, dim, rank, seed):
    return np.clip(np.round(40.0 * reflib.gen_lowrank(n, dim, rank, seed) + 128), 0, 255).astype(np.uint8)


def _norm_i8(x):
    # the caller-side normalisation the reference expects for cosine queries: Utils::Normalize (CommonUtils.h:62-76),
    # arr[j] = (T)(arr[j] / |arr| * base), base = 127, C cast = truncation toward zero
    v = x.astype(np.float64)
    n = np.sqrt((v * v).sum(1, keepdims=True))
    return np.trunc(v / n * 127).astype(np.int8)


def _int16_lowrank(n, dim, rank, seed):
    return np.clip(np.round(6000.0 * reflib.gen_lowrank(n, dim, rank, seed)), -32767, 32767).astype(np.int16)


def _norm_i16(x):
    # Utils::Normalize for int16 (CommonUtils.h:62-76): base = 32767, truncation toward zero
    v = x.astype(np.float64)
    n = np.sqrt((v * v).sum(1, keepdims=True))
    return np.trunc(v / n * 32767).astype(np.int16)


def _int8_lowrank(n, dim, rank, seed):
    # SPACEV-style int8 raw vectors: clamp(round(32 * x), -127, 127)
    return np.clip(np.round(32.0 * reflib.gen_lowrank(n, dim, rank, seed)), -127, 127).astype(np.int8)


# quantized indexes: name -> (raw data gen, raw query gen, M, opq, reconstruct type)
QSPECS = {
    "bkt_pq_6k_32": (lambda: reflib.gen_lowrank(6000, 32, 8, 41), lambda: reflib.gen_lowrank(200, 32, 8, 42), 8, False,
                     reflib.VT_FLOAT),
    "bkt_opq_6k_48": (lambda: reflib.gen_lowrank(6000, 48, 10, 43), lambda: reflib.gen_lowrank(200, 48, 10, 44), 16, True,
                      reflib.VT_FLOAT),
    "bkt_opq_i8_8k_100": (lambda: _int8_lowrank(8000, 100, 16, 45), lambda: _int8_lowrank(200, 100, 16, 46), 50, True,
                          reflib.VT_INT8),
}


def make_quantized(name, force=False):
    folder = os.path.join(reflib.DATA_DIR, name)
    if os.path.exists(os.path.join(folder, "indexloader.ini")) and os.path.exists(
            os.path.join(folder, "queries.npy")) and not force:
        return folder
    if not reflib.have_ref():
        raise RuntimeError("oracle/_ref/libsptag_ref.so missing")
    gen_data, gen_q, m, opq, rtype = QSPECS[name]
    raw = np.ascontiguousarray(gen_data())
    t = time.time()
    os.makedirs(folder, exist_ok=True)
    qz = reflib.train_quantizer(raw.astype(np.float32), m=m, ks=256, opq=opq, rtype=rtype, seed=5, iters=4)
    qpath = os.path.join(folder, "quantizer_src.bin")
    qz.write(qpath)
    codes = reflib.RefQuantizer(qpath).encode(raw)          # the reference's own QuantizeVector
    idx = reflib.RefIndex.build_quantized("BKT", codes, "L2", qpath, threads=os.cpu_count() or 8)
    idx.save(folder)                                         # writes quantizer.bin + [Quantizer] ini section
    np.save(os.path.join(folder, "queries.npy"), np.ascontiguousarray(gen_q()))
    print("built %-18s n=%d raw dim=%d M=%d in %.1fs" % (name, raw.shape[0], raw.shape[1], m, time.time() - t), flush=True)
    return folder


# indexes with tombstones: name -> (base spec name of SPECS-like tuple, fraction deleted, seed).  Built by the
# reference, then VectorIndex::DeleteIndex(id) on a random subset plus the exact nearest neighbours of some queries
# (so that deleted vectors are popped, skipped as results and still expanded), then SaveIndex -> deletes.bin
DSPECS = {
    "bkt_l2_deleted_6k_32": ("BKT", "L2", lambda: reflib.gen_lowrank(6000, 32, 8, 91), lambda: reflib.gen_lowrank(200, 32, 8, 92), 0.25, 93),
    "bkt_cos_deleted_5k_64": ("BKT", "Cosine", lambda: reflib.gen_lowrank(5000, 64, 10, 94),
                              lambda: reflib.normalize_rows(reflib.gen_lowrank(200, 64, 10, 95)), 0.3, 96),
}


def make_deleted(name, force=False):
    folder = os.path.join(reflib.DATA_DIR, name)
    if os.path.exists(os.path.join(folder, "indexloader.ini")) and os.path.exists(
            os.path.join(folder, "queries.npy")) and not force:
        return folder
    if not reflib.have_ref():
        raise RuntimeError("oracle/_ref/libsptag_ref.so missing")
    algo, metric, gen_data, gen_q, frac, seed = DSPECS[name]
    data = np.ascontiguousarray(gen_data())
    q = np.ascontiguousarray(gen_q())
    t = time.time()
    idx = reflib.RefIndex.build(algo, data, metric, threads=os.cpu_count() or 8)
    rng = np.random.default_rng(seed)
    dele = set(np.nonzero(rng.random(data.shape[0]) < frac)[0].tolist())
    ids, _, _ = idx.search(q[:60], 3, threads=4)          # the top-3 of the first 60 queries go too
    dele.update(int(v) for v in ids.ravel() if v >= 0)
    idx.delete(np.array(sorted(dele), np.int32))
    idx.save(folder)
    np.save(os.path.join(folder, "queries.npy"), q)
    print("built %-18s n=%d dim=%d deleted=%d in %.1fs" % (name, data.shape[0], data.shape[1], len(dele), time.time() - t),
          flush=True)
    return folder


def make(name, force=False):
    if name in QSPECS:
        return make_quantized(name, force)
    if name in DSPECS:
        return make_deleted(name, force)
    folder = os.path.join(reflib.DATA_DIR, name)
    if os.path.exists(os.path.join(folder, "indexloader.ini")) and os.path.exists(
            os.path.join(folder, "queries.npy")) and not force:
        return folder
    if not reflib.have_ref():
        raise RuntimeError("oracle/_ref/libsptag_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    algo, metric, gen_data, gen_q, params = SPECS[name]
    data = np.ascontiguousarray(gen_data())
    t = time.time()
    idx = reflib.RefIndex.build(algo, data, metric, threads=os.cpu_count() or 8, params=params)
    idx.save(folder)
    np.save(os.path.join(folder, "queries.npy"), np.ascontiguousarray(gen_q()))
    print("built %-18s n=%d dim=%d in %.1fs" % (name, data.shape[0], data.shape[1], time.time() - t), flush=True)
    return folder


if __name__ == "__main__":
    names = sys.argv[1:] or (list(SPECS) + list(QSPECS) + list(DSPECS))
    for nm in names:
        make(nm)
