"""GPU parity tests (run with -m gpu on a B200): the CUDA search path, called through the C ABI,
against the oracle (oracle/sptag_oracle.c) on the same reference-built index files -- ids bit-exact,
distances bit-exact (same summation tree, tolerance 0 ulp; the north star allows 1e-4 relative),
and the per-query work counters equal to the reference's WorkSpace counters."""
import os

import numpy as np
import pytest

import reflib
from conftest import data_folder

pytestmark = pytest.mark.gpu

BKT_SETS = ["algo_line_bkt", "bkt_l2_20k_32", "bkt_cos_10k_128", "bkt_l2_10k_128", "bkt_l2_5k_100",
            "bkt_l2_3k_30", "bkt_cos_3k_768", "bkt_l2_dups", "bkt2_l2_6k_32"]


def _compare(idx, files, q, k, mc, tag):
    from sptag_b200 import capi
    idx.set_param("MaxCheck", mc)
    ids, dists, stats = idx.search(q, k, want_stats=True)
    o = reflib.OracleIndex(files)
    o.max_check = mc
    ids_o, d_o, st_o = o.search(q, k)
    bad = np.nonzero((ids != ids_o).any(axis=1))[0]
    assert bad.size == 0, "%s MaxCheck=%d: %d/%d queries differ, first %d: gpu %s oracle %s" % (
        tag, mc, bad.size, q.shape[0], bad[0], ids[bad[0]], ids_o[bad[0]])
    assert np.array_equal(dists.view(np.int32), d_o.view(np.int32)), (tag, mc)
    for a, b in [(capi.ST_CHECKED, reflib.ST_CHECKED), (capi.ST_TREE_CHECKED, reflib.ST_TREE_CHECKED),
                 (capi.ST_NG_LEFT, reflib.ST_NG_LEFT),
                 (capi.ST_SPT_LEFT, reflib.ST_SPT_LEFT), (capi.ST_NDIST, reflib.ST_NDIST),
                 (capi.ST_NEXPAND, reflib.ST_NEXPAND), (capi.ST_NTREE, reflib.ST_NTREE)]:
        assert np.array_equal(stats[:, a], st_o[:, b]), (tag, mc, "stat", a)
    assert not stats[:, capi.ST_FLAGS].any()


@pytest.mark.parametrize("name", BKT_SETS)
def test_bkt_search_bit_exact(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    k = 3 if name == "algo_line_bkt" else 10
    idx = B200Index.load(folder)
    try:
        for mc in [8192, 1024, 64]:
            _compare(idx, files, q, k, mc, name)
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["kdt_l2_10k_64", "kdt2_l2_6k_32"])
def test_kdt_search_bit_exact(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        for mc in [8192, 1024, 64]:
            _compare(idx, files, q, 10, mc, name)
        idx.set_param("B200.NGCacheEntries", 8)  # KDT keeps ~MaxCheck entries in NGQueue: force the HBM spill
        _compare(idx, files, q, 10, 2048, name + " spill")
    finally:
        idx.close()


def test_algo_line_known_answer_on_gpu():
    # Test/src/AlgoTest.cpp:163-201
    from sptag_b200 import B200Index
    folder = data_folder("algo_line_bkt")
    idx = B200Index.load(folder)
    q = np.array([[0] * 10, [2] * 10, [4] * 10], np.float32)
    ids, dists = idx.search(q, 3)
    assert [set(r) for r in ids.tolist()] == [{0, 1, 2}, {2, 1, 3}, {4, 3, 5}]
    assert dists.tolist() == [[0, 10, 40], [0, 10, 10], [0, 10, 10]]
    idx.close()


@pytest.mark.parametrize("knobs", [
    {"B200.StageRows": 2, "B200.Stages": 1},
    {"B200.StageRows": 6, "B200.Stages": 3},
    {"B200.NGCacheEntries": 7, "B200.SPTCacheEntries": 5},       # nearly everything spills to HBM
    {"B200.NGCacheEntries": 64, "B200.SPTCacheEntries": 4096},
    {"B200.QueriesPerSM": 1},
    {"B200.QueriesPerSM": 16},
    {"B200.VisitedLog": 1},                                     # log + selective clear of the visited bitmap
    {"B200.VisitedLog": 1, "B200.VisitedLogEntries": 300},      # log overflow -> full clear
])
def test_tuning_knobs_do_not_change_results(knobs):
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_20k_32")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:200]
    idx = B200Index.load(folder)
    for name, v in knobs.items():
        idx.set_param(name, v)
    try:
        for mc in [2048, 128]:
            _compare(idx, files, q, 10, mc, str(knobs))
    finally:
        idx.close()


@pytest.mark.parametrize("knobs", [
    {},                                                  # default: one-stage ring, slot count from the batch size
    {"B200.QueriesPerSM": 17},                           # the register-capped (96 registers) 20-slot instantiation
    {"B200.QueriesPerSM": 20},
    {"B200.QueriesPerSM": 16, "B200.Stages": 2},         # two-stage ring
    {"B200.QueriesPerSM": 6, "B200.Stages": 2, "B200.NGCacheEntries": 64, "B200.SPTCacheEntries": 32},
    {"B200.QueriesPerSM": 20, "B200.NGCacheEntries": 32, "B200.SPTCacheEntries": 16},  # queues mostly in the HBM arenas
])
@pytest.mark.parametrize("name", ["bkt_l2_10k_128", "bkt_cos_10k_128", "kdt_l2_8k_128"])
def test_512_byte_row_variants(name, knobs):
    """128-d float rows run the fixed-shape fast path; its ring depth and its register-capped high-residency
    instantiation are chosen per launch (sptag_b200.cu configure) -- every variant must return the reference's bits."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:300]
    idx = B200Index.load(folder)
    for pname, v in knobs.items():
        idx.set_param(pname, v)
    try:
        for mc in [8192, 512]:
            _compare(idx, files, q, 10, mc, "%s %s" % (name, knobs))
    finally:
        idx.close()


@pytest.mark.parametrize("k", [1, 5, 32, 33, 100, 500])
def test_result_counts(k):
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_10k_128")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:100]
    idx = B200Index.load(folder)
    try:
        _compare(idx, files, q, k, 1024, "k=%d" % k)
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["bkt_l2_10k_128", "kdt_l2_10k_64", "bkt_l2_dups", "bkt_i8_l2_5k_100"])
def test_large_k_result_set(name):
    """K > 32 keeps the result set unordered in HBM (append, then overwrite-the-worst + rescan) and sorts it with a
    bitonic network: small K at a large budget exercises the overwrite path on nearly every accepted point, K = 1024
    the padded sort; both must reproduce QueryResultSet's heap (QueryResultSet.h:77-120) exactly."""
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:100]
    idx = B200Index.load(folder)
    try:
        for k, mc in [(40, 8192), (64, 2048), (333, 8192), (1024, 8192)]:
            _compare(idx, files, q, k, mc, "%s k=%d" % (name, k))
    finally:
        idx.close()


def test_full_heap_replacement_path():
    # tiny MaxCheck AND tiny MaxCheckForRefineGraph make Heap::insert hit its count == length branch (Heap.h:43-49)
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_20k_32")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:200]
    idx = B200Index.load(folder)
    try:
        for mc in [16, 48]:
            idx.set_param("MaxCheck", mc)
            idx.set_param("MaxCheckForRefineGraph", mc)
            ids, dists, stats = idx.search(q, 10, want_stats=True)
            o = reflib.OracleIndex(files)
            o.max_check = mc
            o.max_check_refine = mc
            ids_o, d_o, st_o = o.search(q, 10)
            assert np.array_equal(ids, ids_o)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
            assert np.array_equal(stats[:, 3], st_o[:, reflib.ST_SPT_LEFT])
    finally:
        idx.close()


def test_deleted_vectors_are_skipped():
    from sptag_b200 import B200Index, capi
    folder = data_folder("bkt_l2_10k_128")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:100]
    rng = np.random.default_rng(3)
    deleted = (rng.random(files.n) < 0.3).astype(np.int8)
    files.deleted = deleted
    files.num_deleted = int(deleted.sum())
    idx = B200Index.create(algo=capi.ALGO_BKT, value_type=capi.VT_FLOAT, metric=files.metric, vectors=files.vectors,
                           graph=files.graph, tree_starts=files.tree_starts, tree_nodes=files.nodes[:files.node_count],
                           deleted=deleted, num_deleted=files.num_deleted)
    try:
        idx.set_param("MaxCheck", 2048)
        ids, dists = idx.search(q, 10)
        o = reflib.OracleIndex(files)
        o.max_check = 2048
        ids_o, d_o, _ = o.search(q, 10)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
        assert not deleted[ids[ids >= 0]].any()
    finally:
        idx.close()


def test_id_offset_for_shards():
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_3k_30")
    q = np.load(os.path.join(folder, "queries.npy"))[:50]
    a = B200Index.load(folder)
    b = B200Index.load(folder, id_offset=1000000)
    ia, da = a.search(q, 10)
    ib, db = b.search(q, 10)
    assert np.array_equal(np.where(ia >= 0, ia + 1000000, ia), ib)
    assert np.array_equal(da, db)
    a.close()
    b.close()


@pytest.mark.parametrize("metric", [0, 1])
def test_distance_kernel_bit_exact_all_dims(metric):
    """The inner loop alone (sptag_b200_distance_batch) vs the oracle's AVX-512 tree for many dims incl. tails."""
    from sptag_b200 import B200Index, capi
    rng = np.random.default_rng(11)
    L = reflib.ora()
    for dim in [1, 3, 4, 7, 8, 12, 15, 16, 17, 24, 28, 31, 32, 33, 64, 100, 127, 128, 131, 200, 384, 768, 960, 1000]:
        n, nq, per = 257, 9, 33
        x = rng.standard_normal((n, dim), dtype=np.float32)
        q = rng.standard_normal((nq, dim), dtype=np.float32)
        graph = np.full((n, 4), -1, np.int32)
        nodes = np.array([[n, 1, 2], [0, -1, -1], [-1, -1, -1]], np.int32)
        idx = B200Index.create(algo=capi.ALGO_BKT, value_type=capi.VT_FLOAT, metric=metric, vectors=x, graph=graph,
                               tree_starts=np.array([0], np.int32), tree_nodes=nodes)
        ids = rng.integers(0, n, (nq, per)).astype(np.int32)
        ids[0, 0] = -1
        out = idx.distance_batch(q, ids)
        exp = np.empty_like(out)
        for i in range(nq):
            a = np.ascontiguousarray(np.repeat(q[i:i + 1], per, 0))
            b = np.ascontiguousarray(x[np.maximum(ids[i], 0)])
            L.ora_distance_f32_many(metric, 16, a.ctypes.data, b.ctypes.data, dim, per, exp[i].ctypes.data)
        exp[0, 0] = L.ora_max_dist()
        assert np.array_equal(out.view(np.int32), exp.view(np.int32)), dim
        idx.close()


def test_merge_topk_matches_comparator():
    import torch
    from sptag_b200 import capi
    rng = np.random.default_rng(5)
    G, nq, k = 4, 300, 10
    d = np.sort(rng.integers(0, 50, (G, nq, k)).astype(np.float32), axis=2)  # many ties
    ids = rng.permutation(G * nq * k).astype(np.int32).reshape(G, nq, k)
    # sort each list by (dist, id)
    for g in range(G):
        for i in range(nq):
            order = np.lexsort((ids[g, i], d[g, i]))
            ids[g, i], d[g, i] = ids[g, i][order], d[g, i][order]
    ids[1, :, 7:] = -1
    d[1, :, 7:] = reflib.ora().ora_max_dist()
    t_ids = torch.from_numpy(ids).cuda()
    t_d = torch.from_numpy(d).cuda()
    o_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    capi.merge_topk(0, t_ids.data_ptr(), t_d.data_ptr(), G, nq, k, o_ids.data_ptr(), o_d.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for i in range(nq):
        cand = [(d[g, i, j], ids[g, i, j]) for g in range(G) for j in range(k) if ids[g, i, j] >= 0]
        cand.sort()
        exp = cand[:k]
        assert o_ids[i].tolist() == [c[1] for c in exp]
        assert o_d[i].tolist() == [c[0] for c in exp]


# ---------------------------------------------------------------------------------------------
# PQ / OPQ quantized indexes: raw queries in, device-side QuantizeVector,
# SDC table distances -- all bit-exact against the oracle (itself pinned to the reference)
# ---------------------------------------------------------------------------------------------
QUANT_SETS = ["bkt_pq_6k_32", "bkt_opq_6k_48", "bkt_opq_i8_8k_100"]


@pytest.mark.parametrize("name", QUANT_SETS)
def test_quantize_vector_bit_exact(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)          # picks up [Quantizer] QuantizerFilePath
    try:
        codes = idx.quantize(q, files.quantizer.m)
        exp = reflib.OracleQuantizer(files.quantizer).encode(q)
        assert np.array_equal(codes, exp)
    finally:
        idx.close()


@pytest.mark.parametrize("name", QUANT_SETS)
def test_quantized_bkt_search_bit_exact(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        for mc in [8192, 1024, 128]:
            _compare(idx, files, q, 10, mc, name)
    finally:
        idx.close()


# ---------------------------------------------------------------------------------------------
# the committed golden fixtures (tests/golden/*.npz): reference-built indexes + the REFERENCE'S OWN outputs.
# These need nothing but the repository (no tests/_data, no oracle/_ref), so they pin the device path to the
# reference even on a box that only has a fresh checkout.
# ---------------------------------------------------------------------------------------------
def _golden_names():
    import glob
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(g))


@pytest.mark.parametrize("name", _golden_names())
def test_gpu_matches_reference_golden_outputs(name):
    from sptag_b200 import B200Index, capi
    import test_oracle_pin as pin
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    files = reflib.IndexFiles.__new__(reflib.IndexFiles)
    pin._files_from_npz(files, g)
    idx = B200Index.create(algo=capi.ALGO_KDT if files.algo == "KDT" else capi.ALGO_BKT, value_type=capi.VT_FLOAT,
                           metric=files.metric, vectors=files.vectors, graph=files.graph,
                           tree_starts=files.tree_starts, tree_nodes=files.nodes)
    try:
        q = np.ascontiguousarray(g["queries"])
        k = int(g["k"])
        for i, mc in enumerate(g["max_checks"].tolist()):
            idx.set_param("MaxCheck", int(mc))
            ids, dists, stats = idx.search(q, k, want_stats=True)
            assert np.array_equal(ids, g["ref_ids"][i]), (name, mc)
            assert np.array_equal(dists.view(np.int32), g["ref_dists"][i].view(np.int32)), (name, mc)
            # the reference's WorkSpace counters: checked leaves, tree-checked leaves, queue sizes at exit
            assert np.array_equal(stats[:, capi.ST_CHECKED], g["ref_stats"][i][:, 0]), (name, mc)
            assert np.array_equal(stats[:, capi.ST_NG_LEFT], g["ref_stats"][i][:, 2]), (name, mc)
            assert np.array_equal(stats[:, capi.ST_SPT_LEFT], g["ref_stats"][i][:, 3]), (name, mc)
    finally:
        idx.close()


# ---------------------------------------------------------------------------------------------
# int8 / uint8 element types (DistanceUtils integer variants)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["bkt_i8_cos_6k_64", "bkt_u8_l2_6k_128", "bkt_i8_l2_5k_100", "kdt_i8_l2_6k_32",
                                  "bkt_i16_l2_5k_64", "bkt_i16_cos_5k_40", "bkt_i16_l2_4k_27", "kdt_i16_l2_5k_32"])
def test_integer_index_search_bit_exact(name):
    from sptag_b200 import B200Index
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        for mc in [8192, 1024, 128]:
            _compare(idx, files, q, 10, mc, name)
    finally:
        idx.close()


@pytest.mark.parametrize("vt,dt,lo,hi", [(0, np.int8, -127, 128), (1, np.uint8, 0, 256), (2, np.int16, -32768, 32768),
                                         (2, np.int16, -2000, 2000)])
@pytest.mark.parametrize("metric", [0, 1])
def test_integer_distance_kernel_bit_exact_all_dims(vt, dt, lo, hi, metric):
    from sptag_b200 import B200Index, capi
    rng = np.random.default_rng(13)
    L = reflib.ora()
    for dim in [1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 24, 27, 31, 32, 33, 40, 48, 63, 64, 65, 80, 96, 100, 127, 128, 131,
                200, 256, 258]:
        n, nq, per = 129, 7, 21
        x = rng.integers(lo, hi, (n, dim)).astype(dt)
        q = rng.integers(lo, hi, (nq, dim)).astype(dt)
        graph = np.full((n, 4), -1, np.int32)
        nodes = np.array([[n, 1, 2], [0, -1, -1], [-1, -1, -1]], np.int32)
        idx = B200Index.create(algo=capi.ALGO_BKT, value_type=vt, metric=metric, vectors=x, graph=graph,
                               tree_starts=np.array([0], np.int32), tree_nodes=nodes)
        ids = rng.integers(0, n, (nq, per)).astype(np.int32)
        qc = np.ascontiguousarray(q)
        out = np.empty(ids.shape, np.float32)
        capi._check(capi.lib().sptag_b200_distance_batch(idx._h, qc.ctypes.data, nq, ids.ctypes.data, per, out.ctypes.data))
        exp = np.empty_like(out)
        for i in range(nq):
            for jj in range(per):
                a = np.ascontiguousarray(q[i])
                b = np.ascontiguousarray(x[ids[i, jj]])
                exp[i, jj] = L.ora_distance(metric, vt, 16, a.ctypes.data, b.ctypes.data, dim)
        assert np.array_equal(out.view(np.int32), exp.view(np.int32)), (vt, metric, dim)
        idx.close()


# ---------------------------------------------------------------------------------------------
# size-independent properties of the result lists (what the full-size bench run also relies on)
# ---------------------------------------------------------------------------------------------
def test_result_list_properties():
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_20k_32")
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        idx.set_param("MaxCheck", 2048)
        ids, dists = idx.search(q, 10)
        # ascending by (dist, id), ids unique and in range, no unfilled slot on a 20k index
        assert (np.diff(dists, axis=1) >= 0).all()
        ties = np.diff(dists, axis=1) == 0
        assert (np.diff(ids, axis=1)[ties] > 0).all()
        assert (ids >= 0).all() and (ids < idx.num_vectors).all()
        assert all(len(set(r)) == len(r) for r in ids.tolist())
        # reported distance == the inner loop evaluated on (query, id): the list is self-consistent
        assert np.array_equal(idx.distance_batch(q, ids).view(np.int32), dists.view(np.int32))
        # a query's result does not depend on its position in the batch or on its neighbours in the batch
        perm = np.random.default_rng(1).permutation(q.shape[0])
        ids_p, dists_p = idx.search(q[perm], 10)
        assert np.array_equal(ids_p, ids[perm]) and np.array_equal(dists_p, dists[perm])
        rep = np.repeat(q[:7], 5, axis=0)
        ids_r, _ = idx.search(rep, 10)
        assert np.array_equal(ids_r, np.repeat(ids[:7], 5, axis=0))
        # idempotence: the same batch again gives the same bits
        ids2, dists2 = idx.search(q, 10)
        assert np.array_equal(ids2, ids) and np.array_equal(dists2.view(np.int32), dists.view(np.int32))
        # a larger budget never returns a worse k-th distance on this data
        idx.set_param("MaxCheck", 8192)
        _, dists_big = idx.search(q, 10)
        assert (dists_big[:, -1] <= dists[:, -1]).mean() > 0.99
    finally:
        idx.close()


def test_empty_and_single_query_batches():
    from sptag_b200 import B200Index, capi
    folder = data_folder("bkt_l2_3k_30")
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        assert capi.lib().sptag_b200_search(idx._h, None, 0, 10, None, None, None) == 0   # empty batch is a no-op
        one_ids, one_d = idx.search(q[:1], 10)
        all_ids, all_d = idx.search(q, 10)
        assert np.array_equal(one_ids[0], all_ids[0]) and np.array_equal(one_d[0], all_d[0])
        with pytest.raises(capi.SptagB200Error):
            idx.search(q, 2049)          # K > 2048 is rejected loudly, not truncated
        with pytest.raises(capi.SptagB200Error):
            idx.set_param("NoSuchParameter", 1)
    finally:
        idx.close()


@pytest.mark.parametrize("name", QUANT_SETS)
def test_quantized_adc_search_bit_exact(name):
    # VectorIndex::SetQuantizerADC(true): per-query asymmetric distance tables instead of the SDC table
    from sptag_b200 import B200Index, capi
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    try:
        idx.set_param("EnableADC", 1)
        for mc in [8192, 256]:
            idx.set_param("MaxCheck", mc)
            ids, dists, stats = idx.search(q, 10, want_stats=True)
            o = reflib.OracleIndex(files)
            o.max_check = mc
            o.enable_adc = True
            ids_o, d_o, st_o = o.search(q, 10)
            assert np.array_equal(ids, ids_o), (name, mc)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32)), (name, mc)
            assert np.array_equal(stats[:, capi.ST_NDIST], st_o[:, reflib.ST_NDIST])
        idx.set_param("EnableADC", 0)       # and back to SDC on the same handle
        _compare(idx, files, q, 10, 1024, name + " sdc-after-adc")
    finally:
        idx.close()


# ---------------------------------------------------------------------------------------------
# SearchIndexWithFilter on the device
# ---------------------------------------------------------------------------------------------
def test_filter_known_answer_on_gpu():
    from sptag_b200 import B200Index
    folder = data_folder("algo_line_bkt")
    idx = B200Index.load(folder)
    allowed = np.ones(idx.num_vectors, np.uint8)
    allowed[2] = 0                                   # FilterTest.cpp:40-60: metadata "2" is rejected
    ids, _ = idx.search_filtered(np.array([[0] * 10, [2] * 10, [4] * 10], np.float32), 3, allowed)
    assert ids.tolist() == [[0, 1, 3], [1, 3, 0], [4, 3, 5]]
    idx.close()


@pytest.mark.parametrize("name", ["bkt_l2_20k_32", "bkt_l2_dups", "bkt_cos_10k_128"])
def test_filtered_search_bit_exact(name):
    from sptag_b200 import B200Index, capi
    folder = data_folder(name)
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:150]
    allowed = (np.random.default_rng(5).random(files.n) < 0.5).astype(np.uint8)
    idx = B200Index.load(folder)
    try:
        for mc in [0, 512]:
            ids, dists, stats = idx.search_filtered(q, 10, allowed, max_check=mc, want_stats=True)
            o = reflib.OracleIndex(files)
            o.filter = allowed
            if mc:
                o.max_check = mc
            ids_o, d_o, st_o = o.search(q, 10)
            assert np.array_equal(ids, ids_o), (name, mc)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32)), (name, mc)
            assert np.array_equal(stats[:, capi.ST_CHECKED], st_o[:, reflib.ST_CHECKED])
        _compare(idx, files, q, 10, 1024, name + " unfiltered-after-filtered")   # the override does not stick
    finally:
        idx.close()


def test_filter_rejected_on_kdt():
    from sptag_b200 import B200Index, capi
    folder = data_folder("kdt_l2_10k_64")
    idx = B200Index.load(folder)
    q = np.load(os.path.join(folder, "queries.npy"))[:4]
    with pytest.raises(capi.SptagB200Error):     # "Not Support Filter on KDT Index!" (KDTIndex.cpp:361-365)
        idx.search_filtered(q, 10, np.ones(idx.num_vectors, np.uint8))
    idx.close()


def test_concurrent_callers_on_one_handle():
    """SearchIndex is const and re-entrant in the reference (OpenMP callers, thread pools); the C ABI serialises
    callers of one handle on its mutex.  Four host threads with different batches, K and budgets must each get
    exactly what a lone caller gets."""
    import threading
    from sptag_b200 import B200Index
    folder = data_folder("bkt_l2_10k_128")
    files = reflib.IndexFiles(folder)
    q = np.load(os.path.join(folder, "queries.npy"))
    idx = B200Index.load(folder)
    idx.set_param("MaxCheck", 1024)
    o = reflib.OracleIndex(files)
    o.max_check = 1024
    jobs = [(q[0:80], 10), (q[80:150], 5), (q[150:260], 40), (q[260:300], 1)]
    expect = [o.search(b, k)[:2] for b, k in jobs]
    out = [None] * len(jobs)
    errs = []

    def run(i):
        try:
            for _ in range(5):
                out[i] = idx.search(jobs[i][0], jobs[i][1])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    try:
        assert not errs, errs
        for (ids, dists), (ids_o, d_o) in zip(out, expect):
            assert np.array_equal(ids, ids_o)
            assert np.array_equal(dists.view(np.int32), d_o.view(np.int32))
    finally:
        idx.close()
