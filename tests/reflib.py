"""Test-side helpers (TEST INFRASTRUCTURE): ctypes bindings for

* ``oracle/_ref/libsptag_ref.so``   -- the UNMODIFIED reference compiled by ``oracle/Makefile ref``
* ``oracle/_build/libsptag_oracle.so`` -- our plain-C restatement (``oracle/sptag_oracle.c``)

plus a numpy reader for the reference's on-disk index folder (formats: SURVEY.md section 8 f1;
``Dataset.h:146-180``, ``NeighborhoodGraph.h:606-615``, ``BKTree.h:635-645``, ``KDTree.h:123-133``,
``Labelset.h:78-83``, ``VectorIndex.cpp:197-222``).  Nothing here is imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsptag_ref.so")
ORA_SO = os.path.join(ROOT, "oracle", "_build", "libsptag_oracle.so")
DATA_DIR = os.path.join(ROOT, "tests", "_data")

VT_INT8, VT_UINT8, VT_INT16, VT_FLOAT = 0, 1, 2, 3
NP_OF_VT = {VT_INT8: np.int8, VT_UINT8: np.uint8, VT_INT16: np.int16, VT_FLOAT: np.float32}
VT_OF_NAME = {"Int8": VT_INT8, "UInt8": VT_UINT8, "Int16": VT_INT16, "Float": VT_FLOAT}
METRIC_OF_NAME = {"L2": 0, "Cosine": 1, "InnerProduct": 2}
ALGO_OF_NAME = {"BKT": 0, "KDT": 1}
ORA_ST_COUNT = 8
ST_CHECKED, ST_TREE_CHECKED, ST_NG_LEFT, ST_SPT_LEFT, ST_NDIST, ST_NEXPAND, ST_NTREE = range(7)


def have_ref():
    return os.path.exists(REF_SO)


def build_port():
    """(Re)build the C restatement; cheap (gcc, <2 s)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])


# ------------------------------------------------------------------------------------------------
# reference shim
# ------------------------------------------------------------------------------------------------
_ref = None


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_build.restype = C.c_void_p
        L.ref_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.ref_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.ref_get_param.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_num_samples.argtypes = [C.c_void_p]
        L.ref_dim.argtypes = [C.c_void_p]
        L.ref_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_enable_stats.argtypes = [C.c_void_p]
        L.ref_search_one_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_distance.restype = C.c_float
        L.ref_distance.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_distance_f32_isa.restype = C.c_float
        L.ref_distance_f32_isa.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_distance_f32_many.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_build_quantized.restype = C.c_void_p
        L.ref_build_quantized.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.ref_quantizer_load.restype = C.c_void_p
        L.ref_quantizer_load.argtypes = [C.c_char_p]
        L.ref_quantizer_m.argtypes = [C.c_void_p]
        L.ref_quantizer_reconstruct_dim.argtypes = [C.c_void_p]
        L.ref_quantizer_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_rebuild_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_quantizer_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_quantizer_l2.restype = C.c_float
        L.ref_quantizer_l2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_set_adc.argtypes = [C.c_void_p, C.c_int]
        L.ref_search_filtered.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p]
        L.ref_search_each.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_refine_nodes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_iter_open.restype = C.c_void_p
        L.ref_iter_open.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_iter_next.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_iter_close.argtypes = [C.c_void_p]
        L.ref_iter_scan.restype = C.c_longlong
        L.ref_iter_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_double)]
        L.ref_delete.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_nearest_open.restype = C.c_void_p
        L.ref_nearest_open.argtypes = [C.c_void_p, C.c_int]
        L.ref_nearest_next.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_nearest_close.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_search_each_flag.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p]
        L.ref_iter_open_flag.restype = C.c_void_p
        L.ref_iter_open_flag.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_quiet(3)  # warnings and errors only
        _ref = L
    return _ref


class RefIndex:
    """The reference's VectorIndex behind the shim."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference index handle is null")
        self.h = C.c_void_p(handle)

    @classmethod
    def build(cls, algo, data, metric, threads=8, params=""):
        data = np.ascontiguousarray(data)
        vt = {np.dtype(np.int8): VT_INT8, np.dtype(np.uint8): VT_UINT8,
              np.dtype(np.int16): VT_INT16, np.dtype(np.float32): VT_FLOAT}[data.dtype]
        h = ref().ref_build(ALGO_OF_NAME[algo], vt, METRIC_OF_NAME[metric], data.ctypes.data,
                            data.shape[0], data.shape[1], threads, params.encode())
        return cls(h)

    @classmethod
    def load(cls, folder):
        return cls(ref().ref_load(folder.encode()))

    @classmethod
    def load_memory(cls, ini_text, vectors, graph, nodes, tree_starts):
        """VectorIndex::LoadIndex(config, blobs): the blobs are the bytes of vectors.bin / tree.bin / graph.bin."""
        def blob(header, *arrays):
            total = 4 * len(header) + sum(a.nbytes for a in arrays)
            b = np.empty(total, np.uint8)
            b[:4 * len(header)] = np.array(header, np.int32).view(np.uint8)
            at = 4 * len(header)
            for a in arrays:
                b[at:at + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
                at += a.nbytes
            return b
        vb = blob([vectors.shape[0], vectors.shape[1]], vectors)
        gb = blob([graph.shape[0], graph.shape[1]], np.ascontiguousarray(graph, np.int32))
        ts = np.ascontiguousarray(tree_starts, np.int32)
        tb = np.concatenate([np.array([ts.shape[0]], np.int32), ts, np.array([nodes.shape[0]], np.int32),
                             np.ascontiguousarray(nodes, np.int32).reshape(-1)]).view(np.uint8)
        L = ref()
        L.ref_load_memory.restype = C.c_void_p
        L.ref_load_memory.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        h = L.ref_load_memory(ini_text.encode(), vb.ctypes.data, vb.nbytes, tb.ctypes.data, tb.nbytes, gb.ctypes.data,
                              gb.nbytes)
        idx = cls(h)
        idx._blobs = (vb, tb, gb)  # Dataset::Load(char*) keeps pointing into the blobs (Dataset.h:191-204)
        return idx

    def save(self, folder):
        os.makedirs(folder, exist_ok=True)
        rc = ref().ref_save(self.h, folder.encode())
        if rc != 0:
            raise RuntimeError("SaveIndex failed: %d" % rc)

    def set_param(self, name, value):
        return ref().ref_set_param(self.h, name.encode(), str(value).encode())

    def search(self, queries, k, threads=0):
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        sec = C.c_double()
        rc = ref().ref_search_batch(self.h, queries.ctypes.data, nq, k, threads, ids.ctypes.data,
                                    dists.ctypes.data, C.byref(sec))
        if rc != 0:
            raise RuntimeError("reference SearchIndex failed: %d" % rc)
        return ids, dists, sec.value

    @classmethod
    def build_quantized(cls, algo, codes, metric, quantizer_file, threads=8, params=""):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        h = ref().ref_build_quantized(ALGO_OF_NAME[algo], METRIC_OF_NAME[metric], quantizer_file.encode(),
                                      codes.ctypes.data, codes.shape[0], codes.shape[1], threads, params.encode())
        return cls(h)

    def search_filtered(self, queries, k, allowed, max_check=0, threads=0):
        """VectorIndex::SearchIndexWithFilter with allowed[id] (uint8) as the predicate."""
        queries = np.ascontiguousarray(queries)
        allowed = np.ascontiguousarray(allowed, np.uint8)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        bad = ref().ref_search_filtered(self.h, queries.ctypes.data, nq, queries.strides[0], k, allowed.ctypes.data,
                                        max_check, threads, ids.ctypes.data, dists.ctypes.data)
        if bad:
            raise RuntimeError("SearchIndexWithFilter failed for %d queries" % bad)
        return ids, dists

    def set_adc(self, enable):
        ref().ref_set_adc(self.h, 1 if enable else 0)

    def search_each(self, queries, k, threads=0):
        """Per-query overload on RAW queries (needed for quantized indexes)."""
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        sec = C.c_double()
        ref().ref_search_each(self.h, queries.ctypes.data, nq, queries.strides[0], k, threads, ids.ctypes.data,
                              dists.ctypes.data, C.byref(sec))
        return ids, dists, sec.value

    def refine_nodes(self, first, num, cef, neighborhood=32, rng_factor=1.0, threads=0):
        """NeighborhoodGraph::RefineNode per node against the loaded graph (graph not modified):
        RefineSearchIndex + RelativeNeighborhoodGraph::RebuildNeighbors.  -> (rows, result ids, result dists)."""
        rows = np.empty((num, neighborhood), np.int32)
        ids = np.empty((num, cef + 1), np.int32)
        dists = np.empty((num, cef + 1), np.float32)
        rc = ref().ref_refine_nodes(self.h, first, num, cef, neighborhood, rng_factor, threads, rows.ctypes.data,
                                    ids.ctypes.data, dists.ctypes.data)
        if rc != 0:
            raise RuntimeError("reference refine failed: %d" % rc)
        return rows, ids, dists

    def rebuild_graph(self, graph, neighborhood):
        """NeighborhoodGraph::RebuildGraph run by the reference (one thread) on `graph` [n, stride >= 2*neighborhood];
        returns the rows afterwards (the first `neighborhood` entries of a row are its neighbours)."""
        g = np.ascontiguousarray(graph, np.int32).copy()
        rc = ref().ref_rebuild_graph(self.h, g.ctypes.data, g.shape[0], g.shape[1], neighborhood)
        if rc != 0:
            raise RuntimeError("ref_rebuild_graph failed")
        return g

    def search_flag(self, queries, k, search_deleted, threads=0):
        """SearchIndex(QueryResult&, p_searchDeleted) per query."""
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        ref().ref_search_each_flag(self.h, queries.ctypes.data, nq, queries.strides[0], k, threads,
                                   1 if search_deleted else 0, ids.ctypes.data, dists.ctypes.data)
        return ids, dists

    def delete(self, ids):
        """VectorIndex::DeleteIndex(id) for every id (tombstones)."""
        ids = np.ascontiguousarray(ids, np.int32)
        bad = ref().ref_delete(self.h, ids.ctypes.data, ids.shape[0])
        if bad:
            raise RuntimeError("DeleteIndex failed for %d ids" % bad)

    def iterator_scan(self, queries, batch, rounds, threads=0):
        """All-cores CPU baseline: one ResultIterator per query, rounds x Next(batch) -> (results, seconds)."""
        queries = np.ascontiguousarray(queries)
        sec = C.c_double()
        n = ref().ref_iter_scan(self.h, queries.ctypes.data, queries.shape[0], queries.strides[0], batch, rounds, threads,
                                C.byref(sec))
        return int(n), sec.value

    def iterator(self, query, search_deleted=False):
        """VectorIndex::GetIterator: the reference's own ResultIterator for one query."""
        return _RefIterator(self, np.ascontiguousarray(query), search_deleted)

    def enable_stats(self):
        return ref().ref_enable_stats(self.h)

    def search_one_stats(self, query, k):
        query = np.ascontiguousarray(query)
        ids = np.empty(k, np.int32)
        dists = np.empty(k, np.float32)
        stats = np.zeros(4, np.int32)
        ref().ref_search_one_stats(self.h, query.ctypes.data, k, ids.ctypes.data, dists.ctypes.data,
                                   stats.ctypes.data)
        return ids, dists, stats


class RefNearestScan:
    """SearchIndexIterativeFromNeareast driven like SPANN drives its head index: k results per call."""

    def __init__(self, index, query, k):
        self.index, self.query, self.k = index, np.ascontiguousarray(query), k
        self.ws = ref().ref_nearest_open(index.h, k)
        self.first = True

    def next(self):
        ids = np.empty(self.k, np.int32)
        dists = np.empty(self.k, np.float32)
        ok = ref().ref_nearest_next(self.index.h, self.ws, self.query.ctypes.data, self.k, 1 if self.first else 0,
                                    ids.ctypes.data, dists.ctypes.data)
        self.first = False
        return bool(ok), ids, dists

    def close(self):
        if self.ws:
            ref().ref_nearest_close(self.index.h, self.ws)
            self.ws = None


class _RefIterator:
    def __init__(self, index, query, search_deleted=False):
        self.query = query            # the iterator borrows the target buffer
        self.index = index
        self.h = ref().ref_iter_open_flag(index.h, query.ctypes.data, 1 if search_deleted else 0)
        if not self.h:
            raise RuntimeError("GetIterator returned null")

    def next(self, batch):
        ids = np.empty(batch, np.int32)
        dists = np.empty(batch, np.float32)
        relaxed = C.c_int()
        count = ref().ref_iter_next(self.h, batch, ids.ctypes.data, dists.ctypes.data, C.byref(relaxed))
        return count, ids, dists, bool(relaxed.value)

    def close(self):
        if self.h:
            ref().ref_iter_close(self.h)
            self.h = None


# ------------------------------------------------------------------------------------------------
# on-disk index folder -> numpy
# ------------------------------------------------------------------------------------------------
class IndexFiles:
    """Flat numpy view of a reference index folder."""

    def __init__(self, folder):
        self.folder = folder
        self.params = {}
        with open(os.path.join(folder, "indexloader.ini")) as f:
            for line in f:
                line = line.strip()
                if "=" in line and not line.startswith("["):
                    k, v = line.split("=", 1)
                    self.params[k] = v
        p = self.params
        self.algo = p["IndexAlgoType"]
        self.value_type = VT_OF_NAME[p["ValueType"]]
        self.metric = METRIC_OF_NAME[p.get("DistCalcMethod", "Cosine")]
        dt = NP_OF_VT[self.value_type]

        raw = np.fromfile(os.path.join(folder, p.get("VectorFilePath", "vectors.bin")), dtype=np.uint8)
        self.n, self.dim = (int(v) for v in raw[:8].view(np.int32))
        self.vectors = raw[8:8 + self.n * self.dim * np.dtype(dt).itemsize].view(dt).reshape(self.n, self.dim)

        raw = np.fromfile(os.path.join(folder, p.get("GraphFilePath", "graph.bin")), dtype=np.int32)
        gn, self.degree = int(raw[0]), int(raw[1])
        assert gn == self.n
        self.graph = raw[2:2 + gn * self.degree].reshape(gn, self.degree)

        raw = np.fromfile(os.path.join(folder, p.get("TreeFilePath", "tree.bin")), dtype=np.int32)
        self.tree_num = int(raw[0])
        self.tree_starts = np.ascontiguousarray(raw[1:1 + self.tree_num])
        self.node_count = int(raw[1 + self.tree_num])
        body = raw[2 + self.tree_num:]
        if self.algo == "BKT":
            nodes = body[:self.node_count * 3].reshape(self.node_count, 3)
            # LoadTrees appends a (-1,-1,-1) sentinel if the last node is not one (BKTree.h:662)
            if self.node_count > 0 and nodes[-1, 0] != -1:
                nodes = np.vstack([nodes, np.array([[-1, -1, -1]], np.int32)])
            self.nodes = np.ascontiguousarray(nodes)
        else:
            self.nodes = np.ascontiguousarray(body[:self.node_count * 4].reshape(self.node_count, 4))

        dpath = os.path.join(folder, p.get("DeleteVectorFilePath", "deletes.bin"))
        self.num_deleted = 0
        self.deleted = None
        if os.path.exists(dpath):
            raw = np.fromfile(dpath, dtype=np.uint8)
            self.num_deleted = int(raw[:4].view(np.int32)[0])
            rows = int(raw[4:8].view(np.int32)[0])
            self.deleted = np.ascontiguousarray(raw[12:12 + rows].view(np.int8))

        self.quantizer = None
        qf = p.get("QuantizerFilePath")
        if qf and os.path.exists(os.path.join(folder, qf)):
            self.quantizer = Quantizer.read(os.path.join(folder, qf))

    def int_param(self, name, default):
        return int(self.params.get(name, default))


# ------------------------------------------------------------------------------------------------
# quantizer.bin (PQQuantizer::SaveQuantizer, PQQuantizer.h:226-239; OPQQuantizer::SaveQuantizer,
# OPQQuantizer.h:133-147): uint8 qtype, uint8 rtype, int32 M, int32 Ks, int32 dimPerSub, codebooks, [rotation]
# ------------------------------------------------------------------------------------------------
Q_NONE, Q_PQ, Q_OPQ = 0, 1, 2


class Quantizer:
    def __init__(self, qtype, rtype, codebooks, rotation=None):
        self.qtype, self.rtype = qtype, rtype
        self.codebooks = np.ascontiguousarray(codebooks, np.float32)   # [M, Ks, dsub]
        self.m, self.ks, self.dsub = self.codebooks.shape
        self.rotation = None if rotation is None else np.ascontiguousarray(rotation, np.float32)
        self.dim = self.m * self.dsub

    def write(self, path):
        with open(path, "wb") as f:
            np.array([self.qtype, self.rtype], np.uint8).tofile(f)
            np.array([self.m, self.ks, self.dsub], np.int32).tofile(f)
            self.codebooks.tofile(f)
            if self.qtype == Q_OPQ:
                self.rotation.tofile(f)

    @classmethod
    def read(cls, path):
        raw = np.fromfile(path, dtype=np.uint8)
        qtype, rtype = int(raw[0]), int(raw[1])
        m, ks, dsub = (int(v) for v in raw[2:14].view(np.int32))
        if qtype == Q_PQ and rtype != VT_FLOAT:
            raise NotImplementedError("PQQuantizer<%d>: only float codebooks are handled" % rtype)
        n = m * ks * dsub
        cb = raw[14:14 + 4 * n].view(np.float32).reshape(m, ks, dsub)
        rot = None
        if qtype == Q_OPQ:
            d = m * dsub
            rot = raw[14 + 4 * n:14 + 4 * n + 4 * d * d].view(np.float32).reshape(d, d)
        return cls(qtype, rtype, cb.copy(), None if rot is None else rot.copy())

    def blob(self):
        import io
        b = io.BytesIO()
        b.write(np.array([self.qtype, self.rtype], np.uint8).tobytes())
        b.write(np.array([self.m, self.ks, self.dsub], np.int32).tobytes())
        b.write(self.codebooks.tobytes())
        if self.qtype == Q_OPQ:
            b.write(self.rotation.tobytes())
        return b.getvalue()


def train_quantizer(data, m, ks=256, opq=False, rtype=VT_FLOAT, seed=0, iters=6):
    """Synthetic quantizer for tests/bench: per-subspace k-means codebooks (+ a random orthonormal rotation
    for OPQ).  The reference cannot train OPQ natively (Quantizer/main.cpp:157-163), so SURVEY.md section 7
    prescribes exactly this synthesis."""
    rng = np.random.default_rng(seed)
    x = np.asarray(data, np.float32)
    dim = x.shape[1]
    dsub = dim // m
    assert dsub * m == dim
    rot = None
    if opq:
        qm, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
        rot = qm.astype(np.float32)
        x = x @ rot  # vec . R: the reference computes out[i] = dot(vec, R^T row i) = sum_j vec[j] R[j][i]
    sample = x[rng.choice(x.shape[0], min(x.shape[0], 20000), replace=False)]
    cb = np.zeros((m, ks, dsub), np.float32)
    for i in range(m):
        sub = sample[:, i * dsub:(i + 1) * dsub]
        c = sub[rng.choice(sub.shape[0], ks, replace=sub.shape[0] < ks)].copy()
        for _ in range(iters):
            d = ((sub[:, None, :] - c[None, :, :]) ** 2).sum(-1)
            a = d.argmin(1)
            for j in range(ks):
                sel = sub[a == j]
                if sel.shape[0]:
                    c[j] = sel.mean(0)
        cb[i] = c
    return Quantizer(Q_OPQ if opq else Q_PQ, rtype, cb, rot)


class RefQuantizer:
    def __init__(self, path):
        self.h = C.c_void_p(ref().ref_quantizer_load(path.encode()))
        if not self.h:
            raise RuntimeError("reference could not load quantizer " + path)
        self.m = ref().ref_quantizer_m(self.h)
        self.dim = ref().ref_quantizer_reconstruct_dim(self.h)

    def encode(self, raw):
        raw = np.ascontiguousarray(raw)
        out = np.empty((raw.shape[0], self.m), np.uint8)
        ref().ref_quantizer_encode(self.h, raw.ctypes.data, raw.shape[0], out.ctypes.data)
        return out

    def l2(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return ref().ref_quantizer_l2(self.h, a.ctypes.data, b.ctypes.data)

    def reconstruct(self, codes, dtype):
        codes = np.ascontiguousarray(codes, np.uint8)
        out = np.empty((codes.shape[0], self.dim), dtype)
        ref().ref_quantizer_reconstruct(self.h, codes.ctypes.data, codes.shape[0], out.ctypes.data)
        return out


# ------------------------------------------------------------------------------------------------
# C restatement
# ------------------------------------------------------------------------------------------------
class _OraQuantizer(C.Structure):
    _fields_ = [("qtype", C.c_int32), ("rtype", C.c_int32), ("m", C.c_int32), ("ks", C.c_int32),
                ("dsub", C.c_int32), ("simd_width", C.c_int32), ("enable_adc", C.c_int32), ("codebooks", C.c_void_p),
                ("rotation", C.c_void_p), ("sdc", C.c_void_p), ("rotation_t", C.c_void_p)]


class OracleQuantizer:
    """oracle/sptag_oracle.c quantizer over a Quantizer (tables built by ora_quantizer_init)."""

    def __init__(self, quant, simd_width=16):
        self.q = quant
        self.sdc = np.empty((quant.m, quant.ks, quant.ks), np.float32)
        self.rot_t = np.empty((quant.dim, quant.dim), np.float32) if quant.qtype == Q_OPQ else None
        s = _OraQuantizer()
        s.qtype, s.rtype, s.m, s.ks, s.dsub, s.simd_width = quant.qtype, quant.rtype, quant.m, quant.ks, quant.dsub, simd_width
        s.enable_adc = 0
        s.codebooks = quant.codebooks.ctypes.data
        s.rotation = quant.rotation.ctypes.data if quant.rotation is not None else None
        s.sdc = self.sdc.ctypes.data
        s.rotation_t = self.rot_t.ctypes.data if self.rot_t is not None else None
        self.struct = s
        ora().ora_quantizer_init(C.byref(s))

    def encode(self, raw):
        raw = np.ascontiguousarray(raw)
        out = np.empty((raw.shape[0], self.q.m), np.uint8)
        ora().ora_quantizer_encode(C.byref(self.struct), raw.ctypes.data, raw.shape[0], out.ctypes.data)
        return out

    def l2(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return ora().ora_quantizer_l2(C.byref(self.struct), a.ctypes.data, b.ctypes.data)

    def reconstruct(self, codes, dtype):
        codes = np.ascontiguousarray(codes, np.uint8)
        out = np.empty((codes.shape[0], self.q.dim), dtype)
        ora().ora_quantizer_reconstruct(C.byref(self.struct), codes.ctypes.data, codes.shape[0], out.ctypes.data)
        return out


class _OraIndex(C.Structure):
    _fields_ = [("n", C.c_int32), ("dim", C.c_int32), ("value_type", C.c_int32), ("metric", C.c_int32),
                ("vectors", C.c_void_p), ("degree", C.c_int32), ("graph", C.c_void_p),
                ("tree_kind", C.c_int32), ("tree_num", C.c_int32), ("node_count", C.c_int32),
                ("tree_starts", C.c_void_p), ("nodes", C.c_void_p), ("deleted", C.c_void_p),
                ("num_deleted", C.c_int32), ("max_check", C.c_int32), ("max_check_refine", C.c_int32),
                ("initial_pivots", C.c_int32), ("other_pivots", C.c_int32),
                ("no_better_threshold", C.c_int32), ("simd_width", C.c_int32), ("quantizer", C.c_void_p),
                ("filter", C.c_void_p)]


_ora = None


def ora():
    global _ora
    if _ora is None:
        if not os.path.exists(ORA_SO):
            build_port()
        L = C.CDLL(ORA_SO)
        L.ora_max_dist.restype = C.c_float
        L.ora_distance.restype = C.c_float
        L.ora_distance.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_distance_f32_many.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_int32, C.c_void_p]
        L.ora_search_batch.argtypes = [C.POINTER(_OraIndex), C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_refine_nodes.argtypes = [C.POINTER(_OraIndex), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_iter_open.restype = C.c_void_p
        L.ora_iter_open.argtypes = [C.POINTER(_OraIndex), C.c_void_p]
        L.ora_iter_next.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        L.ora_iter_close.argtypes = [C.c_void_p]
        L.ora_iter_next_from_nearest.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.ora_quantizer_init.argtypes = [C.POINTER(_OraQuantizer)]
        L.ora_quantizer_encode.argtypes = [C.POINTER(_OraQuantizer), C.c_void_p, C.c_int32, C.c_void_p]
        L.ora_rebuild_graph.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.ora_quantizer_reconstruct.argtypes = [C.POINTER(_OraQuantizer), C.c_void_p, C.c_int32, C.c_void_p]
        L.ora_quantizer_l2.restype = C.c_float
        L.ora_quantizer_l2.argtypes = [C.POINTER(_OraQuantizer), C.c_void_p, C.c_void_p]
        _ora = L
    return _ora


class OracleIndex:
    """oracle/sptag_oracle.c over an IndexFiles."""

    def __init__(self, files, simd_width=16):
        self.files = files
        self.simd_width = simd_width
        self.max_check = files.int_param("MaxCheck", 8192)
        self.max_check_refine = files.int_param("MaxCheckForRefineGraph", 8192)
        self.initial_pivots = files.int_param("NumberOfInitialDynamicPivots", 50)
        self.other_pivots = files.int_param("NumberOfOtherDynamicPivots", 4)
        self.no_better_threshold = files.int_param("ThresholdOfNumberOfContinuousNoBetterPropagation", 3)
        self.oq = OracleQuantizer(files.quantizer, simd_width) if getattr(files, "quantizer", None) is not None else None
        self.enable_adc = False
        self.filter = None   # numpy uint8 [n]: SearchIndexWithFilter semantics
        self.search_deleted = False   # p_searchDeleted: tombstones are ignored (dispatch flag, BKTIndex.cpp:473)

    def _struct(self):
        f = self.files
        s = _OraIndex()
        s.n, s.dim, s.value_type, s.metric = f.n, f.dim, f.value_type, f.metric
        s.vectors = f.vectors.ctypes.data
        s.degree = f.degree
        s.graph = f.graph.ctypes.data
        s.tree_kind = ALGO_OF_NAME[f.algo]
        s.tree_num = f.tree_num
        s.node_count = f.nodes.shape[0]
        s.tree_starts = f.tree_starts.ctypes.data
        s.nodes = f.nodes.ctypes.data
        use_del = f.deleted is not None and f.num_deleted > 0 and not self.search_deleted
        s.deleted = f.deleted.ctypes.data if use_del else None
        s.num_deleted = f.num_deleted if use_del else 0
        s.max_check = self.max_check
        s.max_check_refine = self.max_check_refine
        s.initial_pivots = self.initial_pivots
        s.other_pivots = self.other_pivots
        s.no_better_threshold = self.no_better_threshold
        s.simd_width = self.simd_width
        if self.oq is not None:
            self.oq.struct.enable_adc = 1 if self.enable_adc else 0
        s.quantizer = C.addressof(self.oq.struct) if self.oq is not None else None
        s.filter = self.filter.ctypes.data if self.filter is not None else None
        return s

    def search(self, queries, k, threads=0, want_stats=True):
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        stats = np.zeros((nq, ORA_ST_COUNT), np.int32)
        s = self._struct()
        rc = ora().ora_search_batch(C.byref(s), queries.ctypes.data, nq, k, ids.ctypes.data,
                                    dists.ctypes.data, stats.ctypes.data if want_stats else None, threads)
        assert rc == 0
        return ids, dists, stats

    def iterator(self, query):
        return _OraIterator(self, np.ascontiguousarray(query))

    def refine_nodes(self, first, num, cef, neighborhood=32, rng_factor=1.0, threads=0):
        """One RefineNode pass over [first, first+num) against the current graph (ora_refine_nodes)."""
        rows = np.empty((num, neighborhood), np.int32)
        ids = np.empty((num, cef + 1), np.int32)
        dists = np.empty((num, cef + 1), np.float32)
        s = self._struct()
        rc = ora().ora_refine_nodes(C.byref(s), first, num, cef, neighborhood, rng_factor, rows.ctypes.data,
                                    ids.ctypes.data, dists.ctypes.data, threads)
        assert rc == 0
        return rows, ids, dists


class _OraIterator:
    def __init__(self, oindex, query):
        self.oindex = oindex
        self.struct = oindex._struct()     # keeps the arrays alive through oindex.files
        self.h = ora().ora_iter_open(C.byref(self.struct), query.ctypes.data)
        if not self.h:
            raise RuntimeError("ora_iter_open: unsupported index kind")

    def next(self, batch):
        ids = np.empty(batch, np.int32)
        dists = np.empty(batch, np.float32)
        relaxed = C.c_int32()
        count = ora().ora_iter_next(self.h, batch, ids.ctypes.data, dists.ctypes.data, C.byref(relaxed))
        return count, ids, dists, bool(relaxed.value)

    def next_from_nearest(self, k):
        """SearchIndexIterativeFromNeareast: -> (ok, ids [k], dists [k])."""
        ids = np.empty(k, np.int32)
        dists = np.empty(k, np.float32)
        ok = ora().ora_iter_next_from_nearest(self.h, k, ids.ctypes.data, dists.ctypes.data)
        assert ok >= 0
        return bool(ok), ids, dists

    def close(self):
        if self.h:
            ora().ora_iter_close(self.h)
            self.h = None


# ------------------------------------------------------------------------------------------------
# synthetic data (BASELINE.md generators, numpy flavour -- parity is always checked on the SAME
# saved index files, so the exact libstdc++ sequence is not needed)
# ------------------------------------------------------------------------------------------------
def oracle_rebuild_graph(graph, neighborhood):
    """oracle/sptag_oracle.c ora_rebuild_graph on a copy of `graph` [n, stride]."""
    g = np.ascontiguousarray(graph, np.int32).copy()
    if ora().ora_rebuild_graph(g.ctypes.data, g.shape[0], g.shape[1], neighborhood) != 0:
        raise RuntimeError("ora_rebuild_graph failed")
    return g


def gen_iid(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def gen_lowrank(n, dim, rank, seed, noise=0.1):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((rank, dim), dtype=np.float32) / np.float32(np.sqrt(rank))
    z = rng.standard_normal((n, rank), dtype=np.float32)
    return (z @ A + np.float32(noise) * rng.standard_normal((n, dim), dtype=np.float32)).astype(np.float32)


def normalize_rows(x):
    x = x.astype(np.float32)
    nrm = np.sqrt((x.astype(np.float64) ** 2).sum(1))
    nrm[nrm == 0] = 1
    return (x / nrm[:, None]).astype(np.float32)
