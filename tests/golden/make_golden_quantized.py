"""Golden fixtures for quantized indexes -- TEST INFRASTRUCTURE.

tests/golden/quantized/<name>.npz holds a tiny PQ / OPQ index BUILT BY THE UNMODIFIED REFERENCE over codes the reference's
own QuantizeVector produced (VectorIndex::BuildIndex with a quantizer, SaveIndex), flattened to arrays, the quantizer file's
bytes, raw queries, and the reference's own outputs on it:
  * per-query SearchIndex on RAW queries (SDC, and with SetQuantizerADC(true)) for two MaxCheck values,
  * IQuantizer::ReconstructVector of the first code rows,
  * one RefineNode step per node on a quantized index (reconstruct -> SetTarget -> RefineSearchIndex -> RebuildNeighbors,
    NeighborhoodGraph.h:535-549): result lists and new rows.
/root/reference is not needed to USE the fixtures.  Run (where oracle/_ref exists): python tests/golden/make_golden_quantized.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib  # noqa: E402


def i8(x):
    return np.clip(np.round(x * 32), -127, 127).astype(np.int8)


# name -> (raw data, raw queries, M, opq, reconstruct type, numpy dtype of raw vectors)
CASES = {
    "opq_i8_1500_24": (lambda: i8(reflib.gen_lowrank(1500, 24, 6, 41)), lambda: i8(reflib.gen_lowrank(48, 24, 6, 42)), 6, True,
                       reflib.VT_INT8),
    "pq_f32_1200_16": (lambda: reflib.gen_lowrank(1200, 16, 5, 43), lambda: reflib.gen_lowrank(48, 16, 5, 44), 4, False,
                       reflib.VT_FLOAT),
}
MAX_CHECKS = [8192, 256]
REFINE = (40, 512, 16, 1.0, 300)  # CEF, MaxCheckForRefineGraph, neighbourhood, RNGFactor, nodes


def make(name):
    gen_data, gen_q, m, opq, rtype = CASES[name]
    raw = np.ascontiguousarray(gen_data())
    q = np.ascontiguousarray(gen_q())
    k = 10
    with tempfile.TemporaryDirectory() as tmp:
        qz = reflib.train_quantizer(raw.astype(np.float32), m=m, ks=256, opq=opq, rtype=rtype, seed=7, iters=3)
        qpath = os.path.join(tmp, "quantizer_src.bin")
        qz.write(qpath)
        rq = reflib.RefQuantizer(qpath)
        codes = rq.encode(raw)
        folder = os.path.join(tmp, "idx")
        os.makedirs(folder)
        reflib.RefIndex.build_quantized("BKT", codes, "L2", qpath, threads=8).save(folder)
        files = reflib.IndexFiles(folder)
        assert files.quantizer is not None and np.array_equal(files.vectors, codes)
        r = reflib.RefIndex.load(folder)
        out = {}
        for adc, tag in ((False, "sdc"), (True, "adc")):
            r.set_adc(adc)
            ids_all, d_all = [], []
            for mc in MAX_CHECKS:
                r.set_param("MaxCheck", mc)
                ids, dists, _ = r.search_each(q, k, threads=4)
                ids_all.append(ids.astype(np.int32))
                d_all.append(dists)
            out["ref_ids_" + tag] = np.stack(ids_all)
            out["ref_dists_" + tag] = np.stack(d_all)
        r.set_adc(False)
        cef, mcr, nbh, factor, num = REFINE
        r.set_param("MaxCheckForRefineGraph", mcr)
        rows, rids, rd = r.refine_nodes(0, num, cef, nbh, factor, threads=4)
        names = sorted(files.params)
        with open(os.path.join(folder, files.params.get("QuantizerFilePath", "quantizer.bin")), "rb") as f:
            qblob = np.frombuffer(f.read(), np.uint8).copy()
        out.update(param_names=np.array(names), param_values=np.array([files.params[n] for n in names]),
                   vectors=files.vectors, graph=files.graph, tree_starts=files.tree_starts, nodes=files.nodes,
                   quantizer_blob=qblob, queries=q, k=np.int32(k), max_checks=np.array(MAX_CHECKS, np.int32),
                   reconstructed=rq.reconstruct(codes[:64], raw.dtype),
                   refine_cef=np.int32(cef), refine_max_check=np.int32(mcr), refine_neighborhood=np.int32(nbh),
                   refine_rng_factor=np.float32(factor), refine_rows=rows, refine_ids=rids.astype(np.int32), refine_dists=rd,
                   ref_isa=np.int32(reflib.ref().ref_isa()))
    os.makedirs(os.path.join(HERE, "quantized"), exist_ok=True)
    path = os.path.join(HERE, "quantized", name + ".npz")
    np.savez_compressed(path, **out)
    print("golden quantized", name, raw.shape, "M", m, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        make(nm)
