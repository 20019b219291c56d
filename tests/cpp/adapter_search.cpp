// Drives the C++ adapter (sptag_b200/csrc/vector_index_adapter.hpp) exactly like the reference's own
// callers drive VectorIndex: LoadIndex -> SetParameter -> SearchIndex(batch) -> read BasicResult.
// usage: adapter_search <index folder> <queries.f32> <nq> <dim> <k> <maxcheck> <out.bin> [<refine.bin> <cef> <nodes>]
// With the optional arguments it also runs the builder-side calls: RefineSearchIndex on node 0 and one
// RefineGraphPass (not installed), and writes [nodes x degree] new rows followed by node 0's CEF+1 (VID, Dist) pairs.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sptag_b200/csrc/vector_index_adapter.hpp"

using namespace SPTAG_B200;

int main(int argc, char** argv) {
    if (argc != 8 && argc != 11) {
        std::fprintf(stderr, "usage: %s folder queries.f32 nq dim k maxcheck out.bin\n", argv[0]);
        return 2;
    }
    const int nq = std::atoi(argv[3]), dim = std::atoi(argv[4]), k = std::atoi(argv[5]);
    std::vector<float> q((size_t)nq * dim);
    FILE* f = std::fopen(argv[2], "rb");
    if (!f || std::fread(q.data(), 4, q.size(), f) != q.size()) return 3;
    std::fclose(f);

    std::shared_ptr<VectorIndex> index;
    ErrorCode ec = VectorIndex::LoadIndex(argv[1], index);
    if (ec != ErrorCode::Success) {
        std::fprintf(stderr, "LoadIndex failed: 0x%x (%s)\n", (unsigned)ec, sptag_b200_last_error());
        return 4;
    }
    if (index->GetFeatureDim() != dim) return 5;
    index->SetParameter("MaxCheck", argv[6]);

    // batched overload, caller-owned default-constructed results (VID -1, Dist MaxDist)
    std::vector<BasicResult> res((size_t)nq * k);
    ec = index->SearchIndex(q.data(), nq, k, false, res.data());
    if (ec != ErrorCode::Success) return 6;

    // single-query overload on the first query must agree with row 0 of the batch
    QueryResult one(q.data(), k, false);
    if (index->SearchIndex(one) != ErrorCode::Success) return 7;
    for (int i = 0; i < k; ++i)
        if (one.GetResult(i)->VID != res[i].VID || one.GetResult(i)->Dist != res[i].Dist) return 8;

    f = std::fopen(argv[7], "wb");
    for (auto& r : res) {
        std::fwrite(&r.VID, 4, 1, f);
        std::fwrite(&r.Dist, 4, 1, f);
    }
    std::fclose(f);

    if (argc == 11) {
        const int cef = std::atoi(argv[9]), nodes = std::atoi(argv[10]);
        const int degree = sptag_b200_graph_degree(index->Handle());
        std::vector<std::int32_t> rows((size_t)index->GetNumSamples() * degree);
        if (index->RefineGraphPass(cef, 1.0f, rows.data(), /*install=*/false) != ErrorCode::Success) return 9;
        QueryResult rq(nullptr, cef + 1, false);
        if (index->RefineSearchIndex(0, rq) != ErrorCode::Success) return 10;
        f = std::fopen(argv[8], "wb");
        std::fwrite(rows.data(), 4, (size_t)nodes * degree, f);
        for (int i = 0; i <= cef; ++i) {
            std::fwrite(&rq.GetResult(i)->VID, 4, 1, f);
            std::fwrite(&rq.GetResult(i)->Dist, 4, 1, f);
        }
        std::fclose(f);

        // ResultIterator driven like Test/src/IterativeScanTest.cpp drives the reference's: GetIterator, Next(batch)
        // twice, Close; the (VID, Dist, RelaxedMono) triples of the first query are appended to refine.bin
        std::shared_ptr<ResultIterator> it = index->GetIterator(q.data());
        if (!it) return 11;
        f = std::fopen(argv[8], "ab");
        for (int round = 0; round < 2; ++round) {
            std::shared_ptr<QueryResult> r = it->Next(5);
            if (r->GetResultNum() != 5) return 12;
            for (int i = 0; i < r->GetResultNum(); ++i) {
                std::int32_t mono = r->GetResult(i)->RelaxedMono ? 1 : 0;
                std::fwrite(&r->GetResult(i)->VID, 4, 1, f);
                std::fwrite(&r->GetResult(i)->Dist, 4, 1, f);
                std::fwrite(&mono, 4, 1, f);
            }
        }
        it->Close();
        std::fclose(f);
    }
    return 0;
}
