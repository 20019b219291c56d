// Drives the C++ adapter (sptag_b200/csrc/vector_index_adapter.hpp) exactly like the reference's own
// callers drive VectorIndex: LoadIndex -> SetParameter -> SearchIndex(batch) -> read BasicResult.
// usage: adapter_search <index folder> <queries.f32> <nq> <dim> <k> <maxcheck> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sptag_b200/csrc/vector_index_adapter.hpp"

using namespace SPTAG_B200;

int main(int argc, char** argv) {
    if (argc != 8) {
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
    return 0;
}
