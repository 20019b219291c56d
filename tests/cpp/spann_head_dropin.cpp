// TEST INFRASTRUCTURE (row f4: SPANN head search on the device).
//
// Builds a small SPANN index with the UNMODIFIED reference exactly like Test/src/AlgoTest.cpp:23-45 does
// (SelectHead -> BuildHead -> BuildSSDIndex), loads it twice, and
//   A. takes the in-memory HEAD index of one copy (SPANN::Index<T>::GetMemoryIndex()), wraps it in the real VectorIndex
//      subclass SPTAG::B200::Index, and drives both heads exactly as SPANN drives its head:
//        SPANNIndex.cpp:197-203   QueryResultSet<T>(target, SearchInternalResultNum) -> m_index->SearchIndex(query)
//        SPANNIndex.cpp:259-285   RentWorkSpace -> Reset() -> SearchIndexIterativeFromNeareast(first) -> ... -> End
//      comparing every (VID, Dist) bit for bit;
//   B. swaps the wrapped head INTO the second SPANN::Index (m_index is private there, so this test file -- and only
//      this file -- compiles SPANN/Index.h with `private` opened up; object layout is unchanged) and runs the
//      reference's own SPANN::Index::SearchIndex end to end: head search on the B200, posting lists on the CPU as
//      before.  The final results must equal the all-CPU SPANN index's.
// usage: spann_head_dropin <work dir> [--cpu-only]     (exit code = failed checks)
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include <functional>
#include <shared_mutex>

// everything SPANN/Index.h includes, first and untouched (include guards keep them from being re-read below)
#include "inc/Core/Common.h"
#include "inc/Core/VectorIndex.h"
#include "inc/Core/VectorSet.h"
#include "inc/Core/Common/CommonUtils.h"
#include "inc/Core/Common/DistanceUtils.h"
#include "inc/Core/Common/SIMDUtils.h"
#include "inc/Core/Common/QueryResultSet.h"
#include "inc/Core/Common/BKTree.h"
#include "inc/Core/Common/WorkSpacePool.h"
#include "inc/Core/Common/Labelset.h"
#include "inc/Helper/SimpleIniReader.h"
#include "inc/Helper/StringConvert.h"
#include "inc/Helper/ThreadPool.h"
#include "inc/Helper/ConcurrentSet.h"
#include "inc/Helper/VectorSetReader.h"
#include "inc/Helper/Logging.h"
#include "inc/Core/Common/IQuantizer.h"
#include "inc/Core/SPANN/IExtraSearcher.h"
#include "inc/Core/SPANN/Options.h"
#define private public  // test-only: reach SPANN::Index<T>::m_index (no setter exists, SPANN/Index.h:49,71)
#include "inc/Core/SPANN/Index.h"
#undef private

#include "../../sptag_b200/csrc/sptag_vector_index.hpp"

using namespace SPTAG;

class QuietLogger : public Helper::Logger {
public:
    void Logging(const char*, Helper::LogLevel level, const char*, int, const char*, const char* format, ...) override {
        if (level < Helper::LogLevel::LL_Error) return;
        va_list args;
        va_start(args, format);
        std::vfprintf(stderr, format, args);
        va_end(args);
    }
};

static int g_failed = 0;
static void report(const char* name, bool ok, const std::string& detail = "") {
    std::printf("%s %s %s\n", ok ? "PASS" : "FAIL", name, detail.c_str());
    if (!ok) ++g_failed;
}
static size_t count_diff(const BasicResult* a, const BasicResult* b, size_t n) {
    size_t d = 0;
    for (size_t i = 0; i < n; ++i) d += (a[i].VID != b[i].VID || std::memcmp(&a[i].Dist, &b[i].Dist, 4) != 0) ? 1 : 0;
    return d;
}

static ErrorCode build_spann(const std::vector<float>& x, SizeType n, DimensionType dim, const std::string& out) {
    std::shared_ptr<VectorIndex> idx = VectorIndex::CreateInstance(IndexAlgoType::SPANN, VectorValueType::Float);
    if (!idx) return ErrorCode::Fail;
    // Test/src/AlgoTest.cpp:23-42
    idx->SetParameter("IndexAlgoType", "BKT", "Base");
    idx->SetParameter("DistCalcMethod", "L2", "Base");
    idx->SetParameter("isExecute", "true", "SelectHead");
    idx->SetParameter("NumberOfThreads", "8", "SelectHead");
    idx->SetParameter("Ratio", "0.2", "SelectHead");
    idx->SetParameter("isExecute", "true", "BuildHead");
    idx->SetParameter("RefineIterations", "3", "BuildHead");
    idx->SetParameter("NumberOfThreads", "8", "BuildHead");
    idx->SetParameter("isExecute", "true", "BuildSSDIndex");
    idx->SetParameter("BuildSsdIndex", "true", "BuildSSDIndex");
    idx->SetParameter("NumberOfThreads", "8", "BuildSSDIndex");
    idx->SetParameter("PostingPageLimit", "12", "BuildSSDIndex");
    idx->SetParameter("SearchPostingPageLimit", "12", "BuildSSDIndex");
    idx->SetParameter("InternalResultNum", "64", "BuildSSDIndex");
    idx->SetParameter("SearchInternalResultNum", "64", "BuildSSDIndex");
    ByteArray bytes = ByteArray::Alloc(x.size() * sizeof(float));
    std::memcpy(bytes.Data(), x.data(), x.size() * sizeof(float));
    std::shared_ptr<VectorSet> vec(new BasicVectorSet(bytes, VectorValueType::Float, dim, n));
    std::shared_ptr<MetadataSet> meta;
    ErrorCode ec = idx->BuildIndex(vec, meta);
    if (ec != ErrorCode::Success) return ec;
    return idx->SaveIndex(out);
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s workdir [--cpu-only]\n", argv[0]);
        return 100;
    }
    const bool cpu_only = argc > 2 && std::string(argv[2]) == "--cpu-only";
    SetLogger(std::make_shared<QuietLogger>());
    const SizeType n = 20000;
    const DimensionType dim = 64;
    const int nq = 64, rank = 12;
    // low-rank synthetic: x = z A + 0.1 eps
    std::mt19937 gen(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> A((size_t)rank * dim), x((size_t)n * dim), q((size_t)nq * dim);
    for (auto& v : A) v = nd(gen) / std::sqrt((float)rank);
    auto fill = [&](std::vector<float>& dst, int rows) {
        std::vector<float> z(rank);
        for (int r = 0; r < rows; ++r) {
            for (auto& v : z) v = nd(gen);
            for (int d = 0; d < dim; ++d) {
                float s = 0.f;
                for (int t = 0; t < rank; ++t) s += z[t] * A[(size_t)t * dim + d];
                dst[(size_t)r * dim + d] = s + 0.1f * nd(gen);
            }
        }
    };
    fill(x, n);
    fill(q, nq);

    const std::string folder = std::string(argv[1]) + "/spann_index";
    if (build_spann(x, n, dim, folder) != ErrorCode::Success) {
        std::fprintf(stderr, "the reference could not build the SPANN index\n");
        return 101;
    }
    std::shared_ptr<VectorIndex> cpuSpann, mixSpann;
    if (VectorIndex::LoadIndex(folder, cpuSpann) != ErrorCode::Success || VectorIndex::LoadIndex(folder, mixSpann) != ErrorCode::Success)
        return 102;
    auto* cpuS = dynamic_cast<SPANN::Index<float>*>(cpuSpann.get());
    auto* mixS = dynamic_cast<SPANN::Index<float>*>(mixSpann.get());
    if (!cpuS || !mixS) return 103;
    std::shared_ptr<VectorIndex> cpuHead = cpuS->GetMemoryIndex();
    const int internal = cpuS->GetOptions()->m_searchInternalResultNum;
    std::printf("SPANN index: %d vectors, head index %d vectors (%s), SearchInternalResultNum %d\n", n, cpuHead->GetNumSamples(),
                cpuHead->GetIndexAlgoType() == IndexAlgoType::BKT ? "BKT" : "KDT", internal);

    std::shared_ptr<VectorIndex> gpuHead = cpu_only ? mixS->GetMemoryIndex() : B200::Index::Attach(mixS->GetMemoryIndex(), 0);
    if (!gpuHead) {
        std::fprintf(stderr, "B200::Index::Attach(head) failed: %s\n", sptag_b200_last_error());
        return 104;
    }

    {   // A1: SPANNIndex.cpp:197-203
        size_t bad = 0;
        for (int i = 0; i < nq; ++i) {
            COMMON::QueryResultSet<float> ra(q.data() + (size_t)i * dim, internal), rb(q.data() + (size_t)i * dim, internal);
            cpuHead->SearchIndex(ra);
            gpuHead->SearchIndex(rb);
            bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)internal);
        }
        report("head SearchIndex(K = SearchInternalResultNum)", bad == 0, std::to_string(bad) + " results differ");
    }
    {   // A2: SPANNIndex.cpp:259-285
        size_t bad = 0;
        bool flags = true;
        for (int i = 0; i < nq && i < 24; ++i) {
            const float* t = q.data() + (size_t)i * dim;
            std::unique_ptr<COMMON::WorkSpace> wa = cpuHead->RentWorkSpace(internal), wb = gpuHead->RentWorkSpace(internal);
            COMMON::QueryResultSet<float> ra(t, internal), rb(t, internal);
            for (int call = 0; call < 5; ++call) {
                ra.Reset();
                rb.Reset();
                const bool fa = cpuHead->SearchIndexIterativeFromNeareast(ra, wa.get(), call == 0);
                const bool fb = gpuHead->SearchIndexIterativeFromNeareast(rb, wb.get(), call == 0);
                flags = flags && (fa == fb);
                bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)internal);
            }
            cpuHead->SearchIndexIterativeEnd(std::move(wa));
            gpuHead->SearchIndexIterativeEnd(std::move(wb));
        }
        report("head SearchIndexIterativeFromNeareast x5", flags && bad == 0, std::to_string(bad) + " results differ");
    }
    {   // B: the reference's own SPANN::Index::SearchIndex with the head swapped for the device-backed one
        mixS->m_index = gpuHead;
        size_t bad = 0;
        const int k = 10;
        for (int i = 0; i < nq; ++i) {
            QueryResult ra(q.data() + (size_t)i * dim, k, false), rb(q.data() + (size_t)i * dim, k, false);
            ErrorCode ea = cpuSpann->SearchIndex(ra);
            ErrorCode eb = mixSpann->SearchIndex(rb);
            if (ea != eb) ++bad;
            bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)k);
        }
        report("SPANN::Index::SearchIndex end to end (head on the device, postings on the CPU)", bad == 0,
               std::to_string(bad) + " results differ");
    }
    {   // B2: SPANN's iterator (SPANNIndex.cpp:259-285 through SPANNResultIterator)
        size_t bad = 0;
        bool shape = true;
        for (int i = 0; i < 12; ++i) {
            std::shared_ptr<ResultIterator> ia = cpuSpann->GetIterator(q.data() + (size_t)i * dim);
            std::shared_ptr<ResultIterator> ib = mixSpann->GetIterator(q.data() + (size_t)i * dim);
            if (!ia || !ib) {
                shape = shape && (!ia && !ib);
                continue;
            }
            for (int r = 0; r < 3; ++r) {
                std::shared_ptr<QueryResult> ra = ia->Next(8), rb = ib->Next(8);
                if (ra->GetResultNum() != rb->GetResultNum()) {
                    shape = false;
                    continue;
                }
                bad += count_diff(ra->GetResults(), rb->GetResults(), (size_t)ra->GetResultNum());
            }
            ia->Close();
            ib->Close();
        }
        report("SPANN GetIterator / Next end to end", shape && bad == 0, std::to_string(bad) + " results differ");
    }
    std::printf("%d check(s) failed\n", g_failed);
    return g_failed;
}
