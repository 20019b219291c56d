// TEST INFRASTRUCTURE.  Drives SPTAG::B200::Index (sptag_b200/csrc/sptag_vector_index.hpp) -- the real
// `SPTAG::VectorIndex` subclass -- through the reference's own types and call patterns, next to the unmodified
// reference index it wraps, and compares every returned BasicResult bit for bit.
//   compiled against /root/reference/AnnService headers, linked with oracle/_ref/libsptag_ref.so (the reference,
//   for VectorIndex::LoadIndex and the CPU side of each comparison) and sptag_b200/lib/libsptag_b200.so.
// usage: vector_index_dropin <index folder> <queries.bin> <nq> <k> <maxcheck>
// Exit code = number of failed checks; one line per check on stdout.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "inc/Core/VectorIndex.h"
#include "inc/Helper/Logging.h"
#include "inc/Core/MetadataSet.h"
#include "inc/Core/ResultIterator.h"
#include "inc/Core/Common/QueryResultSet.h"

#include "../../sptag_b200/csrc/sptag_vector_index.hpp"

using namespace SPTAG;

// the reference logs every parameter it loads at Info level
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

static bool same_result(const BasicResult& a, const BasicResult& b, bool with_meta) {
    if (a.VID != b.VID || std::memcmp(&a.Dist, &b.Dist, 4) != 0) return false;
    if (with_meta) {
        if (a.Meta.Length() != b.Meta.Length()) return false;
        if (a.Meta.Length() && std::memcmp(a.Meta.Data(), b.Meta.Data(), a.Meta.Length()) != 0) return false;
    }
    return true;
}

static size_t count_diff(const BasicResult* a, const BasicResult* b, size_t n, bool with_meta) {
    size_t d = 0;
    for (size_t i = 0; i < n; ++i) d += same_result(a[i], b[i], with_meta) ? 0 : 1;
    return d;
}

// metadata "m<i>" for vector i, as a MemMetadataSet over one blob + an offset table (MetadataSet.h:84-92)
static MetadataSet* make_metadata(SizeType n) {
    std::string blob;
    std::vector<std::uint64_t> offsets((size_t)n + 1);
    for (SizeType i = 0; i < n; ++i) {
        offsets[(size_t)i] = blob.size();
        blob += "m" + std::to_string(i);
    }
    offsets[(size_t)n] = blob.size();
    ByteArray meta = ByteArray::Alloc(blob.size());
    std::memcpy(meta.Data(), blob.data(), blob.size());
    ByteArray offs = ByteArray::Alloc(offsets.size() * sizeof(std::uint64_t));
    std::memcpy(offs.Data(), offsets.data(), offsets.size() * sizeof(std::uint64_t));
    return new MemMetadataSet(meta, offs, n);
}

int main(int argc, char** argv) {
    if (argc != 6) {
        std::fprintf(stderr, "usage: %s folder queries.bin nq k maxcheck\n", argv[0]);
        return 100;
    }
    const int nq = std::atoi(argv[3]), k = std::atoi(argv[4]);
    SetLogger(std::make_shared<QuietLogger>());
    std::shared_ptr<VectorIndex> cpu;
    if (VectorIndex::LoadIndex(argv[1], cpu) != ErrorCode::Success || !cpu) {
        std::fprintf(stderr, "reference LoadIndex failed\n");
        return 101;
    }
    const size_t qsize = GetValueTypeSize(cpu->GetVectorValueType()) * (size_t)cpu->GetFeatureDim();
    std::vector<char> q((size_t)nq * qsize);
    FILE* f = std::fopen(argv[2], "rb");
    if (!f || std::fread(q.data(), 1, q.size(), f) != q.size()) return 102;
    std::fclose(f);
    const bool bkt = cpu->GetIndexAlgoType() == IndexAlgoType::BKT;

    // a second, untouched load is the CPU side of the comparisons; `gpu` wraps the first
    std::shared_ptr<VectorIndex> ref;
    if (VectorIndex::LoadIndex(argv[1], ref) != ErrorCode::Success) return 103;
    std::shared_ptr<VectorIndex> gpu = B200::Index::Attach(cpu, 0);
    if (!gpu) {
        std::fprintf(stderr, "B200::Index::Attach failed: %s\n", sptag_b200_last_error());
        return 104;
    }
    // the one seam existing callers hold: std::shared_ptr<VectorIndex>
    ref->SetParameter("MaxCheck", argv[5]);
    gpu->SetParameter("MaxCheck", argv[5]);
    report("parameters reach the wrapped index", cpu->GetParameter("MaxCheck") == std::string(argv[5]));

    const size_t n = (size_t)nq * k;
    {   // VectorIndex::SearchIndex(batch), default-constructed BasicResults, no metadata
        std::vector<BasicResult> a(n), b(n);
        ref->SearchIndex(q.data(), nq, k, false, a.data());
        ErrorCode ec = gpu->SearchIndex(q.data(), nq, k, false, b.data());
        report("SearchIndex(batch)", ec == ErrorCode::Success && count_diff(a.data(), b.data(), n, false) == 0,
               std::to_string(count_diff(a.data(), b.data(), n, false)) + " of " + std::to_string(n) + " results differ");
    }
    ref->SetMetadata(make_metadata(ref->GetNumSamples()));
    gpu->SetMetadata(make_metadata(gpu->GetNumSamples()));
    {   // with metadata (BKTIndex.cpp:611-618)
        std::vector<BasicResult> a(n), b(n);
        ref->SearchIndex(q.data(), nq, k, true, a.data());
        gpu->SearchIndex(q.data(), nq, k, true, b.data());
        bool any_meta = false;
        for (size_t i = 0; i < n; ++i) any_meta = any_meta || b[i].Meta.Length() > 0;
        report("SearchIndex(batch, withMeta)", any_meta && count_diff(a.data(), b.data(), n, true) == 0);
    }
    {   // AnnIndex::BatchSearch (Wrappers/src/CoreInterface.cpp:229-238): one QueryResult of nq*k slots as the buffer
        QueryResult ra(q.data(), nq * k, true), rb(q.data(), nq * k, true);
        ref->SearchIndex(q.data(), nq, k, true, ra.GetResults());
        gpu->SearchIndex(q.data(), nq, k, true, rb.GetResults());
        report("AnnIndex::BatchSearch pattern", count_diff(ra.GetResults(), rb.GetResults(), n, true) == 0);
    }
    {   // AnnIndex::Search / SearchWithMetaData (CoreInterface.cpp:206-227): SearchIndex(QueryResult&), both p_searchDeleted
        size_t bad = 0;
        for (int flag = 0; flag < 2; ++flag)
            for (int i = 0; i < nq && i < 64; ++i) {
                QueryResult ra(q.data() + (size_t)i * qsize, k, true), rb(q.data() + (size_t)i * qsize, k, true);
                ref->SearchIndex(ra, flag != 0);
                gpu->SearchIndex(rb, flag != 0);
                bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)k, true);
            }
        report("SearchIndex(QueryResult&, p_searchDeleted)", bad == 0, std::to_string(bad) + " results differ");
    }
    {   // SearchIndexWithFilter: a predicate on the metadata bytes (Test/src/FilterTest.cpp pattern)
        auto keep = [](const ByteArray& meta) -> bool {
            return meta.Length() > 1 && ((meta.Data()[meta.Length() - 1] - '0') % 3) != 0;
        };
        size_t bad = 0;
        bool codes_match = true;
        for (int mc = 0; mc < 2; ++mc)
            for (int i = 0; i < nq && i < 48; ++i) {
                QueryResult ra(q.data() + (size_t)i * qsize, k, true), rb(q.data() + (size_t)i * qsize, k, true);
                ErrorCode ea = ref->SearchIndexWithFilter(ra, keep, mc ? 2048 : 0, false);
                ErrorCode eb = gpu->SearchIndexWithFilter(rb, keep, mc ? 2048 : 0, false);
                codes_match = codes_match && (ea == eb);
                if (ea == ErrorCode::Success) bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)k, true);
            }
        report("SearchIndexWithFilter", codes_match && bad == 0, std::to_string(bad) + " results differ");
    }
    {   // RefineSearchIndex with a base vector as the target (NeighborhoodGraph::RefineNode, NeighborhoodGraph.h:534-545)
        size_t bad = 0;
        const int rk = 33;
        for (SizeType node = 0; node < 40 && node < ref->GetNumSamples(); node += 3) {
            QueryResult ra(ref->GetSample(node), rk, false), rb(ref->GetSample(node), rk, false);
            ref->RefineSearchIndex(ra, false);
            gpu->RefineSearchIndex(rb, false);
            bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)rk, false);
        }
        report("RefineSearchIndex", bad == 0, std::to_string(bad) + " results differ");
    }
    {   // GetIterator -> the reference's own ResultIterator class on top of the overridden virtuals
        size_t bad = 0;
        bool shape = true;
        for (int i = 0; i < nq && i < 24; ++i) {
            std::shared_ptr<ResultIterator> ia = ref->GetIterator(q.data() + (size_t)i * qsize, false);
            std::shared_ptr<ResultIterator> ib = gpu->GetIterator(q.data() + (size_t)i * qsize, false);
            if (!ia || !ib) {
                shape = shape && (!ia && !ib);  // KDT: both nullptr
                continue;
            }
            const int batches[4] = {7, 7, 3, 5};
            for (int r = 0; r < 4; ++r) {
                std::shared_ptr<QueryResult> ra = ia->Next(batches[r]), rb = ib->Next(batches[r]);
                if (ra->GetResultNum() != rb->GetResultNum() || ia->GetRelaxedMono() != ib->GetRelaxedMono()) {
                    shape = false;
                    continue;
                }
                for (int j = 0; j < ra->GetResultNum(); ++j) {
                    if (!same_result(*ra->GetResult(j), *rb->GetResult(j), true)) ++bad;
                    if (ra->GetResult(j)->RelaxedMono != rb->GetResult(j)->RelaxedMono) ++bad;
                }
            }
            ia->Close();
            ib->Close();
        }
        report("GetIterator / ResultIterator::Next", shape && bad == 0, std::to_string(bad) + " results differ");
    }
    if (bkt) {  // SPANN's head-index pattern (SPANNIndex.cpp:259-285): Rent -> FromNeareast(first) -> FromNeareast... -> End
        size_t bad = 0;
        bool flags = true;
        const int hk = 32;
        for (int i = 0; i < nq && i < 16; ++i) {
            const void* t = q.data() + (size_t)i * qsize;
            std::unique_ptr<COMMON::WorkSpace> wa = ref->RentWorkSpace(hk), wb = gpu->RentWorkSpace(hk);
            QueryResult ra(t, hk, false), rb(t, hk, false);
            for (int call = 0; call < 4; ++call) {
                ra.Reset();
                rb.Reset();
                const bool fa = ref->SearchIndexIterativeFromNeareast(ra, wa.get(), call == 0, false);
                const bool fb = gpu->SearchIndexIterativeFromNeareast(rb, wb.get(), call == 0, false);
                flags = flags && (fa == fb);
                bad += count_diff(ra.GetResults(), rb.GetResults(), (size_t)hk, false);
            }
            ref->SearchIndexIterativeEnd(std::move(wa));
            gpu->SearchIndexIterativeEnd(std::move(wb));
        }
        report("SearchIndexIterativeFromNeareast (SPANN head pattern)", flags && bad == 0, std::to_string(bad) + " results differ");
    }
    {   // a mutation through the wrapper re-synchronises the device copy before the next search
        const SizeType victim = 3;
        ref->DeleteIndex(victim);
        gpu->DeleteIndex(victim);
        std::vector<BasicResult> a(n), b(n);
        ref->SearchIndex(q.data(), nq, k, false, a.data());
        gpu->SearchIndex(q.data(), nq, k, false, b.data());
        bool victim_gone = true;
        for (size_t i = 0; i < n; ++i) victim_gone = victim_gone && b[i].VID != victim;
        report("DeleteIndex then search", victim_gone && count_diff(a.data(), b.data(), n, false) == 0);
    }
    std::printf("%d check(s) failed\n", g_failed);
    return g_failed;
}
