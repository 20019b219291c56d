// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" shim over the UNMODIFIED reference library (compiled from /root/reference by
// oracle/Makefile `ref`).  It only calls the reference's own public API:
//   VectorIndex::CreateInstance / SetParameter / BuildIndex / SaveIndex / LoadIndex
//                                   (AnnService/inc/Core/VectorIndex.h:28-214)
//   VectorIndex::SearchIndex(const void*,int,int,bool,BasicResult*)   (VectorIndex.cpp:454-463)
//   VectorIndex::SearchIndex(QueryResult&)                            (BKTIndex.cpp:595-620)
//   COMMON::DistanceCalcSelector<T> / DistanceUtils::Compute*_{SSE,AVX,AVX512}
//                                   (inc/Core/Common/DistanceUtils.h:118-163)
// so that python (ctypes) tests and bench.py's cpu_baseline / --impl reference leg can run the
// reference itself.  Nothing in the product library links against this file.
#include "inc/Core/VectorIndex.h"
#include "inc/Core/SearchQuery.h"
#include "inc/Core/Common/DistanceUtils.h"
#include "inc/Core/Common/InstructionUtils.h"
#include "inc/Core/Common/WorkSpace.h"
#include "inc/Core/Common/IQuantizer.h"
#include "inc/Core/MetadataSet.h"
#include "inc/Core/Common/QueryResultSet.h"
#include "inc/Core/ResultIterator.h"
#include "inc/Core/Common/RelativeNeighborhoodGraph.h"
#include "inc/Helper/Logging.h"

#include <omp.h>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace SPTAG;

namespace {

struct RefHandle {
    std::shared_ptr<VectorIndex> index;
};

// silence the reference's Info-level chatter (build progress, "Delete workspace happens!")
class QuietLogger : public Helper::Logger {
public:
    explicit QuietLogger(Helper::LogLevel lvl) : m_level(lvl) {}
    void Logging(const char* title, Helper::LogLevel level, const char* file, int line,
                 const char* func, const char* format, ...) override {
        if (level < m_level) return;
        va_list args;
        va_start(args, format);
        std::fprintf(stderr, "[ref:%d] ", (int)level);
        std::vfprintf(stderr, format, args);
        va_end(args);
    }
private:
    Helper::LogLevel m_level;
};

// Work-space factory that lets us read the reference's per-query counters
// (WorkSpace.h:303-308) without patching it.  SetWorkSpaceFactory (BKT/Index.h:199-216)
// dynamic_casts from IWorkSpaceFactory<IWorkSpace> to IWorkSpaceFactory<WorkSpace>, so the
// object must inherit both (SURVEY.md 8a').
struct SideA : COMMON::IWorkSpaceFactory<COMMON::IWorkSpace> {
    std::unique_ptr<COMMON::IWorkSpace> GetWorkSpace() override { return nullptr; }
    void ReturnWorkSpace(std::unique_ptr<COMMON::IWorkSpace>) override {}
    virtual ~SideA() {}
};
struct SideB : COMMON::IWorkSpaceFactory<COMMON::WorkSpace> {
    static thread_local std::unique_ptr<COMMON::WorkSpace> tl_ws;
    static thread_local COMMON::WorkSpace* tl_last;
    std::unique_ptr<COMMON::WorkSpace> GetWorkSpace() override { return std::move(tl_ws); }
    void ReturnWorkSpace(std::unique_ptr<COMMON::WorkSpace> ws) override {
        tl_last = ws.get();
        tl_ws = std::move(ws);
    }
    virtual ~SideB() {}
};
thread_local std::unique_ptr<COMMON::WorkSpace> SideB::tl_ws;
thread_local COMMON::WorkSpace* SideB::tl_last = nullptr;
struct PinnedFactory : SideA, SideB {};

void apply_params(VectorIndex* idx, const char* params) {
    // "Name=Value;Name=Value"
    if (!params) return;
    std::string s(params);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find(';', pos);
        if (end == std::string::npos) end = s.size();
        std::string kv = s.substr(pos, end - pos);
        size_t eq = kv.find('=');
        if (eq != std::string::npos)
            idx->SetParameter(kv.substr(0, eq).c_str(), kv.substr(eq + 1).c_str());
        pos = end + 1;
    }
}

}  // namespace

// One NeighborhoodGraph::RefineNode(index, node, updateNeighbors=false, searchDeleted=false, CEF) per node
// (NeighborhoodGraph.h:534-545) WITHOUT writing the index's graph: the reference's own RefineSearchIndex
// (BKTIndex.cpp:698-711 / KDTIndex.cpp:367-390) on the loaded index, then the reference's own
// RelativeNeighborhoodGraph::RebuildNeighbors (RelativeNeighborhoodGraph.h:20-38) into out_graph.
template <typename T>
static int refine_nodes_t(VectorIndex* idx, int first, int num, int cef, int neighborhood, float rng_factor,
                          int* out_graph, int* res_ids, float* res_dists) {
    COMMON::RelativeNeighborhoodGraph rng;
    rng.m_iNeighborhoodSize = neighborhood;
    rng.m_fRNGFactor = rng_factor;
    const int k = cef + 1;
#pragma omp parallel for schedule(dynamic, 10)
    for (int i = 0; i < num; ++i) {
        const int node = first + i;
        COMMON::QueryResultSet<T> query((const T*)idx->GetSample(node), k);
        void* rec_query = nullptr;
        if (idx->m_pQuantizer) {  // RefineNode on a quantized index (NeighborhoodGraph.h:538-543), verbatim
            rec_query = ALIGN_ALLOC(idx->m_pQuantizer->ReconstructSize());
            idx->m_pQuantizer->ReconstructVector((const uint8_t*)query.GetTarget(), rec_query);
            query.SetTarget((T*)rec_query, idx->m_pQuantizer);
        }
        idx->RefineSearchIndex(query, false);
        rng.RebuildNeighbors(idx, node, out_graph + (size_t)i * neighborhood, query.GetResults(), k);
        if (rec_query) ALIGN_FREE(rec_query);
        for (int j = 0; j < k; ++j) {
            if (res_ids) res_ids[(size_t)i * k + j] = query.GetResult(j)->VID;
            if (res_dists) res_dists[(size_t)i * k + j] = query.GetResult(j)->Dist;
        }
    }
    return 0;
}

extern "C" {

// 512 / 256 / 128 / 0: which DistanceUtils variant the reference's cpuid dispatch selects here
int ref_isa(void) {
    if (COMMON::InstructionSet::AVX512()) return 512;
    if (COMMON::InstructionSet::AVX2() || COMMON::InstructionSet::AVX()) return 256;
    if (COMMON::InstructionSet::SSE2() || COMMON::InstructionSet::SSE()) return 128;
    return 0;
}

void ref_quiet(int min_level) {
    SetLogger(std::make_shared<QuietLogger>((Helper::LogLevel)min_level));
}

// algo: 0 BKT, 1 KDT.  value_type: 0 int8, 1 uint8, 2 int16, 3 float.  metric: 0 L2, 1 Cosine.
void* ref_build(int algo, int value_type, int metric, const void* data, int n, int dim,
                int threads, const char* params) {
    auto idx = VectorIndex::CreateInstance((IndexAlgoType)algo, (VectorValueType)value_type);
    if (!idx) return nullptr;
    idx->SetParameter("DistCalcMethod", metric == 0 ? "L2" : (metric == 1 ? "Cosine" : "InnerProduct"));
    idx->SetParameter("NumberOfThreads", std::to_string(threads).c_str());
    apply_params(idx.get(), params);
    if (idx->BuildIndex(data, n, dim) != ErrorCode::Success) return nullptr;
    auto* h = new RefHandle();
    h->index = idx;
    return h;
}

int ref_save(void* h, const char* folder) {
    return (int)((RefHandle*)h)->index->SaveIndex(std::string(folder));
}

void* ref_load(const char* folder) {
    std::shared_ptr<VectorIndex> idx;
    if (VectorIndex::LoadIndex(std::string(folder), idx) != ErrorCode::Success || !idx) return nullptr;
    auto* h = new RefHandle();
    h->index = idx;
    return h;
}

// VectorIndex::LoadIndex(config, blobs, index) (VectorIndex.cpp:745-792): the reference loads an index from memory
// blobs that hold exactly the bytes of vectors.bin / tree.bin / graph.bin; `config` is the indexloader.ini text.
// The index keeps pointing INTO the blobs (Dataset::Load(char*), Dataset.h:191-204): they must outlive it.  Used where an index exists only in memory
// (bench.py's shard legs: eight 8-GB folders would not fit a scratch disk).
void* ref_load_memory(const char* config, const void* vectors, unsigned long long vectors_len, const void* tree,
                      unsigned long long tree_len, const void* graph, unsigned long long graph_len) {
    std::vector<ByteArray> blobs;
    blobs.push_back(ByteArray((std::uint8_t*)vectors, vectors_len, false));
    blobs.push_back(ByteArray((std::uint8_t*)tree, tree_len, false));
    blobs.push_back(ByteArray((std::uint8_t*)graph, graph_len, false));
    std::shared_ptr<VectorIndex> idx;
    if (VectorIndex::LoadIndex(std::string(config), blobs, idx) != ErrorCode::Success || !idx) return nullptr;
    auto* h = new RefHandle();
    h->index = idx;
    return h;
}

void ref_free(void* h) { delete (RefHandle*)h; }

int ref_set_param(void* h, const char* name, const char* value) {
    return (int)((RefHandle*)h)->index->SetParameter(name, value);
}

int ref_get_param(void* h, const char* name, char* out, int cap) {
    std::string v = ((RefHandle*)h)->index->GetParameter(name);
    std::snprintf(out, cap, "%s", v.c_str());
    return (int)v.size();
}

int ref_num_samples(void* h) { return ((RefHandle*)h)->index->GetNumSamples(); }
int ref_dim(void* h) { return ((RefHandle*)h)->index->GetFeatureDim(); }
const void* ref_sample(void* h, int i) { return ((RefHandle*)h)->index->GetSample(i); }

// THE reference call the product replaces: VectorIndex::SearchIndex(batch) (VectorIndex.cpp:454-463).
// Results are pre-initialised the way a default-constructed BasicResult is (SearchResult.h:72).
// Wall-clock seconds of the SearchIndex call only are written to *seconds.
int ref_search_batch(void* h, const void* queries, int nq, int k, int threads,
                     int* ids, float* dists, double* seconds) {
    auto& idx = ((RefHandle*)h)->index;
    std::vector<BasicResult> res((size_t)nq * k);
    if (threads > 0) omp_set_num_threads(threads);
    auto t0 = std::chrono::steady_clock::now();
    ErrorCode ec = idx->SearchIndex(queries, nq, k, false, res.data());
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].VID;
        dists[i] = res[i].Dist;
    }
    return (int)ec;
}

// ---- quantized indexes (PQ / OPQ): IQuantizer.h, PQQuantizer.h:110-180, OPQQuantizer.h:96-121 ----

// Build a BKT/KDT index over uint8 PQ codes with the quantizer loaded from `quantizer_file`
// (VectorIndex::LoadQuantizer, VectorIndex.cpp:548-563; the file format is PQQuantizer::SaveQuantizer /
// OPQQuantizer::SaveQuantizer).  codes: n x M uint8.
void* ref_build_quantized(int algo, int metric, const char* quantizer_file, const void* codes, int n, int m,
                          int threads, const char* params) {
    auto idx = VectorIndex::CreateInstance((IndexAlgoType)algo, VectorValueType::UInt8);
    if (!idx) return nullptr;
    if (idx->LoadQuantizer(std::string(quantizer_file)) != ErrorCode::Success) return nullptr;
    idx->SetParameter("DistCalcMethod", metric == 0 ? "L2" : "Cosine");
    idx->SetParameter("NumberOfThreads", std::to_string(threads).c_str());
    apply_params(idx.get(), params);
    // SetParameter("DistCalcMethod") re-selects m_fComputeDistance from the quantizer (BKTIndex.cpp:1001)
    if (idx->BuildIndex(codes, n, m) != ErrorCode::Success) return nullptr;
    auto* h = new RefHandle();
    h->index = idx;
    return h;
}

// Stand-alone quantizer (no index): returns an opaque handle holding shared_ptr<IQuantizer>
void* ref_quantizer_load(const char* quantizer_file) {
    auto ptr = SPTAG::f_createIO();
    if (!ptr->Initialize(quantizer_file, std::ios::binary | std::ios::in)) return nullptr;
    auto q = COMMON::IQuantizer::LoadIQuantizer(ptr);
    if (!q) return nullptr;
    return new std::shared_ptr<COMMON::IQuantizer>(q);
}
int ref_quantizer_m(void* q) { return (*(std::shared_ptr<COMMON::IQuantizer>*)q)->GetNumSubvectors(); }
int ref_quantizer_reconstruct_dim(void* q) { return (*(std::shared_ptr<COMMON::IQuantizer>*)q)->ReconstructDim(); }
// IQuantizer::QuantizeVector(vec, out, ADC=false): raw vectors (ReconstructSize() bytes each) -> M code bytes each
void ref_quantizer_encode(void* q, const void* raw, int nvec, unsigned char* out) {
    auto& quant = *(std::shared_ptr<COMMON::IQuantizer>*)q;
    const size_t rs = quant->ReconstructSize(), m = quant->GetNumSubvectors();
#pragma omp parallel for
    for (int i = 0; i < nvec; ++i) quant->QuantizeVector((const std::uint8_t*)raw + i * rs, out + i * m, false);
}
// IQuantizer::ReconstructVector: M code bytes -> ReconstructSize() bytes of the reconstruct type
void ref_quantizer_reconstruct(void* q, const unsigned char* codes, int nvec, void* out) {
    auto& quant = *(std::shared_ptr<COMMON::IQuantizer>*)q;
    const size_t rs = quant->ReconstructSize(), m = quant->GetNumSubvectors();
    for (int i = 0; i < nvec; ++i) quant->ReconstructVector(codes + i * m, (std::uint8_t*)out + i * rs);
}
// IQuantizer::L2Distance on two code vectors (SDC table lookups when ADC is off, PQQuantizer.h:110-128)
float ref_quantizer_l2(void* q, const unsigned char* a, const unsigned char* b) {
    return (*(std::shared_ptr<COMMON::IQuantizer>*)q)->L2Distance(a, b);
}

// VectorIndex::SetQuantizerADC (VectorIndex.h:136-138): asymmetric distance (query -> per-sub-vector distance table,
// PQQuantizer.h:114-119, :141-157) instead of the default symmetric SDC table look-up.  Not serialized with the index.
void ref_set_adc(void* h, int enable) { ((RefHandle*)h)->index->SetQuantizerADC(enable != 0); }

// Per-query overload over a batch of RAW queries (what IndexSearcher does for quantized indexes,
// IndexSearcher/main.cpp:179-206; the batched overload strides by code bytes and is not usable here,
// SURVEY.md 8b).  stride_bytes = bytes between consecutive raw queries.
int ref_search_each(void* h, const void* queries, int nq, long long stride_bytes, int k, int threads,
                    int* ids, float* dists, double* seconds) {
    auto& idx = ((RefHandle*)h)->index;
    if (threads > 0) omp_set_num_threads(threads);
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 10)
    for (int i = 0; i < nq; ++i) {
        QueryResult res((const char*)queries + (size_t)i * stride_bytes, k, false);
        idx->SearchIndex(res);
        for (int j = 0; j < k; ++j) {
            ids[(size_t)i * k + j] = res.GetResult(j)->VID;
            dists[(size_t)i * k + j] = res.GetResult(j)->Dist;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}

// VectorIndex::SearchIndexWithFilter (BKTIndex.cpp:622-647) with a filter over per-vector metadata.  The shim attaches
// a MemMetadataSet whose record i is the 4-byte id i (VectorIndex::SetMetadata), and the filter callback looks the id
// up in `allowed` (1 = may be returned).  max_check 0 = the index's MaxCheck, as in the reference.
int ref_search_filtered(void* h, const void* queries, int nq, long long stride_bytes, int k, const unsigned char* allowed,
                        int max_check, int threads, int* ids, float* dists) {
    auto& idx = ((RefHandle*)h)->index;
    const int n = idx->GetNumSamples();
    if (idx->GetMetadata() == nullptr) {
        ByteArray meta = ByteArray::Alloc((size_t)n * 4);
        ByteArray offs = ByteArray::Alloc(((size_t)n + 1) * sizeof(std::uint64_t));
        for (int i = 0; i < n; ++i) {
            std::memcpy(meta.Data() + (size_t)i * 4, &i, 4);
            ((std::uint64_t*)offs.Data())[i] = (std::uint64_t)i * 4;
        }
        ((std::uint64_t*)offs.Data())[n] = (std::uint64_t)n * 4;
        idx->SetMetadata(new MemMetadataSet(meta, offs, n));
    }
    std::function<bool(const ByteArray&)> f = [allowed](const ByteArray& m) -> bool {
        int id;
        std::memcpy(&id, m.Data(), 4);
        return allowed[id] != 0;
    };
    if (threads > 0) omp_set_num_threads(threads);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 10) reduction(+ : bad)
    for (int i = 0; i < nq; ++i) {
        QueryResult res((const char*)queries + (size_t)i * stride_bytes, k, false);
        if (idx->SearchIndexWithFilter(res, f, max_check) != ErrorCode::Success) bad++;
        for (int j = 0; j < k; ++j) {
            ids[(size_t)i * k + j] = res.GetResult(j)->VID;
            dists[(size_t)i * k + j] = res.GetResult(j)->Dist;
        }
    }
    return bad;
}

// NeighborhoodGraph::RebuildGraph (NeighborhoodGraph.h:404-456) run by the reference itself on a graph handed in as
// [n x stride] rows (2 x neighborhood candidates each), single-threaded: the reference's loop updates its in-degree array
// from every OpenMP thread without synchronisation, so only its one-thread order is a function of the input.
// h: a loaded index with n samples (RebuildGraph ends with a GraphAccuracyEstimation log line that reads them).
int ref_rebuild_graph(void* h, int* graph, int n, int stride, int neighborhood) {
    auto& idx = ((RefHandle*)h)->index;
    if (idx->GetNumSamples() != n) return 1;
    std::vector<char> mem(sizeof(SizeType) + sizeof(DimensionType) + (size_t)n * stride * sizeof(SizeType));
    *(SizeType*)mem.data() = n;
    *(DimensionType*)(mem.data() + sizeof(SizeType)) = stride;
    memcpy(mem.data() + sizeof(SizeType) + sizeof(DimensionType), graph, (size_t)n * stride * sizeof(SizeType));
    COMMON::RelativeNeighborhoodGraph g;
    if (g.LoadGraph(mem.data(), idx->m_iDataBlockSize, idx->m_iDataCapacity) != ErrorCode::Success) return 1;
    g.m_iNeighborhoodSize = neighborhood;  // BuildGraph halves it before the call (NeighborhoodGraph.h:388-390)
    const int before = omp_get_max_threads();
    omp_set_num_threads(1);
    switch (idx->GetVectorValueType()) {
    case VectorValueType::Float: g.RebuildGraph<float>(idx.get()); break;
    case VectorValueType::Int8: g.RebuildGraph<std::int8_t>(idx.get()); break;
    case VectorValueType::UInt8: g.RebuildGraph<std::uint8_t>(idx.get()); break;
    case VectorValueType::Int16: g.RebuildGraph<std::int16_t>(idx.get()); break;
    default: omp_set_num_threads(before); return 1;
    }
    omp_set_num_threads(before);
    for (int i = 0; i < n; ++i) memcpy(graph + (size_t)i * stride, g[i], (size_t)stride * sizeof(SizeType));
    return 0;
}

int ref_refine_nodes(void* h, int first, int num, int cef, int neighborhood, float rng_factor, int threads,
                     int* out_graph, int* res_ids, float* res_dists) {
    auto& idx = ((RefHandle*)h)->index;
    if (threads > 0) omp_set_num_threads(threads);
    switch (idx->GetVectorValueType()) {
    case VectorValueType::Float: return refine_nodes_t<float>(idx.get(), first, num, cef, neighborhood, rng_factor, out_graph, res_ids, res_dists);
    case VectorValueType::Int8: return refine_nodes_t<std::int8_t>(idx.get(), first, num, cef, neighborhood, rng_factor, out_graph, res_ids, res_dists);
    case VectorValueType::UInt8: return refine_nodes_t<std::uint8_t>(idx.get(), first, num, cef, neighborhood, rng_factor, out_graph, res_ids, res_dists);
    case VectorValueType::Int16: return refine_nodes_t<std::int16_t>(idx.get(), first, num, cef, neighborhood, rng_factor, out_graph, res_ids, res_dists);
    default: return 1;
    }
}

// VectorIndex::GetIterator / ResultIterator::Next / Close (ResultIterator.cpp) -- the reference's own iterator object.
void* ref_iter_open(void* h, const void* query) {
    auto& idx = ((RefHandle*)h)->index;
    std::shared_ptr<ResultIterator> it = idx->GetIterator(query);
    if (!it) return nullptr;
    return new std::shared_ptr<ResultIterator>(it);
}

// -> resultCount; ids/dists get `batch` entries (the QueryResult buffer after Next), *relaxed = GetRelaxedMono()
int ref_iter_next(void* itp, int batch, int* ids, float* dists, int* relaxed) {
    auto& it = *(std::shared_ptr<ResultIterator>*)itp;
    std::shared_ptr<QueryResult> res = it->Next(batch);
    const int count = res->GetResultNum();
    for (int j = 0; j < batch; ++j) {
        ids[j] = j < count ? res->GetResult(j)->VID : -1;
        dists[j] = j < count ? res->GetResult(j)->Dist : MaxDist;
    }
    if (relaxed) *relaxed = it->GetRelaxedMono() ? 1 : 0;
    return count;
}

void ref_iter_close(void* itp) {
    auto* p = (std::shared_ptr<ResultIterator>*)itp;
    (*p)->Close();
    delete p;
}

// CPU baseline for iterator scans: one ResultIterator per query, `rounds` x Next(batch), OpenMP over queries.
// Returns the total number of results; *seconds = wall time of the parallel region.
long long ref_iter_scan(void* h, const void* queries, int nq, long long stride_bytes, int batch, int rounds, int threads,
                        double* seconds) {
    auto& idx = ((RefHandle*)h)->index;
    if (threads > 0) omp_set_num_threads(threads);
    long long total = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
    for (int i = 0; i < nq; ++i) {
        std::shared_ptr<ResultIterator> it = idx->GetIterator((const char*)queries + (size_t)i * stride_bytes);
        if (!it) continue;
        for (int r = 0; r < rounds; ++r) total += it->Next(batch)->GetResultNum();
        it->Close();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return total;
}

// VectorIndex::DeleteIndex(const SizeType&) for a list of ids (tombstones, Labelset.h:43-83) -> number of failures
int ref_delete(void* h, const int* ids, int n) {
    auto& idx = ((RefHandle*)h)->index;
    int bad = 0;
    for (int i = 0; i < n; ++i)
        if (idx->DeleteIndex((SizeType)ids[i]) != ErrorCode::Success) bad++;
    return bad;
}

// SearchIndex(QueryResult&, p_searchDeleted) per query (VectorIndex.h:41) and GetIterator(target, p_searchDeleted)
int ref_search_each_flag(void* h, const void* queries, int nq, long long stride_bytes, int k, int threads,
                         int search_deleted, int* ids, float* dists) {
    auto& idx = ((RefHandle*)h)->index;
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 10)
    for (int i = 0; i < nq; ++i) {
        QueryResult res((const char*)queries + (size_t)i * stride_bytes, k, false);
        idx->SearchIndex(res, search_deleted != 0);
        for (int j = 0; j < k; ++j) {
            ids[(size_t)i * k + j] = res.GetResult(j)->VID;
            dists[(size_t)i * k + j] = res.GetResult(j)->Dist;
        }
    }
    return 0;
}

void* ref_iter_open_flag(void* h, const void* query, int search_deleted) {
    auto& idx = ((RefHandle*)h)->index;
    std::shared_ptr<ResultIterator> it = idx->GetIterator(query, search_deleted != 0);
    if (!it) return nullptr;
    return new std::shared_ptr<ResultIterator>(it);
}

// VectorIndex::SearchIndexIterativeFromNeareast (VectorIndex.h:49, BKTIndex.cpp:543-595) driven the way SPANN's
// iterative search drives its head index (SPANNIndex.cpp:259-285): RentWorkSpace(k) once, then one call per batch on a
// freshly Reset() QueryResult of k slots; SearchIndexIterativeEnd returns the work space.
void* ref_nearest_open(void* h, int k) {
    auto& idx = ((RefHandle*)h)->index;
    std::unique_ptr<COMMON::WorkSpace> ws = idx->RentWorkSpace(k);
    return ws.release();
}

int ref_nearest_next(void* h, void* wsp, const void* query, int k, int is_first, int* ids, float* dists) {
    auto& idx = ((RefHandle*)h)->index;
    QueryResult res(query, k, false);
    res.Reset();
    const bool ok = idx->SearchIndexIterativeFromNeareast(res, (COMMON::WorkSpace*)wsp, is_first != 0);
    for (int j = 0; j < k; ++j) {
        ids[j] = res.GetResult(j)->VID;
        dists[j] = res.GetResult(j)->Dist;
    }
    return ok ? 1 : 0;
}

void ref_nearest_close(void* h, void* wsp) {
    auto& idx = ((RefHandle*)h)->index;
    idx->SearchIndexIterativeEnd(std::unique_ptr<COMMON::WorkSpace>((COMMON::WorkSpace*)wsp));
}

// Install the counter-reading factory (single-query stats below need it). Irreversible for h.
int ref_enable_stats(void* h) {
    auto& idx = ((RefHandle*)h)->index;
    std::unique_ptr<COMMON::IWorkSpaceFactory<COMMON::IWorkSpace>> f(static_cast<SideA*>(new PinnedFactory()));
    return (int)idx->SetWorkSpaceFactory(std::move(f));
}

// Single-query VectorIndex::SearchIndex(QueryResult&) + the reference's own counters:
// stats[0]=m_iNumberOfCheckedLeaves, [1]=m_iNumberOfTreeCheckedLeaves, [2]=NGQueue.size(), [3]=SPTQueue.size()
int ref_search_one_stats(void* h, const void* query, int k, int* ids, float* dists, int* stats) {
    auto& idx = ((RefHandle*)h)->index;
    QueryResult res(query, k, false);
    ErrorCode ec = idx->SearchIndex(res);
    for (int i = 0; i < k; ++i) {
        ids[i] = res.GetResult(i)->VID;
        dists[i] = res.GetResult(i)->Dist;
    }
    COMMON::WorkSpace* ws = SideB::tl_last;
    if (stats && ws) {
        stats[0] = ws->m_iNumberOfCheckedLeaves;
        stats[1] = ws->m_iNumberOfTreeCheckedLeaves;
        stats[2] = ws->m_NGQueue.size();
        stats[3] = ws->m_SPTQueue.size();
    }
    return (int)ec;
}

// Distance through the reference's own run-time dispatch (DistanceUtils.h:118-163).
float ref_distance(int metric, int value_type, const void* a, const void* b, int dim) {
    DistCalcMethod m = (DistCalcMethod)metric;
    switch (value_type) {
    case 0: return COMMON::DistanceCalcSelector<std::int8_t>(m)((const std::int8_t*)a, (const std::int8_t*)b, dim);
    case 1: return COMMON::DistanceCalcSelector<std::uint8_t>(m)((const std::uint8_t*)a, (const std::uint8_t*)b, dim);
    case 2: return COMMON::DistanceCalcSelector<std::int16_t>(m)((const std::int16_t*)a, (const std::int16_t*)b, dim);
    default: return COMMON::DistanceCalcSelector<float>(m)((const float*)a, (const float*)b, dim);
    }
}

// Explicit ISA variant for float (isa = 512/256/128/0=scalar template), so the restatement can
// be pinned against every summation tree, not only the one this host dispatches to.
float ref_distance_f32_isa(int isa, int metric, const float* a, const float* b, int dim) {
    using DU = COMMON::DistanceUtils;
    if (metric == 0) {
        if (isa == 512) return DU::ComputeL2Distance_AVX512(a, b, dim);
        if (isa == 256) return DU::ComputeL2Distance_AVX(a, b, dim);
        if (isa == 128) return DU::ComputeL2Distance_SSE(a, b, dim);
        return DU::ComputeL2Distance<float>(a, b, dim);
    }
    if (isa == 512) return DU::ComputeCosineDistance_AVX512(a, b, dim);
    if (isa == 256) return DU::ComputeCosineDistance_AVX(a, b, dim);
    if (isa == 128) return DU::ComputeCosineDistance_SSE(a, b, dim);
    return DU::ComputeCosineDistance<float>(a, b, dim);
}

// many pairs at once: a[i*dim..], b[i*dim..] -> out[i]
void ref_distance_f32_many(int isa, int metric, const float* a, const float* b, int dim, int n, float* out) {
    for (int i = 0; i < n; ++i)
        out[i] = ref_distance_f32_isa(isa, metric, a + (size_t)i * dim, b + (size_t)i * dim, dim);
}

}  // extern "C"
