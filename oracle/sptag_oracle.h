/* oracle/sptag_oracle.h -- TEST INFRASTRUCTURE ONLY (see sptag_oracle.c header). */
#ifndef SPTAG_ORACLE_H_
#define SPTAG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* value types / metrics use the reference's enum order (DefinitionList.h:6-9, :36-38) */
enum { ORA_INT8 = 0, ORA_UINT8 = 1, ORA_INT16 = 2, ORA_FLOAT = 3 };
enum { ORA_L2 = 0, ORA_COSINE = 1, ORA_INNERPRODUCT = 2 };
enum { ORA_BKT = 0, ORA_KDT = 1 };

typedef struct {
    int32_t centerid, childStart, childEnd; /* BKTree.h:25-32 */
} ora_bkt_node;

typedef struct {
    int32_t left, right, split_dim;
    float split_value; /* KDTree.h:22-28 */
} ora_kdt_node;

/* PQ / OPQ quantizer (PQQuantizer.h, OPQQuantizer.h; file layout PQQuantizer.h:226-239, OPQQuantizer.h:133-147).
 * Supported: PQQuantizer<float>, OPQQuantizer<T> (codebooks and rotation are float for every T). */
enum { ORA_Q_NONE = 0, ORA_Q_PQ = 1, ORA_Q_OPQ = 2 };
typedef struct {
    int32_t qtype;          /* ORA_Q_* (DefinitionList.h:44-46 order) */
    int32_t rtype;          /* reconstruct type = element type of raw queries (ORA_INT8 ... ORA_FLOAT) */
    int32_t m, ks, dsub;    /* NumSubvectors, KsPerSubvector, DimPerSubvector */
    int32_t simd_width;     /* DistanceUtils variant used for the tables / encoding */
    int32_t enable_adc;     /* IQuantizer::SetEnableADC: queries become M*Ks distance tables (PQQuantizer.h:114-119) */
    const float* codebooks; /* m * ks * dsub */
    const float* rotation;  /* OPQ: (m*dsub)^2 floats as stored (m_OPQMatrix); NULL for PQ */
    float* sdc;             /* m * ks * ks, caller-allocated, filled by ora_quantizer_init */
    float* rotation_t;      /* OPQ: caller-allocated (m*dsub)^2, filled by ora_quantizer_init */
} ora_quantizer;

typedef struct {
    /* vectors.bin (Dataset.h:146-180): row-major n x dim, no padding */
    int32_t n, dim;
    int32_t value_type, metric;
    const void* vectors;
    /* graph.bin (NeighborhoodGraph.h:606-615): n x degree int32, -1 padded */
    int32_t degree;
    const int32_t* graph;
    /* tree.bin (BKTree.h:635-645 / KDTree.h:123-133) */
    int32_t tree_kind, tree_num, node_count;
    const int32_t* tree_starts;
    const void* nodes;
    /* deletes.bin (Labelset.h:78-83): one byte per vector, 1 = deleted; NULL or num_deleted==0 -> none */
    const int8_t* deleted;
    int32_t num_deleted;
    /* search parameters (BKT/ParameterDefinitionList.h:44-49; KDT/ParameterDefinitionList.h) */
    int32_t max_check;               /* MaxCheck */
    int32_t max_check_refine;        /* MaxCheckForRefineGraph (sizes the work space) */
    int32_t initial_pivots;          /* NumberOfInitialDynamicPivots */
    int32_t other_pivots;            /* NumberOfOtherDynamicPivots */
    int32_t no_better_threshold;     /* ThresholdOfNumberOfContinuousNoBetterPropagation (KDT) */
    /* which DistanceUtils variant to restate: 16 = AVX512, 8 = AVX/AVX2, 4 = SSE, 1 = scalar template */
    int32_t simd_width;
    /* non-NULL: the index holds uint8 PQ codes (value_type ORA_UINT8, dim = m) and queries are RAW vectors of
     * quantizer->rtype with m*dsub elements (BKTIndex.cpp:463-469, QueryResultSet.h:46-60) */
    const ora_quantizer* quantizer;
    /* SearchIndexWithFilter (BKTIndex.cpp:622-647): non-NULL = one byte per vector, 0 = the filter callback
     * rejects it (it is still traversed, only never added to the results); BKT only */
    const uint8_t* filter;
} ora_index;

/* per-query counters, all int32 */
enum {
    ORA_ST_CHECKED = 0,   /* WorkSpace::m_iNumberOfCheckedLeaves at exit */
    ORA_ST_TREE_CHECKED,  /* WorkSpace::m_iNumberOfTreeCheckedLeaves (KDT only) */
    ORA_ST_NG_LEFT,       /* NGQueue.size() at exit */
    ORA_ST_SPT_LEFT,      /* SPTQueue.size() at exit */
    ORA_ST_NDIST,         /* distance evaluations (graph neighbours + tree centres/leaves) */
    ORA_ST_NEXPAND,       /* NGQueue pops that read a graph row */
    ORA_ST_NTREE,         /* tree nodes read (popped SPT cells + children scanned) */
    ORA_ST_COUNT = 8
};

float ora_max_dist(void);

float ora_distance(int32_t metric, int32_t value_type, int32_t simd_width,
                   const void* x, const void* y, int32_t dim);

void ora_distance_f32_many(int32_t metric, int32_t simd_width, const float* a, const float* b,
                           int32_t dim, int32_t n, float* out);

/* PQQuantizer::InitializeDistanceTables (PQQuantizer.h:333-348) + OPQQuantizer::m_InitMatrixTranspose */
void ora_quantizer_init(ora_quantizer* q);
/* IQuantizer::QuantizeVector(vec, out, ADC=false): n raw vectors -> n x m code bytes */
void ora_quantizer_encode(const ora_quantizer* q, const void* raw, int32_t n, uint8_t* out);
/* NeighborhoodGraph::RebuildGraph (NeighborhoodGraph.h:404-456) in its single-thread order, in place: graph is
 * [n x stride] with 2*neighborhood candidates per row; afterwards the first `neighborhood` entries of a row are its
 * neighbours.  Returns 0. */
int ora_rebuild_graph(int32_t* graph, int32_t n, int32_t stride, int32_t neighborhood);
/* IQuantizer::ReconstructVector (PQQuantizer.h:196-205, OPQQuantizer.h:124-131): n code rows -> n raw vectors of rtype */
void ora_quantizer_reconstruct(const ora_quantizer* q, const uint8_t* codes, int32_t n, void* out);
/* PQQuantizer::L2Distance with ADC off (SDC table sum, PQQuantizer.h:110-128) */
float ora_quantizer_l2(const ora_quantizer* q, const uint8_t* x, const uint8_t* y);

/* Restatement of VectorIndex::SearchIndex(batch) (VectorIndex.cpp:454-463): nq queries of `dim`
 * elements of the index value type, k results each; ids/dists are [nq*k]; stats is [nq*ORA_ST_COUNT]
 * or NULL.  threads <= 0 -> OpenMP default.  Returns 0 (ErrorCode::Success). */
int ora_search_batch(const ora_index* idx, const void* queries, int32_t nq, int32_t k,
                     int32_t* ids, float* dists, int32_t* stats, int32_t threads);

/* Restatement of one NeighborhoodGraph::RefineNode pass (NeighborhoodGraph.h:534-545 + BKTIndex.cpp:698-711 +
 * RelativeNeighborhoodGraph.h:20-38) over nodes [first_node, first_node+num_nodes) against the index's current
 * graph; out_graph is [num_nodes*neighborhood]; res_ids/res_dists are NULL or [num_nodes*(cef+1)]. */
int ora_refine_nodes(const ora_index* idx, int32_t first_node, int32_t num_nodes, int32_t cef, int32_t neighborhood,
                     float rng_factor, int32_t* out_graph, int32_t* res_ids, float* res_dists, int32_t threads);

/* Restatement of VectorIndex::GetIterator / ResultIterator::Next / Close (ResultIterator.cpp; BKTIndex.cpp:354-427,
 * :650-696): one resumable search.  BKT without quantizer only (the reference's KDT has no iterator). */
typedef struct ora_iterator ora_iterator;
ora_iterator* ora_iter_open(const ora_index* idx, const void* query);
int ora_iter_next(ora_iterator* it, int32_t batch, int32_t* ids, float* dists, int32_t* relaxed_mono);
void ora_iter_close(ora_iterator* it);
/* BKT::Index<T>::SearchIndexIterativeFromNeareast (BKTIndex.cpp:543-595) on an iterator opened with ora_iter_open:
 * first call = the k nearest by a full search + re-seeding from their neighbours, later calls = the next k. */
int ora_iter_next_from_nearest(ora_iterator* it, int32_t k, int32_t* ids, float* dists);

#ifdef __cplusplus
}
#endif
#endif
