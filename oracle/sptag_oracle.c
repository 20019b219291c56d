/* oracle/sptag_oracle.c
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C CPU restatement of the reference's batched search path
 * (microsoft/SPTAG @ /root/reference, paths below relative to AnnService/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library, and there only as the checker / CPU baseline -- never as the product path.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_pin.py) against
 *   (1) the reference's own known-answer tests (Test/src/AlgoTest.cpp:163-201,
 *       Test/cuda/distance_tests.cu:15-17, Test/src/DistanceTest.cpp:36-50) and
 *   (2) outputs of the UNMODIFIED reference compiled here as oracle/_ref/libsptag_ref.so
 *       (ids, distances and the WorkSpace counters, bit for bit, on the same index files).
 *
 * What is restated, with the reference lines each function follows:
 *   distance          src/Core/Common/DistanceUtils.cpp:297-303 (REPEAT), :650-682 (L2 AVX512),
 *                     :1016-1046 (cosine AVX512), :616-648/:982-1014 (AVX), :590-614/:958-980 (SSE),
 *                     inc/Core/Common/DistanceUtils.h:25-79 (scalar templates)
 *   heap              inc/Core/Common/Heap.h:13-106
 *   m_Results         inc/Core/Common/WorkSpace.h:167-225 (DistPriorityQueue)
 *   visited set       inc/Core/Common/WorkSpace.h:43-165 (OptHashPosVector) -- semantically an
 *                     exact set (two tables, then DoubleSize()); restated as a byte map over N
 *   work space        inc/Core/Common/WorkSpace.h:230-320
 *   top-K             inc/Core/Common/QueryResultSet.h:17-120
 *   BKT seed lookup   inc/Core/Common/BKTree.h:696-769 (InitSearchTrees, m_bfs == 0), :771-799 (SearchTrees)
 *   BKT graph search  src/Core/BKT/BKTIndex.cpp:268-352 (Search), :463-508 (dispatch), :595-620
 *   KDT seed lookup   inc/Core/Common/KDTree.h:213-271
 *   KDT graph search  src/Core/KDT/KDTIndex.cpp:182-241, :268-318
 *   batch loop        src/Core/VectorIndex.cpp:454-463
 */
#include "sptag_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Common.h:122  MaxDist = numeric_limits<float>::max() / 10 */
static float kMaxDist(void) { return FLT_MAX / 10; }
float ora_max_dist(void) { return kMaxDist(); }

/* ------------------------------------------------------------------------------------------ */
/* distance                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* Compiled with -ffp-contract=off: every mul/add/sub below rounds separately, which is what the
 * reference's intrinsics do in its g++ -O3 build (vsubps/vmulps/vaddps, no FMA); the reference's
 * plain-C scalar tails ARE contracted by g++ on an FMA target, hence the explicit fmaf() there
 * (re-verified by tests/test_oracle_pin.py). */

static inline float term_f32(int cosine, float x, float y)
{
    if (cosine) return x * y;
    float d = x - y;
    return d * d;
}

static inline float tail_f32(int cosine, float x, float y, float acc)
{
    if (cosine) return fmaf(x, y, acc);
    float d = x - y;
    return fmaf(d, d, acc);
}

static float dist_f32(int cosine, int width, const float* x, const float* y, int len)
{
    int i = 0, j;
    float diff;
    if (width == 16) {
        /* ComputeL2Distance_AVX512 / ComputeCosineDistance_AVX512 (float) */
        float a16[16], a8[8], a4[4];
        for (j = 0; j < 16; j++) a16[j] = 0.0f;
        for (; i + 16 <= len; i += 16)
            for (j = 0; j < 16; j++) a16[j] = a16[j] + term_f32(cosine, x[i + j], y[i + j]);
        for (j = 0; j < 8; j++) a8[j] = a16[j] + a16[j + 8];
        for (; i + 8 <= len; i += 8)
            for (j = 0; j < 8; j++) a8[j] = a8[j] + term_f32(cosine, x[i + j], y[i + j]);
        for (j = 0; j < 4; j++) a4[j] = a8[j] + a8[j + 4];
        for (; i + 4 <= len; i += 4)
            for (j = 0; j < 4; j++) a4[j] = a4[j] + term_f32(cosine, x[i + j], y[i + j]);
        diff = a4[0] + a4[1] + a4[2] + a4[3];
    } else if (width == 8) {
        /* ComputeL2Distance_AVX / ComputeCosineDistance_AVX (float): 16 per trip = two 8-wide adds */
        float a8[8], a4[4];
        for (j = 0; j < 8; j++) a8[j] = 0.0f;
        for (; i + 16 <= len; i += 16) {
            for (j = 0; j < 8; j++) a8[j] = a8[j] + term_f32(cosine, x[i + j], y[i + j]);
            for (j = 0; j < 8; j++) a8[j] = a8[j] + term_f32(cosine, x[i + 8 + j], y[i + 8 + j]);
        }
        for (j = 0; j < 4; j++) a4[j] = a8[j] + a8[j + 4];
        for (; i + 4 <= len; i += 4)
            for (j = 0; j < 4; j++) a4[j] = a4[j] + term_f32(cosine, x[i + j], y[i + j]);
        diff = a4[0] + a4[1] + a4[2] + a4[3];
    } else if (width == 4) {
        /* ComputeL2Distance_SSE / ComputeCosineDistance_SSE (float) */
        float a4[4];
        for (j = 0; j < 4; j++) a4[j] = 0.0f;
        for (; i + 4 <= len; i += 4)
            for (j = 0; j < 4; j++) a4[j] = a4[j] + term_f32(cosine, x[i + j], y[i + j]);
        diff = a4[0] + a4[1] + a4[2] + a4[3];
    } else {
        /* DistanceUtils.h:25-44 / :60-79 scalar templates: one accumulator.  g++ -O3 SLP-vectorises the
         * 4-unrolled body (vsubps/vmulps, then four separately rounded vaddss) and FMA-contracts only
         * the scalar remainder loop (checked against the compiled reference in tests/test_oracle_pin.py) */
        diff = 0.0f;
        for (; i + 4 <= len; i += 4)
            for (j = 0; j < 4; j++) diff = diff + term_f32(cosine, x[i + j], y[i + j]);
    }
    for (; i < len; i++) diff = tail_f32(cosine, x[i], y[i], diff);
    /* float base is 1 (CommonUtils.h GetBase<float>) */
    return cosine ? 1 - diff : diff;
}

/* int8 / uint8 variants (DistanceUtils.cpp:305-558 L2, :684-874 cosine; helpers :17-296).
 * One SIMD step over 4*W bytes yields W float lanes; lane t = (128-bit lane L = t/4, position p = t%4) holds the
 * EXACT int32 sum of the four terms at byte offsets 16L + 2p + {0, 1, 8, 9} (unpacklo/unpackhi_epi8 interleave +
 * madd_epi16 + add_epi32), converted with cvtepi32_ps and accumulated in fp32 like the float kernels. */
static inline int elem_i8u8(const void* p, int is_unsigned, int i)
{
    return is_unsigned ? (int)((const uint8_t*)p)[i] : (int)((const int8_t*)p)[i];
}

static inline float lane_term_i8(int cosine, int is_unsigned, const void* x, const void* y, int base, int t)
{
    static const int add[4] = {0, 1, 8, 9};
    const int off = base + 16 * (t / 4) + 2 * (t % 4);
    int32_t s = 0;
    for (int k = 0; k < 4; k++) {
        int a = elem_i8u8(x, is_unsigned, off + add[k]), b = elem_i8u8(y, is_unsigned, off + add[k]);
        s += cosine ? a * b : (a - b) * (a - b);
    }
    return (float)s;
}

static float dist_i8u8(int cosine, int width, int is_unsigned, const void* x, const void* y, int len)
{
    int i = 0, j;
    float diff;
    float a16[16], a8[8], a4[4];
    if (width == 16) {
        for (j = 0; j < 16; j++) a16[j] = 0.0f;
        for (; i + 64 <= len; i += 64)
            for (j = 0; j < 16; j++) a16[j] = a16[j] + lane_term_i8(cosine, is_unsigned, x, y, i, j);
        for (j = 0; j < 8; j++) a8[j] = a16[j] + a16[j + 8];
        for (; i + 32 <= len; i += 32)
            for (j = 0; j < 8; j++) a8[j] = a8[j] + lane_term_i8(cosine, is_unsigned, x, y, i, j);
        for (j = 0; j < 4; j++) a4[j] = a8[j] + a8[j + 4];
        for (; i + 16 <= len; i += 16)
            for (j = 0; j < 4; j++) a4[j] = a4[j] + lane_term_i8(cosine, is_unsigned, x, y, i, j);
        diff = a4[0] + a4[1] + a4[2] + a4[3];
    } else if (width == 8) {
        for (j = 0; j < 8; j++) a8[j] = 0.0f;
        for (; i + 32 <= len; i += 32)
            for (j = 0; j < 8; j++) a8[j] = a8[j] + lane_term_i8(cosine, is_unsigned, x, y, i, j);
        for (j = 0; j < 4; j++) a4[j] = a8[j] + a8[j + 4];
        for (; i + 16 <= len; i += 16)
            for (j = 0; j < 4; j++) a4[j] = a4[j] + lane_term_i8(cosine, is_unsigned, x, y, i, j);
        diff = a4[0] + a4[1] + a4[2] + a4[3];
    } else if (width == 4) {
        for (j = 0; j < 4; j++) a4[j] = 0.0f;
        for (; i + 16 <= len; i += 16)
            for (j = 0; j < 4; j++) a4[j] = a4[j] + lane_term_i8(cosine, is_unsigned, x, y, i, j);
        diff = a4[0] + a4[1] + a4[2] + a4[3];
    } else {
        diff = 0.0f;
    }
    /* plain-C tails on (float) casts; every value here is an integer far below 2^24 for the supported
     * dimensions, so the 4-unrolled / FMA-contracted distinction of the float kernels cannot change a bit */
    for (; i + 4 <= len; i += 4)
        for (j = 0; j < 4; j++)
            diff = diff + term_f32(cosine, (float)elem_i8u8(x, is_unsigned, i + j), (float)elem_i8u8(y, is_unsigned, i + j));
    for (; i < len; i++)
        diff = tail_f32(cosine, (float)elem_i8u8(x, is_unsigned, i), (float)elem_i8u8(y, is_unsigned, i), diff);
    /* base^2 - dot: 127^2 = 16129 (int8), 255^2 = 65025 (uint8) (DistanceUtils.cpp:775, :869) */
    return cosine ? (float)(is_unsigned ? 65025 : 16129) - diff : diff;
}

/* int16 variants, AVX-512 build only (DistanceUtils.cpp:559-596 L2, :930-967 cosine; helpers :263-289).
 * One 512-bit step covers 32 elements and yields 16 float lanes, t = 4L + p (128-bit lane L, position p):
 *   cosine: _mm512_mul_epi16 = cvtepi32_ps(madd_epi16): lane t = (float)(int32)(x[2t]*y[2t] + x[2t+1]*y[2t+1]);
 *   L2:     _mm512_sqdf_epi16: unpacklo/hi_epi16 sign-extend, dlo = (float)(x[8L+p]-y[8L+p]),
 *           dhi = (float)(x[8L+4+p]-y[8L+4+p]); g++ -O3 contracts add_ps(mul_ps(dlo,dlo), mul_ps(dhi,dhi)) into
 *           fma(dhi, dhi, dlo*dlo) (vmulps + vfmadd132ps in the compiled reference), the accumulation stays a vaddps.
 * 256-/128-bit steps do the same on 8 / 4 lanes.  Plain-C tails: the 4-unrolled statements are FMA-contracted
 * (vfmadd*ss), the single-element remainder loops are NOT (vmulss + vaddss); all checked against the compiled
 * reference in tests/test_oracle_pin.py -- squares of int16 differences exceed 2^24, so every rounding shows. */
static inline float lane_term_i16(int cosine, const int16_t* x, const int16_t* y, int base, int t)
{
    if (cosine) {
        int32_t s = (int32_t)((uint32_t)((int32_t)x[base + 2 * t] * y[base + 2 * t]) +
                              (uint32_t)((int32_t)x[base + 2 * t + 1] * y[base + 2 * t + 1]));
        return (float)s;
    }
    const int off = base + 8 * (t / 4) + (t % 4);
    const float dlo = (float)((int32_t)x[off] - (int32_t)y[off]);
    const float dhi = (float)((int32_t)x[off + 4] - (int32_t)y[off + 4]);
    return fmaf(dhi, dhi, dlo * dlo);
}

static float dist_i16(int cosine, int width, const int16_t* x, const int16_t* y, int len)
{
    int i = 0, j;
    float diff;
    float a16[16], a8[8], a4[4];
    if (width != 16) return NAN; /* the AVX / SSE int16 variants are not restated */
    for (j = 0; j < 16; j++) a16[j] = 0.0f;
    for (; i + 32 <= len; i += 32)
        for (j = 0; j < 16; j++) a16[j] = a16[j] + lane_term_i16(cosine, x, y, i, j);
    for (j = 0; j < 8; j++) a8[j] = a16[j] + a16[j + 8];
    for (; i + 16 <= len; i += 16)
        for (j = 0; j < 8; j++) a8[j] = a8[j] + lane_term_i16(cosine, x, y, i, j);
    for (j = 0; j < 4; j++) a4[j] = a8[j] + a8[j + 4];
    for (; i + 8 <= len; i += 8)
        for (j = 0; j < 4; j++) a4[j] = a4[j] + lane_term_i16(cosine, x, y, i, j);
    diff = a4[0] + a4[1] + a4[2] + a4[3];
    for (; i + 4 <= len; i += 4)
        for (j = 0; j < 4; j++) diff = tail_f32(cosine, (float)x[i + j], (float)y[i + j], diff);
    for (; i < len; i++) {
        if (cosine) {
            float c = (float)x[i] * (float)y[i];
            diff = diff + c;
        } else {
            float c = (float)x[i] - (float)y[i];
            c = c * c;
            diff = diff + c;
        }
    }
    /* base^2 - dot with the INTEGER literal 1073676289 = 32767^2 converted to float (DistanceUtils.cpp:966) */
    return cosine ? (float)1073676289 - diff : diff;
}

float ora_distance(int32_t metric, int32_t value_type, int32_t simd_width,
                   const void* x, const void* y, int32_t dim)
{
    int cosine = (metric != ORA_L2);
    if (value_type == ORA_FLOAT)
        return dist_f32(cosine, simd_width, (const float*)x, (const float*)y, dim);
    if (value_type == ORA_INT8 || value_type == ORA_UINT8)
        return dist_i8u8(cosine, simd_width, value_type == ORA_UINT8, x, y, dim);
    if (value_type == ORA_INT16) return dist_i16(cosine, simd_width, (const int16_t*)x, (const int16_t*)y, dim);
    return NAN;
}

void ora_distance_f32_many(int32_t metric, int32_t simd_width, const float* a, const float* b,
                           int32_t dim, int32_t n, float* out)
{
    int cosine = (metric != ORA_L2);
    for (int32_t i = 0; i < n; i++)
        out[i] = dist_f32(cosine, simd_width, a + (size_t)i * dim, b + (size_t)i * dim, dim);
}

/* ------------------------------------------------------------------------------------------ */
/* PQ / OPQ quantizer  (PQQuantizer.h:110-180, :333-348; OPQQuantizer.h:96-121, :198-206)      */
/* ------------------------------------------------------------------------------------------ */

void ora_quantizer_init(ora_quantizer* q)
{
    /* InitializeDistanceTables: table[i][j][k] = L2(codebook[i][j], codebook[i][k]) via the selected
     * DistanceUtils float variant */
    const int m = q->m, ks = q->ks, d = q->dsub;
    for (int i = 0; i < m; i++) {
        const float* base = q->codebooks + (size_t)i * ks * d;
        for (int j = 0; j < ks; j++)
            for (int k = 0; k < ks; k++)
                q->sdc[((size_t)i * ks + j) * ks + k] = dist_f32(0, q->simd_width, base + (size_t)j * d, base + (size_t)k * d, d);
    }
    if (q->qtype == ORA_Q_OPQ) {
        const int dim = m * d;
        for (int i = 0; i < dim; i++)
            for (int j = 0; j < dim; j++) q->rotation_t[(size_t)i * dim + j] = q->rotation[(size_t)j * dim + i];
    }
}

static float raw_elem(const void* raw, int rtype, size_t i)
{
    switch (rtype) {
    case ORA_INT8: return (float)((const int8_t*)raw)[i];
    case ORA_UINT8: return (float)((const uint8_t*)raw)[i];
    case ORA_INT16: return (float)((const int16_t*)raw)[i];
    default: return ((const float*)raw)[i];
    }
}

static void quantize_one(const ora_quantizer* q, const void* raw, uint8_t* out, float* tmp /* 2*dim */)
{
    const int m = q->m, ks = q->ks, d = q->dsub, dim = m * d;
    float* vec = tmp;
    float* rot = tmp + dim;
    for (int i = 0; i < dim; i++) vec[i] = raw_elem(raw, q->rtype, (size_t)i);
    const float* src = vec;
    if (q->qtype == ORA_Q_OPQ) {
        /* m_VectorMatrixMultiply(m_OPQMatrix_T, vec, out): out[i] = m_base - m_fdot(vec, row_i), m_base = 1,
         * m_fdot = float cosine distance = 1 - dot (OPQQuantizer.h:198-206) */
        for (int i = 0; i < dim; i++)
            rot[i] = 1 - dist_f32(1, q->simd_width, vec, q->rotation_t + (size_t)i * dim, dim);
        src = rot;
    }
    /* PQQuantizer::QuantizeVector, ADC off: first codeword with the strictly smallest L2 distance */
    for (int i = 0; i < m; i++) {
        int best = -1;
        float mind = INFINITY;
        const float* cb = q->codebooks + (size_t)i * ks * d;
        for (int j = 0; j < ks; j++) {
            float dist = dist_f32(0, q->simd_width, src + (size_t)i * d, cb + (size_t)j * d, d);
            if (dist < mind) {
                best = j;
                mind = dist;
            }
        }
        out[i] = (uint8_t)best;
    }
}

/* PQQuantizer::QuantizeVector with ADC enabled (PQQuantizer.h:141-157): table[i][j] = L2(subvector_i, codeword_ij);
 * OPQ rotates first (OPQQuantizer.h:96-121) */
static void adc_table_one(const ora_quantizer* q, const void* raw, float* table, float* tmp /* 2*dim */)
{
    const int m = q->m, ks = q->ks, d = q->dsub, dim = m * d;
    float* vec = tmp;
    float* rot = tmp + dim;
    for (int i = 0; i < dim; i++) vec[i] = raw_elem(raw, q->rtype, (size_t)i);
    const float* src = vec;
    if (q->qtype == ORA_Q_OPQ) {
        for (int i = 0; i < dim; i++)
            rot[i] = 1 - dist_f32(1, q->simd_width, vec, q->rotation_t + (size_t)i * dim, dim);
        src = rot;
    }
    for (int i = 0; i < m; i++)
        for (int j = 0; j < ks; j++)
            table[(size_t)i * ks + j] = dist_f32(0, q->simd_width, src + (size_t)i * d,
                                                 q->codebooks + ((size_t)i * ks + j) * d, d);
}

/* PQQuantizer::L2Distance, ADC branch (PQQuantizer.h:114-119): pX is the query's table */
static float adc_l2(const ora_quantizer* q, const float* table, const uint8_t* y)
{
    float out = 0;
    for (int i = 0; i < q->m; i++) out += table[(size_t)i * q->ks + y[i]];
    return out;
}

void ora_quantizer_encode(const ora_quantizer* q, const void* raw, int32_t n, uint8_t* out)
{
    static const size_t elem[4] = {1, 1, 2, 4};
    const int dim = q->m * q->dsub;
    const size_t rs = elem[q->rtype] * (size_t)dim;
#pragma omp parallel
    {
        float* tmp = (float*)malloc(sizeof(float) * 2 * (size_t)dim);
#pragma omp for
        for (int32_t i = 0; i < n; i++)
            quantize_one(q, (const char*)raw + (size_t)i * rs, out + (size_t)i * q->m, tmp);
        free(tmp);
    }
}

/* IQuantizer::ReconstructVector for one code row.
 * PQQuantizer<T>::ReconstructVector (PQQuantizer.h:196-205): the codewords, copied.
 * OPQQuantizer<T>::ReconstructVector (OPQQuantizer.h:124-131): the float codewords, then
 * m_VectorMatrixMultiply<T>(m_OPQMatrix, pre, out): out[i] = (T)(m_base - m_fdot(pre, row_i of m_OPQMatrix)), m_base = 1,
 * m_fdot = float cosine distance (OPQQuantizer.h:198-206); the (T) cast is C's truncation toward zero. */
static void reconstruct_one(const ora_quantizer* q, const uint8_t* code, void* out, float* tmp /* dim */)
{
    const int m = q->m, ks = q->ks, d = q->dsub, dim = m * d;
    float* pre = (q->qtype == ORA_Q_OPQ) ? tmp : (float*)out; /* PQ is restated for float codebooks only */
    for (int i = 0; i < m; i++)
        memcpy(pre + (size_t)i * d, q->codebooks + ((size_t)i * ks + code[i]) * d, sizeof(float) * (size_t)d);
    if (q->qtype != ORA_Q_OPQ) return;
    for (int i = 0; i < dim; i++) {
        const float v = 1 - dist_f32(1, q->simd_width, pre, q->rotation + (size_t)i * dim, dim);
        switch (q->rtype) {
        case ORA_INT8: ((int8_t*)out)[i] = (int8_t)(int32_t)v; break;
        case ORA_UINT8: ((uint8_t*)out)[i] = (uint8_t)(int32_t)v; break;
        case ORA_INT16: ((int16_t*)out)[i] = (int16_t)(int32_t)v; break;
        default: ((float*)out)[i] = v; break;
        }
    }
}

void ora_quantizer_reconstruct(const ora_quantizer* q, const uint8_t* codes, int32_t n, void* out)
{
    static const size_t elem[4] = {1, 1, 2, 4};
    const int dim = q->m * q->dsub;
    const size_t rs = elem[q->rtype] * (size_t)dim;
    float* tmp = (float*)malloc(sizeof(float) * (size_t)dim);
    for (int32_t i = 0; i < n; i++) reconstruct_one(q, codes + (size_t)i * q->m, (char*)out + (size_t)i * rs, tmp);
    free(tmp);
}

float ora_quantizer_l2(const ora_quantizer* q, const uint8_t* x, const uint8_t* y)
{
    float out = 0;
    for (int i = 0; i < q->m; i++) out += q->sdc[((size_t)i * q->ks + x[i]) * q->ks + y[i]];
    return out;
}

/* ------------------------------------------------------------------------------------------ */
/* NeighborhoodGraph::RebuildGraph (NeighborhoodGraph.h:404-456), the in-degree repair that BuildGraph runs after its
 * refine passes when EnableRebuild is set: rows hold 2 x neighborhood candidates (row stride `stride` >= that); the first
 * neighborhood/2 stay, the other neighborhood/2 slots are refilled from entries [neighborhood/2, 2 x neighborhood) --
 * first the ones whose target has an in-degree below neighborhood/2, then the earliest others -- in index order, the
 * in-degree array following every change.  The reference runs the node loop under OpenMP without synchronising the
 * in-degree array (its result depends on thread timing); this is its single-thread order, node by node.            */
/* ------------------------------------------------------------------------------------------ */
int ora_rebuild_graph(int32_t* graph, int32_t n, int32_t stride, int32_t neighborhood)
{
    if (n < 0 || neighborhood < 2 || stride < 2 * neighborhood) return 1;
    int32_t* indegree = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    uint8_t* reserve = (uint8_t*)malloc((size_t)2 * neighborhood);
    for (int32_t i = 0; i < n; i++) {
        const int32_t* outnodes = graph + (size_t)i * stride;
        for (int32_t j = 0; j < neighborhood; j++)
            if (outnodes[j] >= 0) indegree[outnodes[j]]++;
    }
    const int rebuild_threshold = neighborhood / 2;
    const int rebuildstart = neighborhood / 2;
    for (int32_t i = 0; i < n; i++) {
        int32_t* outnodes = graph + (size_t)i * stride;
        memset(reserve, 0, (size_t)2 * neighborhood);
        int total = 0;
        for (int32_t j = rebuildstart; j < neighborhood * 2; j++)
            if (outnodes[j] >= 0 && indegree[outnodes[j]] < rebuild_threshold) {
                reserve[j] = 1;
                total++;
            }
        for (int32_t j = rebuildstart; j < neighborhood * 2 && total < neighborhood - rebuildstart; j++) {
            if (!reserve[j]) {
                reserve[j] = 1;
                total++;
            }
        }
        for (int32_t j = rebuildstart, z = rebuildstart; j < neighborhood; j++) {
            while (!reserve[z]) z++;
            if (outnodes[j] >= 0) indegree[outnodes[j]] = indegree[outnodes[j]] - 1;
            if (outnodes[z] >= 0) indegree[outnodes[z]] = indegree[outnodes[z]] + 1;
            outnodes[j] = outnodes[z];
            z++;
        }
    }
    free(reserve);
    free(indegree);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Heap<NodeDistPair>  (Heap.h:13-106, SearchResult.h:11-27)                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t node;
    float distance;
} pair_t;

typedef struct {
    pair_t* heap; /* 1-based; heap[0] stays the default (-1, MaxDist) */
    int length, count, lastlevel;
} heap_t;

static void heap_resize(heap_t* h, int size)
{
    free(h->heap);
    h->length = size;
    h->heap = (pair_t*)malloc(sizeof(pair_t) * ((size_t)size + 1));
    for (int i = 0; i <= size; i++) {
        h->heap[i].node = -1;
        h->heap[i].distance = kMaxDist();
    }
    h->count = 0;
    h->lastlevel = (int)pow(2.0, floor(log2((float)size)));
}

static void heap_clear(heap_t* h, int size)
{
    if (size > h->length) heap_resize(h, size);
    h->count = 0;
}

static inline pair_t* heap_top(heap_t* h) { return h->count == 0 ? &h->heap[0] : &h->heap[1]; }

static void heap_insert(heap_t* h, pair_t value)
{
    int loc;
    pair_t* a = h->heap;
    if (h->count == h->length) {
        int maxi = h->lastlevel;
        for (int i = h->lastlevel + 1; i <= h->length; i++)
            if (a[maxi].distance < a[i].distance) maxi = i;
        if (value.distance > a[maxi].distance) return;
        loc = maxi;
    } else {
        loc = ++(h->count);
    }
    int par = (loc >> 1);
    while (par > 0 && value.distance < a[par].distance) {
        a[loc] = a[par];
        loc = par;
        par >>= 1;
    }
    a[loc] = value;
}

static inline void pair_swap(pair_t* x, pair_t* y)
{
    pair_t t = *x;
    *x = *y;
    *y = t;
}

static void heap_heapify(heap_t* h)
{
    pair_t* a = h->heap;
    int parent = 1, next = 2;
    while (next < h->count) {
        if (a[next].distance > a[next + 1].distance) next++;
        if (a[next].distance < a[parent].distance) {
            pair_swap(&a[parent], &a[next]);
            parent = next;
            next <<= 1;
        } else
            break;
    }
    if (next == h->count && a[next].distance < a[parent].distance) pair_swap(&a[parent], &a[next]);
}

/* T& pop(): returns the slot the old root was swapped into */
static pair_t heap_pop(heap_t* h)
{
    if (h->count == 0) return h->heap[0];
    pair_swap(&h->heap[1], &h->heap[h->count]);
    h->count--;
    heap_heapify(h);
    return h->heap[h->count + 1];
}

/* ------------------------------------------------------------------------------------------ */
/* DistPriorityQueue  (WorkSpace.h:167-225)                                                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int size, length, count;
    float* data;
} dpq_t;

static void dpq_clear(dpq_t* q, int count_)
{
    if (count_ > q->size) {
        q->size = count_;
        free(q->data);
        q->data = (float*)malloc(sizeof(float) * ((size_t)count_ + 1));
    }
    q->data[1] = kMaxDist();
    q->length = 1;
    q->count = count_;
}

static int dpq_insert(dpq_t* q, float dist)
{
    float* d = q->data;
    if (dist > d[1]) return 0;
    if (q->length == q->count) {
        d[1] = dist;
        int parent = 1, next = 2;
        while (next < q->length) {
            if (d[next] < d[next + 1]) next++;
            if (d[next] > d[parent]) {
                float t = d[parent];
                d[parent] = d[next];
                d[next] = t;
                parent = next;
                next <<= 1;
            } else
                break;
        }
        if (next == q->length && d[next] > d[parent]) {
            float t = d[parent];
            d[parent] = d[next];
            d[next] = t;
        }
    } else {
        int next = ++(q->length), parent = (next >> 1);
        while (parent > 0 && dist > d[parent]) {
            d[next] = d[parent];
            next = parent;
            parent >>= 1;
        }
        d[next] = dist;
    }
    return 1;
}

static inline float dpq_worst(const dpq_t* q) { return q->data[1]; }

/* ------------------------------------------------------------------------------------------ */
/* QueryResultSet  (QueryResultSet.h:17-120)                                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t vid;
    float dist;
} res_t;

static inline int res_less(res_t a, res_t b)
{
    return (a.dist < b.dist) || ((a.dist == b.dist) && (a.vid < b.vid));
}

static void res_heapify(res_t* r, int count)
{
    int parent = 0, next = 1, maxidx = count - 1;
    while (next < maxidx) {
        if (res_less(r[next], r[next + 1])) next++;
        if (res_less(r[parent], r[next])) {
            res_t t = r[next];
            r[next] = r[parent];
            r[parent] = t;
            parent = next;
            next = (parent << 1) + 1;
        } else
            break;
    }
    if (next == maxidx && res_less(r[parent], r[next])) {
        res_t t = r[parent];
        r[parent] = r[next];
        r[next] = t;
    }
}

static int res_add_point(res_t* r, int k, int32_t index, float dist)
{
    if (dist < r[0].dist || (dist == r[0].dist && index < r[0].vid)) {
        r[0].vid = index;
        r[0].dist = dist;
        res_heapify(r, k);
        return 1;
    }
    return 0;
}

static void res_sort(res_t* r, int k)
{
    for (int i = k - 1; i >= 0; i--) {
        res_t t = r[0];
        r[0] = r[i];
        r[i] = t;
        res_heapify(r, i);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* WorkSpace  (WorkSpace.h:230-320)                                                            */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint8_t* visited; /* exact set over [0, n) standing in for OptHashPosVector */
    int32_t n;
    heap_t ng, spt;
    dpq_t results;
    int no_better, tree_checked, checked, max_check;
    /* our own accounting (algorithmic bytes) */
    int ndist, nexpand, ntree;
} ws_t;

static void ws_init(ws_t* ws, int32_t n, int max_check_alloc)
{
    memset(ws, 0, sizeof(*ws));
    ws->n = n;
    ws->visited = (uint8_t*)calloc((size_t)n + 1, 1);
    heap_resize(&ws->spt, max_check_alloc * 10);
    heap_resize(&ws->ng, max_check_alloc * 30);
    dpq_clear(&ws->results, max_check_alloc / 16 > 1 ? max_check_alloc / 16 : 1);
}

static void ws_reset(ws_t* ws, int max_check, int result_num)
{
    memset(ws->visited, 0, (size_t)ws->n + 1);
    heap_clear(&ws->spt, max_check * 10);
    heap_clear(&ws->ng, max_check * 30);
    dpq_clear(&ws->results, max_check / 16 > result_num ? max_check / 16 : result_num);
    ws->no_better = 0;
    ws->tree_checked = 0;
    ws->checked = 0;
    ws->max_check = max_check;
    ws->ndist = ws->nexpand = ws->ntree = 0;
}

static void ws_free(ws_t* ws)
{
    free(ws->visited);
    free(ws->ng.heap);
    free(ws->spt.heap);
    free(ws->results.data);
}

/* returns nonzero if idx was already present (OptHashPosVector::CheckAndSet, WorkSpace.h:113-117) */
static inline int ws_check_and_set(ws_t* ws, int32_t idx)
{
    if (ws->visited[idx]) return 1;
    ws->visited[idx] = 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* search                                                                                     */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const ora_index* idx;
    const void* query;
    size_t row_bytes;
    int cosine;
} qctx_t;

static inline float qdist(const qctx_t* c, ws_t* ws, int32_t id)
{
    ws->ndist++;
    const char* row = (const char*)c->idx->vectors + (size_t)id * c->row_bytes;
    if (c->idx->quantizer) { /* m_fComputeDistance = quantizer L2Distance (BKTIndex.cpp:34-50, IQuantizer.cpp:103-114) */
        if (c->idx->quantizer->enable_adc) return adc_l2(c->idx->quantizer, (const float*)c->query, (const uint8_t*)row);
        return ora_quantizer_l2(c->idx->quantizer, (const uint8_t*)c->query, (const uint8_t*)row);
    }
    return ora_distance(c->idx->metric, c->idx->value_type, c->idx->simd_width, c->query, row, c->idx->dim);
}

/* StaticDispatch::CheckFilter / AlwaysTrue (BKTIndex.cpp:455-458, :471-507) */
static inline int check_filter(const ora_index* idx, int32_t id) { return idx->filter == NULL || idx->filter[id] != 0; }

static inline int not_deleted(const ora_index* idx, int32_t id)
{
    /* StaticDispatch::CheckIfNotDeleted / AlwaysTrue (BKTIndex.cpp:437-440, :471-507) */
    if (idx->deleted == NULL || idx->num_deleted == 0) return 1;
    return idx->deleted[id] != 1;
}

/* BKTree::InitSearchTrees (BKTree.h:696-769), default m_bfs == 0 path */
static void bkt_init_search_trees(const qctx_t* c, ws_t* ws)
{
    const ora_index* idx = c->idx;
    const ora_bkt_node* nodes = (const ora_bkt_node*)idx->nodes;
    for (int i = 0; i < idx->tree_num; i++) {
        int32_t start = idx->tree_starts[i];
        const ora_bkt_node* node = &nodes[start];
        ws->ntree++;
        if (node->childStart < 0) {
            pair_t p = {start, qdist(c, ws, node->centerid)};
            heap_insert(&ws->spt, p);
        } else {
            for (int32_t begin = node->childStart; begin < node->childEnd; begin++) {
                ws->ntree++;
                pair_t p = {begin, qdist(c, ws, nodes[begin].centerid)};
                heap_insert(&ws->spt, p);
            }
        }
    }
}

/* BKTree::SearchTrees (BKTree.h:771-799) */
static void bkt_search_trees(const qctx_t* c, ws_t* ws, int limits)
{
    const ora_bkt_node* nodes = (const ora_bkt_node*)c->idx->nodes;
    while (ws->spt.count != 0) {
        pair_t bcell = heap_pop(&ws->spt);
        const ora_bkt_node* tnode = &nodes[bcell.node];
        ws->ntree++;
        if (tnode->childStart < 0) {
            if (!ws_check_and_set(ws, tnode->centerid)) {
                ws->checked++;
                pair_t p = {tnode->centerid, bcell.distance};
                heap_insert(&ws->ng, p);
            }
            if (ws->checked >= limits) break;
        } else {
            if (!ws_check_and_set(ws, tnode->centerid)) {
                pair_t p = {tnode->centerid, bcell.distance};
                heap_insert(&ws->ng, p);
            }
            for (int32_t begin = tnode->childStart; begin < tnode->childEnd; begin++) {
                ws->ntree++;
                pair_t p = {begin, qdist(c, ws, nodes[begin].centerid)};
                heap_insert(&ws->spt, p);
            }
        }
    }
}

/* BKT::Index<T>::Search<notDeleted, isDup, checkFilter> (BKTIndex.cpp:268-352).  never_dup = 0: CheckDup, as
 * dispatched by the public SearchIndex (searchDuplicated = true; :463-508, :595-620); never_dup = 1: NeverDup, as
 * dispatched by RefineSearchIndex (searchDuplicated = false; :698-711, StaticDispatch::NeverDup :447-452) */
static void bkt_search(const qctx_t* c, ws_t* ws, res_t* res, int k, int never_dup)
{
    const ora_index* idx = c->idx;
    const ora_bkt_node* nodes = (const ora_bkt_node*)idx->nodes;
    bkt_init_search_trees(c, ws);
    bkt_search_trees(c, ws, idx->initial_pivots);
    const int checkPos = idx->degree - 1;

    while (ws->ng.count != 0) {
        pair_t gnode = heap_pop(&ws->ng);
        int32_t tmpNode = gnode.node;
        const int32_t* node = idx->graph + (size_t)tmpNode * idx->degree;
        ws->nexpand++;

        if (gnode.distance <= res[0].dist) {
            int32_t checkNode = node[checkPos];
            if (checkNode < -1) {
                const ora_bkt_node* tnode = &nodes[-2 - checkNode];
                int32_t i = -tnode->childStart;
                do {
                    if (not_deleted(idx, tmpNode)) {
                        if (check_filter(idx, tmpNode)) {
                            /* CheckDup: stop at the first member that does not enter; NeverDup: add one, stop */
                            if (!res_add_point(res, k, tmpNode, gnode.distance) || never_dup) break;
                        }
                    }
                    if (i <= 0) break;
                    tmpNode = nodes[i].centerid;
                } while (i++ < tnode->childEnd);
            } else {
                if (not_deleted(idx, tmpNode)) {
                    if (check_filter(idx, tmpNode)) res_add_point(res, k, tmpNode, gnode.distance);
                }
            }
        } else {
            if (not_deleted(idx, tmpNode)) {
                if (gnode.distance > dpq_worst(&ws->results) || ws->checked > ws->max_check) {
                    res_sort(res, k);
                    return;
                }
            }
        }
        for (int i = 0; i <= checkPos; i++) {
            int32_t nn_index = node[i];
            if (nn_index < 0) break;
            if (ws_check_and_set(ws, nn_index)) continue;
            float distance2leaf = qdist(c, ws, nn_index);
            ws->checked++;
            if (dpq_insert(&ws->results, distance2leaf)) {
                pair_t p = {nn_index, distance2leaf};
                heap_insert(&ws->ng, p);
            }
        }
        if (heap_top(&ws->ng)->distance > heap_top(&ws->spt)->distance) {
            bkt_search_trees(c, ws, idx->other_pivots + ws->checked);
        }
    }
    res_sort(res, k);
}

/* KDTree::KDTSearch (KDTree.h:233-271); the recursion is a tail call, restated as a loop */
static void kdt_search_node(const qctx_t* c, ws_t* ws, int32_t node, float distBound)
{
    const ora_index* idx = c->idx;
    const ora_kdt_node* nodes = (const ora_kdt_node*)idx->nodes;
    for (;;) {
        if (node < 0) {
            int32_t index = -node - 1;
            if (index >= idx->n) return;
            if (ws_check_and_set(ws, index)) return;
            ++ws->tree_checked;
            ++ws->checked;
            pair_t p = {index, qdist(c, ws, index)};
            heap_insert(&ws->ng, p);
            return;
        }
        const ora_kdt_node* tnode = &nodes[node];
        ws->ntree++;
        /* split test reads the raw (un-quantized) query, KDTree.h:255 */
        float diff = raw_elem(c->query, c->idx->value_type, (size_t)tnode->split_dim) - tnode->split_value;
        /* `distBound + diff * diff` is FMA-contracted in the reference's g++ -O3 build on an FMA
         * target (one vfmadd in KDTree::KDTSearch, checked by disassembly) */
        float distanceBound = fmaf(diff, diff, distBound);
        int32_t otherChild, bestChild;
        if (diff < 0) {
            bestChild = tnode->left;
            otherChild = tnode->right;
        } else {
            otherChild = tnode->left;
            bestChild = tnode->right;
        }
        pair_t p = {otherChild, distanceBound};
        heap_insert(&ws->spt, p);
        node = bestChild;
    }
}

static void kdt_search_trees(const qctx_t* c, ws_t* ws, int limits)
{
    while (ws->spt.count != 0 && ws->checked < limits) {
        pair_t tcell = heap_pop(&ws->spt);
        kdt_search_node(c, ws, tcell.node, tcell.distance);
    }
}

/* KDT::Index<T>::Search<Q, notDeleted> (KDTIndex.cpp:182-241) */
static void kdt_search(const qctx_t* c, ws_t* ws, res_t* res, int k)
{
    const ora_index* idx = c->idx;
    for (int i = 0; i < idx->tree_num; i++) kdt_search_node(c, ws, idx->tree_starts[i], 0);
    kdt_search_trees(c, ws, idx->initial_pivots);
    while (ws->ng.count != 0) {
        pair_t gnode = heap_pop(&ws->ng);
        const int32_t* node = idx->graph + (size_t)gnode.node * idx->degree;
        ws->nexpand++;
        if (not_deleted(idx, gnode.node)) {
            if (!res_add_point(res, k, gnode.node, gnode.distance) && ws->checked > ws->max_check) {
                res_sort(res, k);
                return;
            }
        }
        float upperBound = res[0].dist > gnode.distance ? res[0].dist : gnode.distance;
        int bLocalOpt = 1;
        for (int i = 0; i < idx->degree; i++) {
            int32_t nn_index = node[i];
            if (nn_index < 0) break;
            if (ws_check_and_set(ws, nn_index)) continue;
            float distance2leaf = qdist(c, ws, nn_index);
            if (distance2leaf <= upperBound) bLocalOpt = 0;
            ws->checked++;
            pair_t p = {nn_index, distance2leaf};
            heap_insert(&ws->ng, p);
        }
        if (bLocalOpt)
            ws->no_better++;
        else
            ws->no_better = 0;
        if (ws->no_better > idx->no_better_threshold) {
            if (ws->tree_checked <= ws->checked / 10) {
                kdt_search_trees(c, ws, idx->other_pivots + ws->checked);
            } else if (gnode.distance > res[0].dist) {
                break;
            }
        }
    }
    res_sort(res, k);
}

int ora_search_batch(const ora_index* idx, const void* queries, int32_t nq, int32_t k,
                     int32_t* ids, float* dists, int32_t* stats, int32_t threads)
{
    static const size_t elem[4] = {1, 1, 2, 4};
    const size_t row_bytes = elem[idx->value_type] * (size_t)idx->dim;
    const ora_quantizer* quant = idx->quantizer;
    if (quant && idx->tree_kind != ORA_BKT) return 1; /* quantized KDT not restated */
    if (idx->filter && idx->tree_kind != ORA_BKT) return 1; /* "Not Support Filter on KDT Index!" (KDTIndex.cpp:361-365) */
    const size_t query_bytes = quant ? elem[quant->rtype] * (size_t)(quant->m * quant->dsub) : row_bytes;
    /* a fresh thread's work space: Initialize(max(MaxCheck, MaxCheckForRefineGraph)) then
     * Reset(MaxCheck, K) (BKTIndex.cpp:600-605) */
    const int alloc_check = idx->max_check > idx->max_check_refine ? idx->max_check : idx->max_check_refine;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        ws_t ws;
        ws_init(&ws, idx->n, alloc_check);
        res_t* res = (res_t*)malloc(sizeof(res_t) * (size_t)(k > 0 ? k : 1));
        uint8_t* qcode = quant ? (uint8_t*)malloc((size_t)quant->m) : NULL;
        float* qtable = (quant && quant->enable_adc) ? (float*)malloc(sizeof(float) * (size_t)quant->m * quant->ks) : NULL;
        float* qtmp = quant ? (float*)malloc(sizeof(float) * 2 * (size_t)(quant->m * quant->dsub)) : NULL;
#pragma omp for schedule(dynamic, 10)
        for (int32_t q = 0; q < nq; q++) {
            qctx_t c = {idx, (const char*)queries + (size_t)q * query_bytes, row_bytes, idx->metric != ORA_L2};
            if (quant) { /* QueryResultSet::SetTarget -> QuantizeVector (QueryResultSet.h:46-60) */
                if (quant->enable_adc) {
                    adc_table_one(quant, c.query, qtable, qtmp);
                    c.query = qtable;
                } else {
                    quantize_one(quant, c.query, qcode, qtmp);
                    c.query = qcode;
                }
            }
            for (int i = 0; i < k; i++) {
                res[i].vid = -1;
                res[i].dist = kMaxDist();
            }
            ws_reset(&ws, idx->max_check, k);
            if (idx->tree_kind == ORA_BKT)
                bkt_search(&c, &ws, res, k, 0);
            else
                kdt_search(&c, &ws, res, k);
            for (int i = 0; i < k; i++) {
                ids[(size_t)q * k + i] = res[i].vid;
                dists[(size_t)q * k + i] = res[i].dist;
            }
            if (stats) {
                int32_t* s = stats + (size_t)q * ORA_ST_COUNT;
                s[ORA_ST_CHECKED] = ws.checked;
                s[ORA_ST_TREE_CHECKED] = ws.tree_checked;
                s[ORA_ST_NG_LEFT] = ws.ng.count;
                s[ORA_ST_SPT_LEFT] = ws.spt.count;
                s[ORA_ST_NDIST] = ws.ndist;
                s[ORA_ST_NEXPAND] = ws.nexpand;
                s[ORA_ST_NTREE] = ws.ntree;
                s[7] = 0;
            }
        }
        free(res);
        free(qcode);
        free(qtmp);
        ws_free(&ws);
    }
    return 0;
}


/* RelativeNeighborhoodGraph::RebuildNeighbors (RelativeNeighborhoodGraph.h:20-38): walk the ascending result list,
 * keep a candidate unless an already kept neighbour is closer to it than the node is (RNG rule, factor m_fRNGFactor);
 * distances between base rows use the index's own ComputeDistance (VectorIndex.h:136-139). */
static void rebuild_neighbors(const ora_index* idx, int32_t node, int32_t* nodes, const res_t* results, int num_results,
                              int neighborhood, float rng_factor)
{
    static const size_t elem[4] = {1, 1, 2, 4};
    const size_t row_bytes = elem[idx->value_type] * (size_t)idx->dim;
    const char* base = (const char*)idx->vectors;
    int count = 0;
    for (int j = 0; j < num_results && count < neighborhood; j++) {
        const res_t* item = &results[j];
        if (item->vid < 0) break;
        if (item->vid == node) continue;
        int good = 1;
        for (int k = 0; k < count; k++) {
            float d;
            if (idx->quantizer)
                d = ora_quantizer_l2(idx->quantizer, (const uint8_t*)base + (size_t)nodes[k] * row_bytes,
                                     (const uint8_t*)base + (size_t)item->vid * row_bytes);
            else
                d = ora_distance(idx->metric, idx->value_type, idx->simd_width, base + (size_t)nodes[k] * row_bytes,
                                 base + (size_t)item->vid * row_bytes, idx->dim);
            if (rng_factor * d < item->dist) {
                good = 0;
                break;
            }
        }
        if (good) nodes[count++] = item->vid;
    }
    for (int j = count; j < neighborhood; j++) nodes[j] = -1;
}

/* NeighborhoodGraph::RefineNode(index, node, updateNeighbors=false, searchDeleted=false, CEF)
 * (NeighborhoodGraph.h:534-545) for nodes [first_node, first_node+num_nodes) against the index's CURRENT graph:
 * RefineSearchIndex (BKTIndex.cpp:698-711 / KDTIndex.cpp: Reset(MaxCheckForRefineGraph, CEF+1), searchDuplicated =
 * false) with the node's own row as the query, then RebuildNeighbors into out_graph[i*neighborhood ..].
 * The reference refines IN PLACE under OpenMP, so its pass result depends on thread timing; refining every node
 * against the frozen graph is the deterministic form both the oracle and the device implement.
 * res_ids / res_dists: NULL or [num_nodes*(cef+1)] -- the refine-search result lists. */
int ora_refine_nodes(const ora_index* idx, int32_t first_node, int32_t num_nodes, int32_t cef, int32_t neighborhood,
                     float rng_factor, int32_t* out_graph, int32_t* res_ids, float* res_dists, int32_t threads)
{
    static const size_t elem[4] = {1, 1, 2, 4};
    const size_t row_bytes = elem[idx->value_type] * (size_t)idx->dim;
    const ora_quantizer* quant = idx->quantizer;
    if (idx->filter) return 1;
    if (quant && (idx->tree_kind != ORA_BKT || (quant->qtype == ORA_Q_PQ && quant->rtype != ORA_FLOAT))) return 1;
    /* with ADC on, the reference's RebuildNeighbors passes two code rows to the ADC branch of L2Distance, which reads the
     * first as a float table (out of bounds, PQQuantizer.h:114-119): not a defined operation, not restated */
    if (quant && quant->enable_adc) return 1;
    if (first_node < 0 || num_nodes < 0 || first_node + num_nodes > idx->n || cef < 1 || neighborhood < 1) return 1;
    const int k = cef + 1;
    const int alloc_check = idx->max_check > idx->max_check_refine ? idx->max_check : idx->max_check_refine;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        ws_t ws;
        ws_init(&ws, idx->n, alloc_check);
        res_t* res = (res_t*)malloc(sizeof(res_t) * (size_t)k);
        /* quantized index (NeighborhoodGraph.h:538-543): the node's code row is reconstructed, and SetTarget quantizes
         * the reconstruction again -- that, not the stored row, is what the search uses */
        const int qdim = quant ? quant->m * quant->dsub : 0;
        void* rec = quant ? malloc(4 * (size_t)qdim) : NULL;
        uint8_t* qcode = quant ? (uint8_t*)malloc((size_t)quant->m) : NULL;
        float* qtmp = quant ? (float*)malloc(sizeof(float) * 2 * (size_t)qdim) : NULL;
#pragma omp for schedule(dynamic, 10)
        for (int32_t i = 0; i < num_nodes; i++) {
            const int32_t node = first_node + i;
            qctx_t c = {idx, (const char*)idx->vectors + (size_t)node * row_bytes, row_bytes, idx->metric != ORA_L2};
            if (quant) {
                reconstruct_one(quant, (const uint8_t*)c.query, rec, qtmp);
                quantize_one(quant, rec, qcode, qtmp);
                c.query = qcode;
            }
            for (int j = 0; j < k; j++) {
                res[j].vid = -1;
                res[j].dist = kMaxDist();
            }
            ws_reset(&ws, idx->max_check_refine, k);
            if (idx->tree_kind == ORA_BKT)
                bkt_search(&c, &ws, res, k, 1);
            else
                kdt_search(&c, &ws, res, k);
            rebuild_neighbors(idx, node, out_graph + (size_t)i * neighborhood, res, k, neighborhood, rng_factor);
            if (res_ids)
                for (int j = 0; j < k; j++) res_ids[(size_t)i * k + j] = res[j].vid;
            if (res_dists)
                for (int j = 0; j < k; j++) res_dists[(size_t)i * k + j] = res[j].dist;
        }
        free(res);
        free(rec);
        free(qcode);
        free(qtmp);
        ws_free(&ws);
    }
    return 0;
}


/* ------------------------------------------------------------------------------------------ */
/* ResultIterator (ResultIterator.cpp, BKTIndex.cpp:354-427, :650-696): resumable search       */
/* ------------------------------------------------------------------------------------------ */

struct ora_iterator {
    ora_index idx; /* shallow copy: the arrays stay the caller's */
    ws_t ws;
    void* query;
    int is_first, relaxed_mono, max_batch;
    res_t* res;
};

/* VectorIndex::GetIterator (BKTIndex.cpp:650-657): RentWorkSpace(1) = a fresh work space after
 * Initialize(max(MaxCheck, MaxCheckForRefineGraph)) + Reset(MaxCheck, 1) (:686-696).  BKT, no quantizer. */
ora_iterator* ora_iter_open(const ora_index* idx, const void* query)
{
    static const size_t elem[4] = {1, 1, 2, 4};
    if (idx->tree_kind != ORA_BKT || idx->quantizer || idx->filter) return NULL; /* "ITERATIVE NOT SUPPORT FOR KDT" */
    ora_iterator* it = (ora_iterator*)calloc(1, sizeof(*it));
    it->idx = *idx;
    const size_t qb = elem[idx->value_type] * (size_t)idx->dim;
    it->query = malloc(qb);
    memcpy(it->query, query, qb);
    const int alloc_check = idx->max_check > idx->max_check_refine ? idx->max_check : idx->max_check_refine;
    ws_init(&it->ws, idx->n, alloc_check);
    ws_reset(&it->ws, idx->max_check, 1);
    it->is_first = 1;
    it->relaxed_mono = 0;
    it->max_batch = 0;
    it->res = NULL;
    return it;
}

/* BKT::Index<T>::SearchIterative<notDeleted, isDup> (BKTIndex.cpp:354-427) on the iterator's work space */
static int iter_loop(ora_iterator* it, int batch, res_t* res, int is_first)
{
    const ora_index* idx = &it->idx;
    const ora_bkt_node* nodes = (const ora_bkt_node*)idx->nodes;
    static const size_t elem[4] = {1, 1, 2, 4};
    ws_t* ws = &it->ws;
    qctx_t c = {idx, it->query, elem[idx->value_type] * (size_t)idx->dim, idx->metric != ORA_L2};
    if (is_first) {
        bkt_init_search_trees(&c, ws);
        bkt_search_trees(&c, ws, idx->initial_pivots);
    }
    int count = 0;
    const int checkPos = idx->degree - 1;
    while (ws->ng.count != 0) {
        pair_t gnode = heap_pop(&ws->ng);
        int32_t tmpNode = gnode.node;
        const int32_t* node = idx->graph + (size_t)tmpNode * idx->degree;
        ws->nexpand++;
        if (not_deleted(idx, tmpNode)) {
            res_add_point(res, batch, tmpNode, gnode.distance);
            count++;
            if (gnode.distance > dpq_worst(&ws->results) || ws->checked > ws->max_check) it->relaxed_mono = 1;
        }
        int32_t checkNode = node[checkPos];
        if (checkNode < -1) {
            const ora_bkt_node* tnode = &nodes[-2 - checkNode];
            int32_t i = -tnode->childStart;
            while (i < tnode->childEnd) {
                tmpNode = nodes[i].centerid;
                if (not_deleted(idx, tmpNode)) {
                    float d = qdist(&c, ws, tmpNode);
                    if (!ws_check_and_set(ws, tmpNode)) {
                        pair_t p = {tmpNode, d};
                        heap_insert(&ws->ng, p);
                    }
                }
                i++;
            }
        }
        for (int i = 0; i <= checkPos; i++) {
            int32_t nn_index = node[i];
            if (nn_index < 0) break;
            if (ws_check_and_set(ws, nn_index)) continue;
            float d = qdist(&c, ws, nn_index);
            ws->checked++;
            pair_t p = {nn_index, d};
            heap_insert(&ws->ng, p);
            dpq_insert(&ws->results, d);
        }
        if (heap_top(&ws->ng)->distance > heap_top(&ws->spt)->distance)
            bkt_search_trees(&c, ws, idx->other_pivots + ws->checked);
        if (count >= batch) break;
    }
    return count;
}

/* ResultIterator::Next(batch) -> SearchIndexIterativeNext (BKTIndex.cpp:659-675) -> SearchIterative<notDeleted, isDup>
 * (:354-427).  ids/dists: [batch], ascending over the returned entries, unfilled (-1, MaxDist); returns resultCount;
 * *relaxed_mono = WorkSpace::m_relaxedMono after the call. */
int ora_iter_next(ora_iterator* it, int32_t batch, int32_t* ids, float* dists, int32_t* relaxed_mono)
{
    const ora_index* idx = &it->idx;
    ws_t* ws = &it->ws;
    /* ResultIterator::Next (ResultIterator.cpp:31-42, :52): the first call creates the QueryResult with `batch` slots;
     * afterwards the batch is capped by QueryResult::GetResultNum(), which the previous call left at ITS resultCount
     * (SetResultNum(resultCount)) -- an iterator's batch can only shrink */
    const int requested = batch;
    if (it->res == NULL) {
        it->res = (res_t*)malloc(sizeof(res_t) * (size_t)(batch > 0 ? batch : 1));
    } else if (batch > it->max_batch) {
        batch = it->max_batch;
    }
    for (int i = batch; i < requested; i++) {
        ids[i] = -1;
        dists[i] = kMaxDist();
    }
    res_t* res = it->res;
    for (int i = 0; i < batch; i++) { /* QueryResult::Reset */
        res[i].vid = -1;
        res[i].dist = kMaxDist();
    }
    /* WorkSpace::ResetResult(m_iMaxCheck, batch) (WorkSpace.h:280-286) */
    dpq_clear(&ws->results, idx->max_check / 16 > batch ? idx->max_check / 16 : batch);
    ws->no_better = 0;
    ws->tree_checked = 0;
    ws->checked = 0;

    const int count = iter_loop(it, batch, res, it->is_first);
    it->is_first = 0;
    res_sort(res, batch);
    for (int i = 0; i < batch; i++) {
        ids[i] = res[i].vid;
        dists[i] = res[i].dist;
    }
    if (relaxed_mono) *relaxed_mono = it->relaxed_mono;
    it->max_batch = count; /* m_queryResult->SetResultNum(resultCount) */
    return count;
}

void ora_iter_close(ora_iterator* it)
{
    if (!it) return;
    ws_free(&it->ws);
    free(it->query);
    free(it->res);
    free(it);
}


/* BKT::Index<T>::SearchIndexIterativeFromNeareast (BKTIndex.cpp:543-595), the head-index call of SPANN's iterative
 * search (SPANNIndex.cpp:273-285).  First call: a full SearchIndex(query, workspace, searchDeleted = false,
 * searchDuplicated = true) for the k nearest on the rented work space, then the visited set is cleared (the queues keep
 * whatever the search left in them), every result is marked visited and its unvisited graph neighbours enter NGQueue
 * with their distances.  Later calls: ResetResult + SearchIterative(isFirst = false, batch = k).
 * ids/dists: [k]; returns 1 if the first result slot is a real vector (the reference's bool), else 0. */
int ora_iter_next_from_nearest(ora_iterator* it, int32_t k, int32_t* ids, float* dists)
{
    const ora_index* idx = &it->idx;
    static const size_t elem[4] = {1, 1, 2, 4};
    ws_t* ws = &it->ws;
    if (it->res == NULL) {
        it->max_batch = k;
        it->res = (res_t*)malloc(sizeof(res_t) * (size_t)(k > 0 ? k : 1));
    }
    if (k != it->max_batch) return -1; /* the caller's QueryResult keeps its size (SPANN passes the same object) */
    res_t* res = it->res;
    for (int i = 0; i < k; i++) { /* p_headQueryResults->Reset() */
        res[i].vid = -1;
        res[i].dist = kMaxDist();
    }
    dpq_clear(&ws->results, idx->max_check / 16 > k ? idx->max_check / 16 : k); /* ResetResult(m_iMaxCheck, k) */
    ws->no_better = 0;
    ws->tree_checked = 0;
    ws->checked = 0;
    if (it->is_first) {
        qctx_t c = {idx, it->query, elem[idx->value_type] * (size_t)idx->dim, idx->metric != ORA_L2};
        bkt_search(&c, ws, res, k, 0);
        memset(ws->visited, 0, (size_t)ws->n + 1); /* nodeCheckStatus.clear() */
        const int checkPos = idx->degree - 1;
        for (int i = 0; i < k; i++) {
            const int32_t result = res[i].vid;
            if (result < 0) continue;
            ws_check_and_set(ws, result);
            const int32_t* node = idx->graph + (size_t)result * idx->degree;
            for (int j = 0; j <= checkPos; j++) {
                const int32_t nn = node[j];
                if (nn < 0) break;
                if (ws_check_and_set(ws, nn)) continue;
                pair_t p = {nn, qdist(&c, ws, nn)};
                heap_insert(&ws->ng, p);
            }
        }
        it->is_first = 0;
    } else {
        iter_loop(it, k, res, 0);
        res_sort(res, k);
    }
    for (int i = 0; i < k; i++) {
        ids[i] = res[i].vid;
        dists[i] = res[i].dist;
    }
    return res[0].vid >= 0 ? 1 : 0;
}
