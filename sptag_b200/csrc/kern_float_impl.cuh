// kern_float_impl.cuh -- float-row instantiations of search_kernel for one metric (included by kern_float_*.cu)
#pragma once
#include "kernel_select.h"
#include "search_kernels.cuh"

namespace sptag_b200 {

template <int DIM, bool COSINE>
static SearchKernelFn pick_rpl(int mres_cap, bool kdt, int slots) {
    if (kdt) {  // KDT has no m_Results gate
        if constexpr (DIM == 128) {
            if (slots > 16) return search_kernel<DIM, COSINE, 16, true, false, 0, 20>;
        }
        return search_kernel<DIM, COSINE, 16, true>;
    }
    // register caps (MINB resident single-warp CTAs per SM): 12 -> 168 registers; the kernel is latency-bound per
    // warp, so residency beats a few spilled values (refine passes run the 32-register m_Results file, K = CEF+1)
    if (mres_cap <= 32 * 16) return search_kernel<DIM, COSINE, 16, false, false, 0, 12>;
    if (mres_cap <= 32 * 32) return search_kernel<DIM, COSINE, 32, false, false, 0, 14>;
    // 64 registers per lane: K / CEF+1 up to 2048 (the reference's default RefineGraph schedule searches with
    // CEF x CEFScale + 1 = 2001 results, NeighborhoodGraph.h:459-470)
    if (mres_cap <= 32 * 64) return search_kernel<DIM, COSINE, 64, false, false, 0, 8>;
    return nullptr;
}

// slots: the residency (query slots per SM) the host is aiming at; 512-byte rows have a register-capped variant for up to
// 20 slots (96 registers, 4 bytes of spills; a 24-slot / 80-register variant was measured slower: 561k vs 575-594k QPS)
template <bool COSINE>
static SearchKernelFn pick_dim(int dim, int mres_cap, bool kdt, int slots) {
    switch (dim) {
    case 128:
        if (!kdt && mres_cap <= 32 * 16) {
            if (slots > 16) return search_kernel<128, COSINE, 16, false, false, 0, 20>;
            return search_kernel<128, COSINE, 16, false, false, 0, 16>;  // 128 regs, 16/SM
        }
        return pick_rpl<128, COSINE>(mres_cap, kdt, slots);
    case 768:
        // 15 resident queries per SM (127 registers, no spills) when the m_Results file is the 16-register one
        if (!kdt && mres_cap <= 32 * 16) return search_kernel<768, COSINE, 16, false, false, 0, 15>;
        return pick_rpl<768, COSINE>(mres_cap, kdt, slots);
    default: return pick_rpl<0, COSINE>(mres_cap, kdt, slots);
    }
}

}  // namespace sptag_b200
