// kern_int_impl.cuh -- integer-row instantiations of search_kernel (included by kern_int8.cu / kern_int16.cu)
#pragma once
#include "kernel_select.h"
#include "search_kernels.cuh"

namespace sptag_b200 {

// ELEM 1 int8, 2 uint8, 3 int16
template <bool COSINE, int ELEM>
static SearchKernelFn pick_int(int mres_cap, bool kdt) {
    if (kdt) return search_kernel<0, COSINE, 16, true, false, ELEM>;
    if (mres_cap <= 32 * 16) return search_kernel<0, COSINE, 16, false, false, ELEM>;
    if (mres_cap <= 32 * 32) return search_kernel<0, COSINE, 32, false, false, ELEM, 12>;
    if (mres_cap <= 32 * 64) return search_kernel<0, COSINE, 64, false, false, ELEM, 8>;
    return nullptr;
}

}  // namespace sptag_b200
