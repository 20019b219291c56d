// int16 rows (DistanceUtils.cpp:559-596, :930-967)
#include "kern_int_impl.cuh"
namespace sptag_b200 {
SearchKernelFn pick_int16_kernel(bool cosine, int mres_cap, bool kdt) {
    return cosine ? pick_int<true, 3>(mres_cap, kdt) : pick_int<false, 3>(mres_cap, kdt);
}
}  // namespace sptag_b200
