// tcgen05 / TMA GEMM engine -- placeholder until the tensor-core path lands (phase B).
#include "common.cuh"
#include "gemm.h"

namespace sfb {
int tc_linear_act_forward(const float*, int64_t, const float*, const float*, float*, int64_t, int64_t, int, int, int, int,
                          cudaStream_t) {
    return SFB_TC_UNSUPPORTED;
}
int tc_linear_backward(const float*, int64_t, const float*, int64_t, const float*, int64_t, int, int, int, float*, float*,
                       int64_t, int, float*, cudaStream_t) {
    return SFB_TC_UNSUPPORTED;
}
}  // namespace sfb

extern "C" int sfb200_tc_available(void) { return 0; }
