// common.cuh -- shared helpers of the eventclip_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/eventclip_b200.h"

namespace ec {

void set_error(const char *fmt, ...);

#define EC_CUDA_CHECK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ec::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                          cudaGetErrorString(_e));                                            \
            return EC_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define EC_REQUIRE(cond, ...)                                                                 \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ec::set_error(__VA_ARGS__);                                                       \
            return EC_ERR_ARG;                                                                \
        }                                                                                     \
    } while (0)

// number of SMs of the current device (cached)
int sm_count();

// tcgen05 attention (attention_tc.cu): EC_OK when launched, EC_ERR_UNSUPPORTED when the shape is outside its range
int attention_tc(const void *qkv, void *out, int n_img, int L, int heads, int causal, cudaStream_t stream, float *lse = nullptr);

// tcgen05 attention backward (attention_tc.cu), L <= 256; needs the forward's log-sum-exp
int attention_bwd_tc(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int n_img, int L, int heads,
                     cudaStream_t stream);

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace ec
