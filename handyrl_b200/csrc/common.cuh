// Shared helpers for the sm_100a kernels of the HandyRL learner hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/hrl_b200.h"

namespace hrl {

void set_error(const char *fmt, ...);

#define HRL_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t err__ = (expr);                                                       \
        if (err__ != cudaSuccess) {                                                       \
            hrl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),     \
                           __FILE__, __LINE__);                                           \
            return HRL_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define HRL_REQUIRE(cond, code, ...)        \
    do {                                    \
        if (!(cond)) {                      \
            hrl::set_error(__VA_ARGS__);    \
            return (code);                  \
        }                                   \
    } while (0)

constexpr int kNumSM = 148;  // B200

// reductions over a power-of-two group of W adjacent lanes (W <= 32)
template <int W>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, W));
    return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, W);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming global accesses: data touched exactly once should not displace L1 lines
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }

}  // namespace hrl
