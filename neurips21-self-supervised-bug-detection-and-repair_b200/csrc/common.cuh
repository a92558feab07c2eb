// Shared helpers for the buglab_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/buglab_b200.h"

namespace bl {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (compile-time sizing of partial buffers; launches use num_sms())

// SM count of the CURRENT device, queried once per device (persistent-kernel grid size).
inline int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
        cached[dev] = n;
    }
    return cached[dev];
}

// Records the last CUDA error text for bl_error_string(BL_ERR_CUDA).
void set_cuda_error(cudaError_t e, const char* where);

inline int check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_cuda_error(e, where);
        return BL_ERR_CUDA;
    }
    return BL_OK;
}

inline int check_cuda(cudaError_t e, const char* where) {
    if (e != cudaSuccess) {
        set_cuda_error(e, where);
        return BL_ERR_CUDA;
    }
    return BL_OK;
}

inline unsigned grid_for(int64_t work_items, int threads) {
    int64_t g = (work_items + threads - 1) / threads;
    return (unsigned)(g < 1 ? 1 : g);
}

// exact (erf) GELU, as torch.nn.GELU() default
__device__ __forceinline__ float gelu_exact(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// d/dx GELU(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Order-preserving float <-> int mapping for atomicMax/atomicMin on floats.
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f + 0.0f);  // -0.0 -> +0.0: the two compare equal, as in the reference's float compare
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}

// Counter-based keep mask: one 32-bit hash per element (splitmix64 finaliser over seed+index).
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
__device__ __forceinline__ bool keep_element(uint64_t seed, uint64_t idx, float p_drop) {
    // uniform in [0,1): keep iff u >= p_drop
    const float u = (float)(hash_u32(seed, idx) >> 8) * (1.0f / 16777216.0f);
    return u >= p_drop;
}

// Exact power-of-two scale that brings a table with absolute maximum `amax` to ~2^12 before an fp16 split (fp16
// overflows at 65504 and goes subnormal below 6.1e-5).  amax == 0 or non-finite -> 1.
__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    return exp2f(12.f - ceilf(log2f(amax)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

}  // namespace bl
