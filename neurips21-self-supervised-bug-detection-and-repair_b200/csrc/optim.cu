// Optimiser step over one flat fp32 buffer: global gradient norm, clip, Adam.
// Reference: buglab/models/utils.py:51-52 (Adam lr 1e-4, torch defaults) and
// buglab/models/train.py:104 (clip_gradient_norm=0.5, i.e. torch.nn.utils.clip_grad_norm_).
// The flat gradient buffer is also the single NCCL all-reduce bucket.
#include "common.cuh"

namespace bl {

constexpr int SQN_BLOCKS = 592;  // 4 x 148 SMs, persistent grid-stride

__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
    float s = 0.f;
    const int64_t n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[n4 * 4 + threadIdx.x];
        s += v * v;
    }
    __shared__ float warp_part[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += warp_part[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void sqnorm_final_kernel(const float* __restrict__ partial, int n_partial, float* __restrict__ out) {
    // single warp, fixed order -> deterministic
    float s = 0.f;
    for (int i = threadIdx.x; i < n_partial; i += 32) s += partial[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float max_norm,
            const float* __restrict__ sqnorm, float grad_scale) {
    float coef = grad_scale;
    if (max_norm > 0.f && sqnorm != nullptr) {
        const float total = sqrtf(__ldg(sqnorm)) * grad_scale;
        coef *= fminf(1.0f, max_norm / (total + 1e-6f));
    }
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= step_size * (mi / denom);
    }
}

}  // namespace bl

using namespace bl;

extern "C" int bl_grad_sqnorm(const float* grad, int64_t n, float* sqnorm, float* partial, bl_stream_t stream_) {
    if (n < 0) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (((uintptr_t)grad & 15) != 0) return BL_ERR_INVALID_ARGUMENT;
    sqnorm_partial_kernel<<<SQN_BLOCKS, 256, 0, stream>>>(grad, n, partial);
    sqnorm_final_kernel<<<1, 32, 0, stream>>>(partial, SQN_BLOCKS, sqnorm);
    return check_launch("bl_grad_sqnorm");
}

extern "C" int bl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                            float lr, float beta1, float beta2, float eps, int64_t step, float max_norm,
                            const float* sqnorm, float grad_scale, bl_stream_t stream) {
    if (n < 0 || step < 1) return BL_ERR_INVALID_ARGUMENT;
    if (n == 0) return BL_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<SQN_BLOCKS * 2, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                beta2, eps, (float)bc1, (float)sqrt(bc2),
                                                                max_norm, sqnorm, grad_scale);
    return check_launch("bl_adam_step");
}
