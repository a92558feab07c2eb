// Node-state update pieces around the dense Linear(M->D_out):  LayerNorm  and  Tanh+Dropout.
// ptgnn MlpMessagePassingLayer "state update" (reference call site
// buglab/models/gnnlayerdefs.py:6-23).  Row-wise, HBM-bound, 128-bit accesses, one warp per row.
#include "common.cuh"

namespace bl {

constexpr int LN_MAX_CHUNKS = 8;  // float4 chunks per lane -> dim <= 1024

__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gamma,
                     const float4* __restrict__ beta, int64_t rows, int dim4, float eps,
                     float4* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = x + (size_t)row * dim4;
    float4 v[LN_MAX_CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < dim4) {
            v[i] = __ldg(xr + c);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float inv_dim = 1.0f / (float)(dim4 * 4);
    const float mean = warp_sum(s) * inv_dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < dim4) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float var = warp_sum(q) * inv_dim;
    const float rstd = 1.0f / sqrtf(var + eps);
    float4* yr = y + (size_t)row * dim4;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < dim4) {
            const float4 g = __ldg(gamma + c), b = __ldg(beta + c);
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x;
            o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z;
            o.w = (v[i].w - mean) * rstd * g.w + b.w;
            yr[c] = o;
        }
    }
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

// grid = BL_LN_PARTIALS blocks x 256 threads; block b handles rows {b*8+w + k*PARTIALS*8}.
template <int LN_MAX_CHUNKS>  // float4 chunks per lane: dim <= 128 * LN_MAX_CHUNKS (fewer chunks -> fewer registers -> more CTAs/SM)
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                     const float4* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, int64_t rows, int dim4,
                     float4* __restrict__ dx, float* __restrict__ partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 acc_g[LN_MAX_CHUNKS], acc_b[LN_MAX_CHUNKS];
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        acc_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_dim = 1.0f / (float)(dim4 * 4);
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
        const float m = __ldg(mean + row), rs = __ldg(rstd + row);
        float4 dyv[LN_MAX_CHUNKS], xh[LN_MAX_CHUNKS];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = lane + 32 * i;
            if (c < dim4) {
                const float4 d = __ldg(dy + (size_t)row * dim4 + c);
                const float4 xv = __ldg(x + (size_t)row * dim4 + c);
                const float4 g = __ldg(gamma + c);
                xh[i] = make_float4((xv.x - m) * rs, (xv.y - m) * rs, (xv.z - m) * rs, (xv.w - m) * rs);
                acc_g[i].x += d.x * xh[i].x; acc_g[i].y += d.y * xh[i].y;
                acc_g[i].z += d.z * xh[i].z; acc_g[i].w += d.w * xh[i].w;
                acc_b[i].x += d.x; acc_b[i].y += d.y; acc_b[i].z += d.z; acc_b[i].w += d.w;
                dyv[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
                s1 += (dyv[i].x + dyv[i].y) + (dyv[i].z + dyv[i].w);
                s2 += (dyv[i].x * xh[i].x + dyv[i].y * xh[i].y) + (dyv[i].z * xh[i].z + dyv[i].w * xh[i].w);
            }
        }
        s1 = warp_sum(s1) * inv_dim;
        s2 = warp_sum(s2) * inv_dim;
#pragma unroll
        for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
            const int c = lane + 32 * i;
            if (c < dim4) {
                float4 o;
                o.x = rs * (dyv[i].x - s1 - xh[i].x * s2);
                o.y = rs * (dyv[i].y - s1 - xh[i].y * s2);
                o.z = rs * (dyv[i].z - s1 - xh[i].z * s2);
                o.w = rs * (dyv[i].w - s1 - xh[i].w * s2);
                dx[(size_t)row * dim4 + c] = o;
            }
        }
    }
    // block reduction of the 8 warps' column sums through shared memory (fixed order -> deterministic)
    extern __shared__ float4 smem[];  // [8][dim4] for gamma then [8][dim4] for beta
    float4* sg = smem;
    float4* sb = smem + 8 * dim4;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < dim4) {
            sg[warp * dim4 + c] = acc_g[i];
            sb[warp * dim4 + c] = acc_b[i];
        }
    }
    __syncthreads();
    const int dim = dim4 * 4;
    const float* sgf = reinterpret_cast<const float*>(sg);
    const float* sbf = reinterpret_cast<const float*>(sb);
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            a += sgf[w * dim + j];
            b += sbf[w * dim + j];
        }
        partial[(size_t)blockIdx.x * dim + j] = a;
        partial[(size_t)(gridDim.x + blockIdx.x) * dim + j] = b;
    }
}

// Column sums of the per-block partials: block = 32 columns x 32 row lanes (a single thread walking all 1 184 partials of
// its column took 0.29 ms per call — 1 % of the c2 step); fixed summation order, so still deterministic.
__global__ void __launch_bounds__(1024)
layernorm_bwd_reduce(const float* __restrict__ partial, int num_partials, int dim,
                     float* __restrict__ d_gamma, float* __restrict__ d_beta) {
    __shared__ float sa[32][33], sb[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = blockIdx.x * 32 + tx;
    float a = 0.f, b = 0.f;
    if (j < dim) {
        for (int p = ty; p < num_partials; p += 32) {
            a += partial[(size_t)p * dim + j];
            b += partial[(size_t)(num_partials + p) * dim + j];
        }
    }
    sa[ty][tx] = a;
    sb[ty][tx] = b;
    __syncthreads();
    if (ty == 0 && j < dim) {
        float ta = 0.f, tb = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            ta += sa[r][tx];
            tb += sb[r][tx];
        }
        d_gamma[j] = ta;
        d_beta[j] = tb;
    }
}

__global__ void __launch_bounds__(256)
tanh_dropout_fwd_kernel(const float4* __restrict__ x, int64_t n4, float p_drop, float scale, uint64_t seed,
                        float4* __restrict__ y, float4* __restrict__ t_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    float4 t = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    t_out[i] = t;
    if (p_drop > 0.f) {
        const uint64_t e = (uint64_t)i * 4;
        t.x = keep_element(seed, e + 0, p_drop) ? t.x * scale : 0.f;
        t.y = keep_element(seed, e + 1, p_drop) ? t.y * scale : 0.f;
        t.z = keep_element(seed, e + 2, p_drop) ? t.z * scale : 0.f;
        t.w = keep_element(seed, e + 3, p_drop) ? t.w * scale : 0.f;
    }
    y[i] = t;
}

__global__ void __launch_bounds__(256)
tanh_dropout_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ t, int64_t n4,
                        float p_drop, float scale, uint64_t seed, float4* __restrict__ dx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 d = __ldg(dy + i);
    const float4 tv = __ldg(t + i);
    if (p_drop > 0.f) {
        const uint64_t e = (uint64_t)i * 4;
        d.x = keep_element(seed, e + 0, p_drop) ? d.x * scale : 0.f;
        d.y = keep_element(seed, e + 1, p_drop) ? d.y * scale : 0.f;
        d.z = keep_element(seed, e + 2, p_drop) ? d.z * scale : 0.f;
        d.w = keep_element(seed, e + 3, p_drop) ? d.w * scale : 0.f;
    }
    dx[i] = make_float4(d.x * (1.f - tv.x * tv.x), d.y * (1.f - tv.y * tv.y), d.z * (1.f - tv.z * tv.z),
                        d.w * (1.f - tv.w * tv.w));
}

}  // namespace bl

using namespace bl;

extern "C" int bl_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows,
                                int32_t dim, float eps, float* y, float* mean, float* rstd,
                                bl_stream_t stream) {
    if (rows < 0 || dim <= 0 || (dim & 3) || dim > LN_MAX_CHUNKS * 128) return BL_ERR_INVALID_ARGUMENT;
    if (rows == 0) return BL_OK;
    layernorm_fwd_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)x, (const float4*)gamma, (const float4*)beta, rows, dim / 4, eps, (float4*)y, mean, rstd);
    return check_launch("bl_layernorm_fwd");
}

extern "C" int bl_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                const float* rstd, int64_t rows, int32_t dim, float* dx, float* d_gamma,
                                float* d_beta, float* partial, bl_stream_t stream_) {
    if (rows < 0 || dim <= 0 || (dim & 3) || dim > LN_MAX_CHUNKS * 128) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t smem = (size_t)2 * 8 * dim * sizeof(float);
    if (smem > 48 * 1024) {  // only the 1024-wide instantiation needs the opt-in; per-device attribute, set per call
        int rc = check_cuda(cudaFuncSetAttribute(layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 1024 * 4),
                            "bl_layernorm_bwd attribute");
        if (rc) return rc;
    }
#define BL_LAUNCH_LN_BWD(C)                                                                                         \
    layernorm_bwd_kernel<C><<<BL_LN_PARTIALS, 256, smem, stream>>>((const float4*)dy, (const float4*)x,            \
                                                                  (const float4*)gamma, mean, rstd, rows, dim / 4, \
                                                                  (float4*)dx, partial)
    if (dim <= 128) BL_LAUNCH_LN_BWD(1);
    else if (dim <= 256) BL_LAUNCH_LN_BWD(2);
    else if (dim <= 512) BL_LAUNCH_LN_BWD(4);
    else BL_LAUNCH_LN_BWD(8);
#undef BL_LAUNCH_LN_BWD
    layernorm_bwd_reduce<<<(dim + 31) / 32, dim3(32, 32), 0, stream>>>(partial, BL_LN_PARTIALS, dim, d_gamma, d_beta);
    return check_launch("bl_layernorm_bwd");
}

extern "C" int bl_tanh_dropout_fwd(const float* x, int64_t n, float p_drop, uint64_t seed, float* y,
                                   float* t_out, bl_stream_t stream) {
    if (n < 0 || (n & 3) || p_drop < 0.f || p_drop >= 1.f) return BL_ERR_INVALID_ARGUMENT;
    if (n == 0) return BL_OK;
    tanh_dropout_fwd_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)x, n / 4, p_drop, 1.0f / (1.0f - p_drop), seed, (float4*)y, (float4*)t_out);
    return check_launch("bl_tanh_dropout_fwd");
}

extern "C" int bl_tanh_dropout_bwd(const float* dy, const float* t, int64_t n, float p_drop, uint64_t seed,
                                   float* dx, bl_stream_t stream) {
    if (n < 0 || (n & 3) || p_drop < 0.f || p_drop >= 1.f) return BL_ERR_INVALID_ARGUMENT;
    if (n == 0) return BL_OK;
    tanh_dropout_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, (const float4*)t, n / 4, p_drop, 1.0f / (1.0f - p_drop), seed, (float4*)dx);
    return check_launch("bl_tanh_dropout_bwd");
}
