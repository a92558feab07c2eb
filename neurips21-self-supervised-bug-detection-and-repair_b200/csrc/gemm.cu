// Pair projections on the tensor cores with fp32-class accuracy: split-fp16 ("f16x3") GEMMs.
//
// The hoisted per-type affine of ptgnn's MlpMessagePassingLayer (U = A_k h_src, V = B_k h_tgt + b_k; reference
// call site buglab/models/gnnlayerdefs.py:6-23) is a true dense GEMM over the unique (type,node) pairs, but the
// path must match the reference's fp32 CPU arithmetic to 1e-4 through 8 layers, which a single TF32/BF16 pass cannot
// (1.6e-3 / 5e-3 max error on a 256-long dot product).  Each fp32 operand is split into two fp16 parts
// (x = x1 + x2 + O(2^-22 x); fp16 keeps 11 significand bits per part, bf16 only 8) and the product is evaluated as
// x1*w1 + x1*w2 + x2*w1 by ONE fp16 GEMM whose reduction dimension is the concatenation [x1 | x1 | x2] . [w1 | w2 | w1]
// (3*D long, fp32 accumulation in the tensor-core accumulators).  Emulated max error at D=256: 1.9e-6, the same as
// fp32 SGEMM (1.4e-6); the bf16 split (9e-6) accumulated past 1e-4 after 8 layers and was rejected on the GPU.
// fp16's narrow exponent is handled by an exact power-of-two pre-scale of gradient tables (scale from a device-side
// amax, undone downstream).  The bias rides along as three extra reduction columns [1 1 1] . [b1 b2 b3], padded to 8.
//
// The GEMMs themselves are plain library GEMMs (cublasGemmEx, fp16 x fp16 -> fp32); the split / gather / transposed
// stacking kernels around them are hand-written here.
#include <cublas_v2.h>
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"

namespace bl {

constexpr int BIAS_PAD = 8;  // reduction columns appended for the bias: 3 used + 5 zero

static cublasHandle_t g_handles[64] = {nullptr};
static std::mutex g_handle_mutex;

static int get_handle(cudaStream_t stream, cublasHandle_t* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return BL_ERR_CUDA;
    {
        std::lock_guard<std::mutex> lock(g_handle_mutex);
        if (g_handles[dev] == nullptr) {
            if (cublasCreate(&g_handles[dev]) != CUBLAS_STATUS_SUCCESS) {
                set_cuda_error(cudaErrorUnknown, "cublasCreate");
                return BL_ERR_CUDA;
            }
            cublasSetMathMode(g_handles[dev], CUBLAS_DEFAULT_MATH);
        }
    }
    *out = g_handles[dev];
    if (cublasSetStream(*out, stream) != CUBLAS_STATUS_SUCCESS) return BL_ERR_CUDA;
    return BL_OK;
}

static int check_blas(cublasStatus_t s, const char* where) {
    if (s != CUBLAS_STATUS_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "%s (cublas status %d)", where, (int)s);
        set_cuda_error(cudaErrorUnknown, buf);
        return BL_ERR_CUDA;
    }
    return BL_OK;
}

__device__ __forceinline__ void split2(float x, __half& hi, __half& lo) {
    x = fminf(fmaxf(x, -65000.f), 65000.f);  // saturate instead of overflowing to inf (pre-scaled tables stay far below)
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// out[r, :] = [hi | hi | lo | 1 1 1 0 0 0 0 0]  of row idx[r] (or r) of `table` (times the pow2 scale if amax is given)
__global__ void __launch_bounds__(256)
rows_split3_kernel(const float* __restrict__ table, const int* __restrict__ idx, int64_t num_rows, int D,
                   const float* __restrict__ amax, __half* __restrict__ out) {
    const int D4 = D / 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= num_rows * (D4 + 1)) return;
    const int64_t r = gid / (D4 + 1);
    const int c = (int)(gid - r * (D4 + 1));
    const int64_t row_stride = 3 * (int64_t)D + BIAS_PAD;
    __half* orow = out + r * row_stride;
    if (c == D4) {  // the bias columns
        const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
#pragma unroll
        for (int j = 0; j < BIAS_PAD; ++j) orow[3 * D + j] = (j < 3) ? one : zero;
        return;
    }
    const int64_t src_row = idx ? (int64_t)__ldg(idx + r) : r;
    float4 v = __ldg(reinterpret_cast<const float4*>(table + src_row * D) + c);
    if (amax != nullptr) {
        const float sc = pow2_scale_for(__ldg(amax));
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    __half h[4], l[4];
    split2(v.x, h[0], l[0]); split2(v.y, h[1], l[1]); split2(v.z, h[2], l[2]); split2(v.w, h[3], l[3]);
    uint2 hp, lp;
    hp.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    hp.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    lp.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    lp.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(orow + 4 * c) = hp;
    *reinterpret_cast<uint2*>(orow + D + 4 * c) = hp;
    *reinterpret_cast<uint2*>(orow + 2 * D + 4 * c) = lp;
}

// out[r, :] = [hi | lo] of row idx[r] (or r): the two-part table the weight-gradient GEMM consumes (2*D fp16 per row)
__global__ void __launch_bounds__(256)
rows_split2_kernel(const float* __restrict__ table, const int* __restrict__ idx, int64_t num_rows, int D,
                   const float* __restrict__ amax, __half* __restrict__ out) {
    const int D4 = D / 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= num_rows * D4) return;
    const int64_t r = gid / D4;
    const int c = (int)(gid - r * D4);
    const int64_t src_row = idx ? (int64_t)__ldg(idx + r) : r;
    float4 v = __ldg(reinterpret_cast<const float4*>(table + src_row * D) + c);
    if (amax != nullptr) {
        const float sc = pow2_scale_for(__ldg(amax));
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    __half h[4], l[4];
    split2(v.x, h[0], l[0]); split2(v.y, h[1], l[1]); split2(v.z, h[2], l[2]); split2(v.w, h[3], l[3]);
    uint2 hp, lp;
    hp.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    hp.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    lp.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    lp.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
    __half* orow = out + r * 2 * (int64_t)D;
    *reinterpret_cast<uint2*>(orow + 4 * c) = hp;
    *reinterpret_cast<uint2*>(orow + D + 4 * c) = lp;
}

// d_weight[k, m, col0 + d] = T[k, m, d] + T[k, m, D + d] + T[k, M + m, d] + T[k, M + m, D + d]   (T is [K, 2M, 2D])
__global__ void combine_weight_grad_kernel(const float* __restrict__ T, int64_t K, int M, int D, int ld, int col0,
                                           const float* __restrict__ amax, float* __restrict__ dW) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= K * M * D) return;
    const int d = (int)(gid % D);
    const int m = (int)((gid / D) % M);
    const int64_t k = gid / ((int64_t)D * M);
    const float* t = T + k * 4 * (int64_t)M * D;
    const int64_t r1 = (int64_t)m * 2 * D, r2 = (int64_t)(M + m) * 2 * D;
    float v = (t[r1 + d] + t[r1 + D + d]) + (t[r2 + d] + t[r2 + D + d]);
    if (amax != nullptr) v *= 1.0f / pow2_scale_for(__ldg(amax));
    dW[(k * M + m) * ld + col0 + d] = v;
}

// Forward weights: w3[k, m, :] = [w1 | w2 | w1 | b1 b2 b3 0..] of W[k, m, col0:col0+D] (row stride ld) and bias[k, m].
__global__ void weights_split3_fwd_kernel(const float* __restrict__ W, const float* __restrict__ bias, int64_t KM, int D,
                                          int ld, int col0, __half* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= KM * (D + 1)) return;
    const int64_t km = gid / (D + 1);
    const int d = (int)(gid - km * (D + 1));
    __half* orow = out + km * (3 * (int64_t)D + BIAS_PAD);
    if (d == D) {
        const float b = bias ? bias[km] : 0.f;
        const __half b1 = __float2half_rn(b);
        const float r1 = b - __half2float(b1);
        const __half b2 = __float2half_rn(r1);
        const __half b3 = __float2half_rn(r1 - __half2float(b2));
        orow[3 * D + 0] = b1; orow[3 * D + 1] = b2; orow[3 * D + 2] = b3;
        for (int j = 3; j < BIAS_PAD; ++j) orow[3 * D + j] = __float2half_rn(0.f);
        return;
    }
    __half hi, lo;
    split2(W[km * ld + col0 + d], hi, lo);
    orow[d] = hi;
    orow[D + d] = lo;
    orow[2 * D + d] = hi;
}

// Backward weights, row-stacked for dX = G3 @ B3:  b3[k, j*M + m, d] = part_j(W[k, m, col0 + d]), parts (hi, lo, hi).
__global__ void weights_stack3_bwd_kernel(const float* __restrict__ W, int64_t K, int M, int D, int ld, int col0,
                                          __half* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= K * M * D) return;
    const int d = (int)(gid % D);
    const int m = (int)((gid / D) % M);
    const int64_t k = gid / ((int64_t)D * M);
    __half hi, lo;
    split2(W[(k * M + m) * ld + col0 + d], hi, lo);
    __half* base = out + k * 3 * (int64_t)M * D;
    base[((int64_t)0 * M + m) * D + d] = hi;
    base[((int64_t)1 * M + m) * D + d] = lo;
    base[((int64_t)2 * M + m) * D + d] = hi;
}

// out[k, :] += sum of a slab of rows of type k.  Work items = (type, slab of COLSUM_ROWS rows), enumerated like GEMM
// tiles from the device-side type_ptr; a block reads whole rows with 128-bit loads (dim4 float4 per row, 256/dim4
// row lanes), reduces the lanes through shared memory and adds one float4 per column group with atomics.
constexpr int COLSUM_ROWS = 2048;
__global__ void __launch_bounds__(256)
grouped_colsum_kernel(const float4* __restrict__ rows, const int* __restrict__ type_ptr, int num_types, int dim4,
                      float* __restrict__ out) {
    __shared__ int prefix[130];
    __shared__ float4 part[256];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < num_types; ++k) {
            prefix[k] = acc;
            acc += (type_ptr[k + 1] - type_ptr[k] + COLSUM_ROWS - 1) / COLSUM_ROWS;
        }
        prefix[num_types] = acc;
    }
    __syncthreads();
    const int total = prefix[num_types];
    const int lanes = 256 / dim4;            // row lanes per block (dim4 <= 256)
    const int col = threadIdx.x % dim4, lane = threadIdx.x / dim4;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
        int k = 0;
        while (work >= prefix[k + 1]) ++k;
        const int lo = type_ptr[k] + (work - prefix[k]) * COLSUM_ROWS;
        const int hi = min(lo + COLSUM_ROWS, type_ptr[k + 1]);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < lanes)
            for (int r = lo + lane; r < hi; r += lanes) {
                const float4 v = __ldg(rows + (size_t)r * dim4 + col);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        part[threadIdx.x] = acc;
        __syncthreads();
        if (lane == 0) {
            for (int l = 1; l < lanes; ++l) {
                const float4 v = part[l * dim4 + col];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            float* o = out + ((size_t)k * dim4 + col) * 4;
            atomicAdd(o + 0, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
        }
        __syncthreads();
    }
}

// amax[0] = max |x[i]|  (non-negative floats order as unsigned ints; amax zeroed by the caller side of the launch)
__global__ void __launch_bounds__(256) absmax_kernel(const float4* __restrict__ x, int64_t n4, unsigned* __restrict__ amax_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}

__global__ void unscale_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ amax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= 1.0f / pow2_scale_for(__ldg(amax));
}

}  // namespace bl
using namespace bl;

extern "C" int bl_unscale_pow2(float* x, int64_t n, const float* amax, bl_stream_t stream) {
    if (n < 0 || amax == nullptr) return BL_ERR_INVALID_ARGUMENT;
    if (n == 0) return BL_OK;
    unscale_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, amax);
    return check_launch("bl_unscale_pow2");
}

extern "C" int bl_rows_split3_f16(const float* table, const int32_t* idx, int64_t num_rows, int32_t dim,
                                  const float* amax, void* out, bl_stream_t stream) {
    if (num_rows < 0 || dim <= 0 || (dim & 3)) return BL_ERR_INVALID_ARGUMENT;
    if (num_rows == 0) return BL_OK;
    rows_split3_kernel<<<grid_for(num_rows * (dim / 4 + 1), 256), 256, 0, (cudaStream_t)stream>>>(
        table, idx, num_rows, dim, amax, (__half*)out);
    return check_launch("bl_rows_split3_f16");
}

extern "C" int bl_rows_split2_f16(const float* table, const int32_t* idx, int64_t num_rows, int32_t dim,
                                  const float* amax, void* out, bl_stream_t stream) {
    if (num_rows < 0 || dim <= 0 || (dim & 3)) return BL_ERR_INVALID_ARGUMENT;
    if (num_rows == 0) return BL_OK;
    rows_split2_kernel<<<grid_for(num_rows * (dim / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        table, idx, num_rows, dim, amax, (__half*)out);
    return check_launch("bl_rows_split2_f16");
}

extern "C" int bl_weights_split3_f16(const float* weight, const float* bias, int32_t num_types, int32_t out_dim,
                                      int32_t in_dim, int32_t ld, int32_t col0, void* w3_fwd, void* b3_bwd,
                                      bl_stream_t stream_) {
    if (num_types <= 0 || out_dim <= 0 || in_dim <= 0 || (in_dim & 3)) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t KM = (int64_t)num_types * out_dim;
    if (w3_fwd)
        weights_split3_fwd_kernel<<<grid_for(KM * (in_dim + 1), 256), 256, 0, stream>>>(weight, bias, KM, in_dim, ld, col0,
                                                                                   (__half*)w3_fwd);
    if (b3_bwd)
        weights_stack3_bwd_kernel<<<grid_for(KM * in_dim, 256), 256, 0, stream>>>(weight, num_types, out_dim, in_dim, ld,
                                                                                col0, (__half*)b3_bwd);
    return check_launch("bl_weights_split3_f16");
}

// out[rows of type k, 0:M] = a3[rows, 0:Kp] . w3[k, 0:M, 0:Kp]^T      (row-major; Kp = 3*D + 8)
extern "C" int bl_pair_project_fwd(const void* a3, const void* w3, const int32_t* type_ptr_host, int32_t num_types,
                                   int32_t out_dim, int32_t in_dim, float* out, bl_stream_t stream_) {
    if (num_types <= 0 || out_dim <= 0 || in_dim <= 0) return BL_ERR_INVALID_ARGUMENT;
    cublasHandle_t h;
    int rc = get_handle((cudaStream_t)stream_, &h);
    if (rc) return rc;
    const int Kp = 3 * in_dim + BIAS_PAD;
    const float one = 1.f, zero = 0.f;
    const __half* A = (const __half*)a3;
    const __half* W = (const __half*)w3;
    for (int k = 0; k < num_types; ++k) {
        const int lo = type_ptr_host[k], hi = type_ptr_host[k + 1];
        if (hi <= lo) continue;
        // column-major view: C^T[M, P] = W_k[M, Kp] . A^T[Kp, P]
        rc = check_blas(cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, out_dim, hi - lo, Kp, &one,
                                     W + (size_t)k * out_dim * Kp, CUDA_R_16F, Kp,
                                     A + (size_t)lo * Kp, CUDA_R_16F, Kp, &zero,
                                     out + (size_t)lo * out_dim, CUDA_R_32F, out_dim,
                                     CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT), "bl_pair_project_fwd");
        if (rc) return rc;
    }
    return BL_OK;
}

// d_rows[rows of type k, 0:D] = g3[rows, 0:3M] . b3[k, 0:3M, 0:D]       (g3 row stride = 3*M + 8)
extern "C" int bl_pair_project_bwd_input(const void* g3, const void* b3, const int32_t* type_ptr_host, int32_t num_types,
                                         int32_t out_dim, int32_t in_dim, float* d_rows, bl_stream_t stream_) {
    if (num_types <= 0 || out_dim <= 0 || in_dim <= 0) return BL_ERR_INVALID_ARGUMENT;
    cublasHandle_t h;
    int rc = get_handle((cudaStream_t)stream_, &h);
    if (rc) return rc;
    const int M = out_dim, D = in_dim;
    const int gstride = 3 * M + BIAS_PAD;
    const float one = 1.f, zero = 0.f;
    const __half* G = (const __half*)g3;
    const __half* B = (const __half*)b3;
    for (int k = 0; k < num_types; ++k) {
        const int lo = type_ptr_host[k], hi = type_ptr_host[k + 1];
        if (hi <= lo) continue;
        // C^T[D, P] = B3^T[D, 3M] . G3^T[3M, P]
        rc = check_blas(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, D, hi - lo, 3 * M, &one,
                                     B + (size_t)k * 3 * M * D, CUDA_R_16F, D,
                                     G + (size_t)lo * gstride, CUDA_R_16F, gstride, &zero,
                                     d_rows + (size_t)lo * D, CUDA_R_32F, D,
                                     CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT), "bl_pair_project_bwd_input");
        if (rc) return rc;
    }
    return BL_OK;
}

// d_weight[k, 0:M, col0:col0+D] = sum over the rows of type k of g^T h.  ONE fp16 GEMM per type computes all four
// hi/lo cross products at once,  T_k[2M, 2D] = [g1|g2]^T . [h1|h2]  (operands read once; the g2.h2 block comes for
// free and is added too), then combine_weight_grad_kernel folds the four blocks (and undoes the pow2 pre-scale).
//   g : [g1|g2] at columns [g_col0, g_col0+2M) of a table with row stride g_stride: either the three-part table
//       [P, 3M+8] = [g1|g1|g2|..] (g_col0 = M) or a two-part table [P, 2M] from bl_rows_split2_f16 (g_col0 = 0)
//   a2: [P, 2D]   = [h1|h2]        (bl_rows_split2_f16)
//   tmp: [num_types, 2M, 2D] fp32 scratch
extern "C" int bl_pair_project_bwd_weight(const void* g, int32_t g_stride, int32_t g_col0, const void* a2,
                                          const int32_t* type_ptr_host, int32_t num_types, int32_t out_dim,
                                          int32_t in_dim, const float* amax, float* tmp, float* d_weight, int32_t ld,
                                          int32_t col0, bl_stream_t stream_) {
    if (num_types <= 0 || out_dim <= 0 || in_dim <= 0 || g_stride < 2 * out_dim) return BL_ERR_INVALID_ARGUMENT;
    cublasHandle_t h;
    int rc = get_handle((cudaStream_t)stream_, &h);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int M = out_dim, D = in_dim;
    const int gstride = g_stride, astride = 2 * D;
    const float one = 1.f, zero = 0.f;
    const __half* G = (const __half*)g + g_col0;
    const __half* A = (const __half*)a2;
    for (int k = 0; k < num_types; ++k) {
        const int lo = type_ptr_host[k], hi = type_ptr_host[k + 1];
        float* C = tmp + (size_t)k * 4 * M * D;
        if (hi <= lo) {
            rc = check_cuda(cudaMemsetAsync(C, 0, (size_t)4 * M * D * sizeof(float), stream), "dW tmp memset");
            if (rc) return rc;
            continue;
        }
        // column-major: C^T[2D, 2M] = H_all^T[2D, P] . G_all[P, 2M]
        rc = check_blas(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, 2 * D, 2 * M, hi - lo, &one,
                                     A + (size_t)lo * astride, CUDA_R_16F, astride,
                                     G + (size_t)lo * gstride, CUDA_R_16F, gstride, &zero,
                                     C, CUDA_R_32F, 2 * D, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT),
                        "bl_pair_project_bwd_weight");
        if (rc) return rc;
    }
    combine_weight_grad_kernel<<<grid_for((int64_t)num_types * M * D, 256), 256, 0, stream>>>(tmp, num_types, M, D, ld, col0,
                                                                                          amax, d_weight);
    return check_launch("bl_pair_project_bwd_weight");
}

extern "C" int bl_grouped_colsum(const float* rows, const int32_t* type_ptr, int32_t num_types, int32_t dim, float* out,
                                 bl_stream_t stream_) {
    if (num_types <= 0 || num_types > 128 || dim <= 0 || (dim & 3) || dim > 1024) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_cuda(cudaMemsetAsync(out, 0, (size_t)num_types * dim * sizeof(float), stream), "bl_grouped_colsum memset");
    if (rc) return rc;
    grouped_colsum_kernel<<<4 * num_sms(), 256, 0, stream>>>((const float4*)rows, type_ptr, num_types, dim / 4, out);
    return check_launch("bl_grouped_colsum");
}

extern "C" int bl_absmax(const float* x, int64_t n, float* amax, bl_stream_t stream_) {
    if (n < 0 || (n & 3) || amax == nullptr) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_cuda(cudaMemsetAsync(amax, 0, sizeof(float), stream), "bl_absmax memset");
    if (rc || n == 0) return rc;
    absmax_kernel<<<4 * num_sms(), 256, 0, stream>>>((const float4*)x, n / 4, (unsigned*)amax);
    return check_launch("bl_absmax");
}
