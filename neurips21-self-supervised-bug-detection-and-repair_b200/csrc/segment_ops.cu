// Segment primitives of the localization / repair heads: torch_scatter.scatter_{max,min,sum} and
// scatter_log_softmax (reference buglab/models/utils.py:15-48; call sites
// buglab/models/layers/localizationmodule.py:59,75,105 and buglab/models/gnn.py:299,305).
// Inputs are small (candidates / rewrites of one minibatch); index order is arbitrary, so extremes
// go through order-preserving integer atomics and the arg is resolved in a second pass with
// atomicMin — "first index attaining the extreme wins", exactly torch_scatter's CPU behaviour.
#include "common.cuh"

namespace bl {

__global__ void fill_i32(int* __restrict__ p, int64_t n, int v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// pass 1: ordered-int extreme into key[S*F]
__global__ void seg_extreme_pass1(const float* __restrict__ src, const int* __restrict__ index, int64_t LF,
                                  int F, int is_min, int* __restrict__ key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LF) return;
    const int64_t l = i / F;
    const int f = (int)(i - l * F);
    const int k = float_to_ordered(src[i]);
    int* dst = key + (size_t)index[l] * F + f;
    if (is_min) atomicMin(dst, k); else atomicMax(dst, k);
}
// pass 2: first l whose value equals the extreme
__global__ void seg_extreme_pass2(const float* __restrict__ src, const int* __restrict__ index, int64_t LF,
                                  int F, const int* __restrict__ key, int* __restrict__ arg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LF) return;
    const int64_t l = i / F;
    const int f = (int)(i - l * F);
    const size_t o = (size_t)index[l] * F + f;
    if (float_to_ordered(src[i]) == key[o]) atomicMin(arg + o, (int)l);
}
// pass 3: decode; empty segments -> value 0, arg L
__global__ void seg_extreme_pass3(const int* __restrict__ key, const int* __restrict__ arg, int64_t SF, int L,
                                  float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= SF) return;
    out[i] = (arg[i] >= L) ? 0.f : ordered_to_float(key[i]);
}

__global__ void seg_minmax_bwd_kernel(const float* __restrict__ d_out, const int* __restrict__ arg,
                                      const int* __restrict__ index, int64_t LF, int F,
                                      float* __restrict__ d_src) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LF) return;
    const int64_t l = i / F;
    const int f = (int)(i - l * F);
    const size_t o = (size_t)index[l] * F + f;
    d_src[i] = (arg[o] == (int)l) ? d_out[o] : 0.f;
}

__global__ void seg_sum_kernel(const float* __restrict__ src, const int* __restrict__ index, int64_t LF, int F,
                               float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LF) return;
    const int64_t l = i / F;
    const int f = (int)(i - l * F);
    atomicAdd(out + (size_t)index[l] * F + f, src[i]);
}

// --- log-softmax over segments (1-D) -------------------------------------------------------------
__global__ void lsm_max_kernel(const float* __restrict__ src, const int* __restrict__ index, int64_t L,
                               int* __restrict__ key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) atomicMax(key + index[i], float_to_ordered(src[i]));
}
__global__ void lsm_decode_max(int* __restrict__ key_inout, int64_t S, float* __restrict__ seg_sum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const int k = key_inout[i];
    // empty segment: torch_scatter returns 0 for the max
    const float m = (k == INT_MIN) ? 0.f : ordered_to_float(k);
    reinterpret_cast<float*>(key_inout)[i] = m;
    seg_sum[i] = 0.f;
}
__global__ void lsm_sum_kernel(const float* __restrict__ src, const int* __restrict__ index, int64_t L,
                               const float* __restrict__ seg_max, float* __restrict__ seg_sum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const int s = index[i];
    atomicAdd(seg_sum + s, expf(src[i] - seg_max[s]));
}
__global__ void lsm_out_kernel(const float* __restrict__ src, const int* __restrict__ index, int64_t L,
                               const float* __restrict__ seg_max, const float* __restrict__ seg_sum, float eps,
                               float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const int s = index[i];
    out[i] = (src[i] - seg_max[s]) - logf(seg_sum[s] + eps);
}
__global__ void lsm_bwd_out(const float* __restrict__ d_out, const float* __restrict__ out,
                            const int* __restrict__ index, int64_t L, const float* __restrict__ seg_tmp,
                            float* __restrict__ d_src) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    d_src[i] = d_out[i] - expf(out[i]) * seg_tmp[index[i]];
}

}  // namespace bl

using namespace bl;

extern "C" int bl_segment_minmax(const float* src, const int32_t* index, int64_t L, int32_t F, int64_t S,
                                 int32_t is_min, float* out, int32_t* arg, bl_stream_t stream_) {
    if (L < 0 || F <= 0 || S < 0 || L > 0x7fffffffLL) return BL_ERR_INVALID_ARGUMENT;
    if (S == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int T = 256;
    const int64_t SF = S * F, LF = L * F;
    int* key = reinterpret_cast<int*>(out);  // `out` doubles as the ordered-int key buffer
    fill_i32<<<grid_for(SF, T), T, 0, stream>>>(key, SF, is_min ? INT_MAX : INT_MIN);
    fill_i32<<<grid_for(SF, T), T, 0, stream>>>(arg, SF, (int)L);
    if (LF > 0) {
        seg_extreme_pass1<<<grid_for(LF, T), T, 0, stream>>>(src, index, LF, F, is_min, key);
        seg_extreme_pass2<<<grid_for(LF, T), T, 0, stream>>>(src, index, LF, F, key, arg);
    }
    seg_extreme_pass3<<<grid_for(SF, T), T, 0, stream>>>(key, arg, SF, (int)L, out);
    return check_launch("bl_segment_minmax");
}

extern "C" int bl_segment_minmax_bwd(const float* d_out, const int32_t* arg, const int32_t* index, int64_t L,
                                     int32_t F, float* d_src, bl_stream_t stream) {
    if (L < 0 || F <= 0) return BL_ERR_INVALID_ARGUMENT;
    if (L == 0) return BL_OK;
    seg_minmax_bwd_kernel<<<grid_for(L * F, 256), 256, 0, (cudaStream_t)stream>>>(d_out, arg, index, L * F, F, d_src);
    return check_launch("bl_segment_minmax_bwd");
}

extern "C" int bl_segment_sum(const float* src, const int32_t* index, int64_t L, int32_t F, int64_t S,
                              float* out, bl_stream_t stream_) {
    if (L < 0 || F <= 0 || S < 0) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (S > 0) {
        int rc = check_cuda(cudaMemsetAsync(out, 0, (size_t)S * F * sizeof(float), stream), "bl_segment_sum memset");
        if (rc) return rc;
    }
    if (L == 0 || S == 0) return BL_OK;
    seg_sum_kernel<<<grid_for(L * F, 256), 256, 0, stream>>>(src, index, L * F, F, out);
    return check_launch("bl_segment_sum");
}

extern "C" int bl_segment_log_softmax_fwd(const float* src, const int32_t* index, int64_t L, int64_t S,
                                          float eps, float* out, float* seg_max, float* seg_sum,
                                          bl_stream_t stream_) {
    if (L < 0 || S < 0) return BL_ERR_INVALID_ARGUMENT;
    if (S == 0 || L == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int T = 256;
    int* key = reinterpret_cast<int*>(seg_max);
    fill_i32<<<grid_for(S, T), T, 0, stream>>>(key, S, INT_MIN);
    lsm_max_kernel<<<grid_for(L, T), T, 0, stream>>>(src, index, L, key);
    lsm_decode_max<<<grid_for(S, T), T, 0, stream>>>(key, S, seg_sum);
    lsm_sum_kernel<<<grid_for(L, T), T, 0, stream>>>(src, index, L, seg_max, seg_sum);
    lsm_out_kernel<<<grid_for(L, T), T, 0, stream>>>(src, index, L, seg_max, seg_sum, eps, out);
    return check_launch("bl_segment_log_softmax_fwd");
}

extern "C" int bl_segment_log_softmax_bwd(const float* d_out, const float* out, const int32_t* index,
                                          int64_t L, int64_t S, float* d_src, float* seg_tmp,
                                          bl_stream_t stream_) {
    if (L < 0 || S < 0) return BL_ERR_INVALID_ARGUMENT;
    if (S == 0 || L == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_cuda(cudaMemsetAsync(seg_tmp, 0, (size_t)S * sizeof(float), stream), "lsm bwd memset");
    if (rc) return rc;
    seg_sum_kernel<<<grid_for(L, 256), 256, 0, stream>>>(d_out, index, L, 1, seg_tmp);
    lsm_bwd_out<<<grid_for(L, 256), 256, 0, stream>>>(d_out, out, index, L, seg_tmp, d_src);
    return check_launch("bl_segment_log_softmax_bwd");
}
