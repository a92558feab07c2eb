// Subtoken embedding lookup + masked max-pool over <= T subtokens per node (ptgnn
// StrElementRepresentationModel, token_splitting="subtoken", subtoken_combination="max";
// reference wiring buglab/models/modelregistry.py:57-66,79-82).  Same kernel family as the edge
// aggregate: gather rows, segmented max with arg-routed backward.  One thread per float4 of a node row.
#include "common.cuh"

namespace bl {

__global__ void __launch_bounds__(256)
subtoken_maxpool_fwd_kernel(const float4* __restrict__ emb, const int* __restrict__ ids,
                            const int* __restrict__ lens, int64_t N, int T, int H4, float p_drop,
                            float scale, uint64_t seed, float4* __restrict__ out, int4* __restrict__ arg) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= N * H4) return;
    const int64_t n = gid / H4;
    const int c = (int)(gid - n * H4);
    const int len = min(max(__ldg(lens + n), 0), T);
    float4 best = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    int4 barg = make_int4(0, 0, 0, 0);
    for (int t = 0; t < len; ++t) {
        const int id = __ldg(ids + n * T + t);
        float4 v = __ldg(emb + (size_t)id * H4 + c);
        if (p_drop > 0.f) {
            const uint64_t e = ((uint64_t)(n * T + t) * H4 + c) * 4;
            v.x = keep_element(seed, e + 0, p_drop) ? v.x * scale : 0.f;
            v.y = keep_element(seed, e + 1, p_drop) ? v.y * scale : 0.f;
            v.z = keep_element(seed, e + 2, p_drop) ? v.z * scale : 0.f;
            v.w = keep_element(seed, e + 3, p_drop) ? v.w * scale : 0.f;
        }
        if (v.x > best.x) { best.x = v.x; barg.x = t; }
        if (v.y > best.y) { best.y = v.y; barg.y = t; }
        if (v.z > best.z) { best.z = v.z; barg.z = t; }
        if (v.w > best.w) { best.w = v.w; barg.w = t; }
    }
    if (len == 0) best = make_float4(0.f, 0.f, 0.f, 0.f);
    out[gid] = best;
    arg[gid] = barg;
}

__global__ void __launch_bounds__(256)
subtoken_maxpool_bwd_kernel(const float* __restrict__ d_out, const int* __restrict__ ids,
                            const int* __restrict__ arg, int64_t N, int T, int H, float p_drop, float scale,
                            uint64_t seed, float* __restrict__ d_emb) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= N * H) return;
    const int64_t n = gid / H;
    const int j = (int)(gid - n * H);
    const int t = arg[gid];
    float g = d_out[gid];
    if (p_drop > 0.f) {
        const uint64_t e = (uint64_t)(n * T + t) * H + j;
        g = keep_element(seed, e, p_drop) ? g * scale : 0.f;
    }
    if (g != 0.f) atomicAdd(d_emb + (size_t)ids[n * T + t] * H + j, g);
}

}  // namespace bl

using namespace bl;

extern "C" int bl_subtoken_maxpool_fwd(const float* emb, const int32_t* ids, const int32_t* lens, int64_t N,
                                       int32_t T, int32_t H, float p_drop, uint64_t seed, float* out,
                                       int32_t* arg, bl_stream_t stream) {
    if (N < 0 || T <= 0 || H <= 0 || (H & 3) || p_drop < 0.f || p_drop >= 1.f) return BL_ERR_INVALID_ARGUMENT;
    if (N == 0) return BL_OK;
    subtoken_maxpool_fwd_kernel<<<grid_for(N * (H / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)emb, ids, lens, N, T, H / 4, p_drop, 1.0f / (1.0f - p_drop), seed, (float4*)out, (int4*)arg);
    return check_launch("bl_subtoken_maxpool_fwd");
}

extern "C" int bl_subtoken_maxpool_bwd(const float* d_out, const int32_t* ids, const int32_t* arg, int64_t N,
                                       int32_t T, int32_t H, float p_drop, uint64_t seed, float* d_emb,
                                       bl_stream_t stream) {
    if (N < 0 || T <= 0 || H <= 0 || (H & 3) || p_drop < 0.f || p_drop >= 1.f) return BL_ERR_INVALID_ARGUMENT;
    if (N == 0) return BL_OK;
    subtoken_maxpool_bwd_kernel<<<grid_for(N * H, 256), 256, 0, (cudaStream_t)stream>>>(
        d_out, ids, arg, N, T, H, p_drop, 1.0f / (1.0f - p_drop), seed, d_emb);
    return check_launch("bl_subtoken_maxpool_bwd");
}
