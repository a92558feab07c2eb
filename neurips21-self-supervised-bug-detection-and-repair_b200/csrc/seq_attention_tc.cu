// Row kernels of the tensor-core attention path of the sequence models (seq-great / seq-rat).
//
// Reference semantics: buglab/models/layers/multihead_attention.py:45-85 and relational_multihead_attention.py:71-178 (see
// seq_attention_core.h for the formulas).  The four GEMM-shaped products run on the TMA-fed tcgen05 kernels of gemm_tma.cu,
// one "segment" per (sample, head) with that head's K / V / Q / dO rows as its "weight matrix":
//     forward   S  = Q K^T          bl_tma_project   (n_out = Lp, k_in = 64)
//               O  = P' V           bl_tma_project   (n_out = 64, k_in = Lp)
//     backward  dP = dO V^T         bl_tma_project   (n_out = Lp, k_in = 64)
//               dQ = dS K           bl_tma_project   (n_out = 64, k_in = Lp)
//               dK = dS^T Q,  dV = P'^T dO           bl_tma_weight_grad (m_out = Lp, n_in = 64)
// and the two kernels below do everything between them, one warp per (sample, head, query) row of the [Lp x Lp] score tile:
//     forward  row kernel: typed-edge terms <q_i, bias[tab_e]> added at the entry positions, key mask, softmax (log-sum-exp
//              kept), dropout, P' written as the fp16 hi/lo split table the next GEMM reads, value-bias terms of "rat";
//     backward row kernel: P' recomputed from the saved scores, dS = P (mask dP - delta) written in place of dP, the
//              per-entry table gradients, and the entry terms of dQ.
// Lp = padded length (128, 256 or 512 keys), head size 64 (smaller heads are zero-padded by the caller).
#include <cuda_fp16.h>

#include "common.cuh"
#include "seq_attention_core.h"

namespace bl {
namespace seqtc {

constexpr int kWarps = 8;
constexpr int kMaxLp = 512;
constexpr int kD = 64;

struct RowProblem {
    int B, H, L, Lp, T2;
    const float* q;             // [B*H*Lp, 64] (pre-scaled queries, zero rows beyond L)
    const int32_t* lengths;     // [B]
    const float* bias;          // [T2, H, 64]
    const float* vbias;         // [T2, H, 64] or nullptr
    const int32_t* row_ptr;     // entries of query row (b, i): row_ptr[b * L + i] .. row_ptr[b * L + i + 1]
    const int32_t* row_key;
    const int32_t* row_tab;
    float p_drop;
    uint64_t seed;
    const float* amax_p;        // device scalar: upper bound of P' (1 / (1 - p_drop)), source of the split table's scale
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

__device__ __forceinline__ float dropout_scale(const RowProblem& p, int b, int h, int i, int j) {
    seqatt::Problem q;
    q.p_drop = p.p_drop; q.seed = p.seed; q.H = p.H; q.L = p.L;
    return seqatt::dropout_scale(q, b, h, i, j);
}

__device__ __forceinline__ void store_split4(__half* hi_row, __half* lo_row, int j4, float4 x) {
    const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
    const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(x.x - b01.x, x.y - b01.y), l23 = __floats2half2_rn(x.z - b23.x, x.w - b23.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    reinterpret_cast<uint2*>(hi_row)[j4] = hv;
    reinterpret_cast<uint2*>(lo_row)[j4] = lv;
}

// ---- forward --------------------------------------------------------------------------------------------------------------
// scores [G*Lp, Lp] in: Q K^T; out: the same plus the entry terms (kept for backward).  p_split: [2][G*Lp + 1][Lp] fp16.
__global__ void __launch_bounds__(kWarps * 32)
softmax_fwd_kernel(const RowProblem p, float* __restrict__ scores, float* __restrict__ lse, __half* __restrict__ p_split,
                   float* __restrict__ o_extra) {
    __shared__ __align__(16) float row_buf[kWarps][kMaxLp];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_rows = (int64_t)p.B * p.H * p.Lp;
    const int64_t row = (int64_t)blockIdx.x * kWarps + warp;
    const int L4 = p.Lp >> 2;  // Lp is a multiple of 128: every lane moves whole 16-byte pieces
    __half* const hi_base = p_split;
    __half* const lo_base = p_split + (size_t)(total_rows + 1) * p.Lp;
    if (blockIdx.x == 0 && warp == 0) {  // the zero (padding) row of both parts
        for (int j = lane; j < p.Lp; j += 32) {
            hi_base[(size_t)total_rows * p.Lp + j] = __float2half_rn(0.f);
            lo_base[(size_t)total_rows * p.Lp + j] = __float2half_rn(0.f);
        }
    }
    if (row >= total_rows) return;
    const int i = (int)(row % p.Lp);
    const int g = (int)(row / p.Lp);
    const int h = g % p.H, b = g / p.H;
    const int len = __ldg(p.lengths + b);
    __half* hi_row = hi_base + (size_t)row * p.Lp;
    __half* lo_row = lo_base + (size_t)row * p.Lp;
    float* sm = row_buf[warp];
    if (i >= len) {  // padding query: no probability mass, never read downstream
        for (int j = lane; j < p.Lp; j += 32) {
            hi_row[j] = __float2half_rn(0.f);
            lo_row[j] = __float2half_rn(0.f);
        }
        if (lane == 0) lse[row] = 0.f;
        if (o_extra != nullptr) reinterpret_cast<float2*>(o_extra + (size_t)row * kD)[lane] = make_float2(0.f, 0.f);
        return;
    }
    float* srow = scores + (size_t)row * p.Lp;
    for (int j4 = lane; j4 < L4; j4 += 32) reinterpret_cast<float4*>(sm)[j4] = __ldg(reinterpret_cast<const float4*>(srow) + j4);
    __syncwarp();
    const int e0 = __ldg(p.row_ptr + b * p.L + i), e1 = __ldg(p.row_ptr + b * p.L + i + 1);
    const float2 q2 = __ldg(reinterpret_cast<const float2*>(p.q + (size_t)row * kD) + lane);
    for (int e = e0; e < e1; ++e) {  // ascending keys; repeated (row, key) entries add up
        const int key = __ldg(p.row_key + e);
        const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias + ((size_t)__ldg(p.row_tab + e) * p.H + h) * kD) + lane);
        const float term = warp_sum(fmaf(q2.x, b2.x, q2.y * b2.y));
        if (lane == 0) sm[key] += term;
        __syncwarp();
    }
    float m = -INFINITY;
    for (int j = lane; j < len; j += 32) m = fmaxf(m, sm[j]);
    m = warp_max(m);
    float l = 0.f;
    for (int j = lane; j < len; j += 32) l += expf(sm[j] - m);
    l = warp_sum(l);
    const float row_lse = m + logf(l);
    if (lane == 0) lse[row] = row_lse;
    const float scale = pow2_scale_for(__ldg(p.amax_p));
    for (int j4 = lane; j4 < L4; j4 += 32) {
        const float4 s4 = reinterpret_cast<const float4*>(sm)[j4];
        if (e1 > e0) reinterpret_cast<float4*>(srow)[j4] = s4;  // rows without entries are unchanged
        const int j = 4 * j4;
        float4 pr;
        pr.x = (j < len) ? expf(s4.x - row_lse) * dropout_scale(p, b, h, i, j) : 0.f;
        pr.y = (j + 1 < len) ? expf(s4.y - row_lse) * dropout_scale(p, b, h, i, j + 1) : 0.f;
        pr.z = (j + 2 < len) ? expf(s4.z - row_lse) * dropout_scale(p, b, h, i, j + 2) : 0.f;
        pr.w = (j + 3 < len) ? expf(s4.w - row_lse) * dropout_scale(p, b, h, i, j + 3) : 0.f;
        reinterpret_cast<float4*>(sm)[j4] = pr;
        store_split4(hi_row, lo_row, j4, make_float4(pr.x * scale, pr.y * scale, pr.z * scale, pr.w * scale));
    }
    if (o_extra != nullptr) {  // "rat": sum over the row's entries of p'[key] * vbias[tab]
        __syncwarp();
        float2 acc = make_float2(0.f, 0.f);
        for (int e = e0; e < e1; ++e) {
            const float w = sm[__ldg(p.row_key + e)];
            const float2 v2 = __ldg(reinterpret_cast<const float2*>(p.vbias + ((size_t)__ldg(p.row_tab + e) * p.H + h) * kD) + lane);
            acc.x = fmaf(w, v2.x, acc.x);
            acc.y = fmaf(w, v2.y, acc.y);
        }
        reinterpret_cast<float2*>(o_extra + (size_t)row * kD)[lane] = acc;
    }
}

// ---- backward -------------------------------------------------------------------------------------------------------------
// d_scores [G*Lp, Lp] in: dP = dO V^T; out: dS.  p_split as in forward (recomputed).  dq_extra [G*Lp, 64]: entry terms of dQ.
// d_entry_bias / d_entry_vbias: [entries, H, Dout] with Dout <= 64 the caller's true head size.
__global__ void __launch_bounds__(kWarps * 32)
softmax_bwd_kernel(const RowProblem p, const float* __restrict__ scores, const float* __restrict__ lse,
                   const float* __restrict__ out, const float* __restrict__ d_out, float* __restrict__ d_scores,
                   __half* __restrict__ p_split, float* __restrict__ dq_extra, float* __restrict__ d_entry_bias,
                   float* __restrict__ d_entry_vbias, int d_entry_dim) {
    __shared__ __align__(16) float ds_buf[kWarps][kMaxLp];
    __shared__ __align__(16) float pr_buf[kWarps][kMaxLp];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_rows = (int64_t)p.B * p.H * p.Lp;
    const int64_t row = (int64_t)blockIdx.x * kWarps + warp;
    __half* const hi_base = p_split;
    __half* const lo_base = p_split + (size_t)(total_rows + 1) * p.Lp;
    if (blockIdx.x == 0 && warp == 0) {
        for (int j = lane; j < p.Lp; j += 32) {
            hi_base[(size_t)total_rows * p.Lp + j] = __float2half_rn(0.f);
            lo_base[(size_t)total_rows * p.Lp + j] = __float2half_rn(0.f);
        }
    }
    if (row >= total_rows) return;
    const int i = (int)(row % p.Lp);
    const int g = (int)(row / p.Lp);
    const int h = g % p.H, b = g / p.H;
    const int len = __ldg(p.lengths + b);
    __half* hi_row = hi_base + (size_t)row * p.Lp;
    __half* lo_row = lo_base + (size_t)row * p.Lp;
    float* drow = d_scores + (size_t)row * p.Lp;
    const bool in_plan = i < p.L;
    const int e0 = in_plan ? __ldg(p.row_ptr + b * p.L + i) : 0, e1 = in_plan ? __ldg(p.row_ptr + b * p.L + i + 1) : 0;
    const bool two = 2 * lane + 1 < d_entry_dim, one = 2 * lane < d_entry_dim;
    if (i >= len) {
        for (int j = lane; j < p.Lp; j += 32) {
            hi_row[j] = __float2half_rn(0.f);
            lo_row[j] = __float2half_rn(0.f);
            drow[j] = 0.f;
        }
        reinterpret_cast<float2*>(dq_extra + (size_t)row * kD)[lane] = make_float2(0.f, 0.f);
        for (int e = e0; e < e1; ++e) {  // entries of padding rows cannot exist for well-formed inputs; keep their slots defined
            float* db = d_entry_bias + ((size_t)e * p.H + h) * d_entry_dim;
            if (one) db[2 * lane] = 0.f;
            if (two) db[2 * lane + 1] = 0.f;
            if (d_entry_vbias != nullptr) {
                float* dvb = d_entry_vbias + ((size_t)e * p.H + h) * d_entry_dim;
                if (one) dvb[2 * lane] = 0.f;
                if (two) dvb[2 * lane + 1] = 0.f;
            }
        }
        return;
    }
    float* ds = ds_buf[warp];
    float* pr = pr_buf[warp];
    const float2 g2 = __ldg(reinterpret_cast<const float2*>(d_out + (size_t)row * kD) + lane);
    const float2 o2 = __ldg(reinterpret_cast<const float2*>(out + (size_t)row * kD) + lane);
    const float2 q2 = __ldg(reinterpret_cast<const float2*>(p.q + (size_t)row * kD) + lane);
    const float delta = warp_sum(fmaf(g2.x, o2.x, g2.y * o2.y));  // sum_j p_ij dP_ij = <dO_i, O_i>
    const int L4 = p.Lp >> 2;
    for (int j4 = lane; j4 < L4; j4 += 32) reinterpret_cast<float4*>(ds)[j4] = __ldg(reinterpret_cast<const float4*>(drow) + j4);
    __syncwarp();
    if (p.vbias != nullptr) {
        for (int e = e0; e < e1; ++e) {
            const int key = __ldg(p.row_key + e);
            const float2 v2 = __ldg(reinterpret_cast<const float2*>(p.vbias + ((size_t)__ldg(p.row_tab + e) * p.H + h) * kD) + lane);
            const float term = warp_sum(fmaf(g2.x, v2.x, g2.y * v2.y));
            if (lane == 0) ds[key] += term;
            __syncwarp();
        }
    }
    const float row_lse = __ldg(lse + row);
    const float scale = pow2_scale_for(__ldg(p.amax_p));
    const float* srow = scores + (size_t)row * p.Lp;
    for (int j4 = lane; j4 < L4; j4 += 32) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(srow) + j4);
        const float4 dp4 = reinterpret_cast<const float4*>(ds)[j4];
        const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, dpv[4] = {dp4.x, dp4.y, dp4.z, dp4.w};
        float kept[4], dsv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = 4 * j4 + c;
            kept[c] = 0.f;
            dsv[c] = 0.f;
            if (j < len) {
                const float prob = expf(sv[c] - row_lse);
                const float mask = dropout_scale(p, b, h, i, j);
                kept[c] = prob * mask;
                dsv[c] = prob * (mask * dpv[c] - delta);
            }
        }
        const float4 d4 = make_float4(dsv[0], dsv[1], dsv[2], dsv[3]);
        reinterpret_cast<float4*>(ds)[j4] = d4;
        reinterpret_cast<float4*>(pr)[j4] = make_float4(kept[0], kept[1], kept[2], kept[3]);
        reinterpret_cast<float4*>(drow)[j4] = d4;
        store_split4(hi_row, lo_row, j4, make_float4(kept[0] * scale, kept[1] * scale, kept[2] * scale, kept[3] * scale));
    }
    __syncwarp();
    float2 acc = make_float2(0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
        const int key = __ldg(p.row_key + e);
        const float dsv = ds[key];
        const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias + ((size_t)__ldg(p.row_tab + e) * p.H + h) * kD) + lane);
        acc.x = fmaf(dsv, b2.x, acc.x);
        acc.y = fmaf(dsv, b2.y, acc.y);
        float* db = d_entry_bias + ((size_t)e * p.H + h) * d_entry_dim;
        if (one) db[2 * lane] = dsv * q2.x;
        if (two) db[2 * lane + 1] = dsv * q2.y;
        if (d_entry_vbias != nullptr) {
            const float w = pr[key];
            float* dvb = d_entry_vbias + ((size_t)e * p.H + h) * d_entry_dim;
            if (one) dvb[2 * lane] = w * g2.x;
            if (two) dvb[2 * lane + 1] = w * g2.y;
        }
    }
    reinterpret_cast<float2*>(dq_extra + (size_t)row * kD)[lane] = acc;
}

static bool shape_ok(int B, int H, int L, int Lp, int T2) {
    return B >= 0 && H > 0 && L > 0 && T2 > 0 && (Lp == 128 || Lp == 256 || Lp == 512) && L <= Lp;
}

}  // namespace seqtc
}  // namespace bl

using namespace bl;

extern "C" int bl_seq_attention_tc_supported(int32_t head_dim, int32_t max_len) {
    return head_dim > 0 && head_dim <= seqtc::kD && max_len > 0 && max_len <= seqtc::kMaxLp;
}

extern "C" int bl_seq_softmax_fwd(float* scores, const float* q, const int32_t* lengths, const float* bias, const float* vbias,
                                  const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab, int32_t B, int32_t H,
                                  int32_t L, int32_t Lp, int32_t T2, float p_drop, uint64_t seed, const float* amax_p,
                                  float* lse, void* p_split, float* o_extra, bl_stream_t stream) {
    if (!scores || !q || !lengths || !bias || !row_ptr || !amax_p || !lse || !p_split) return BL_ERR_INVALID_ARGUMENT;
    if ((vbias != nullptr) != (o_extra != nullptr)) return BL_ERR_INVALID_ARGUMENT;
    if (!seqtc::shape_ok(B, H, L, Lp, T2)) return BL_ERR_UNSUPPORTED;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return BL_ERR_INVALID_ARGUMENT;
    const int64_t rows = (int64_t)B * H * Lp;
    seqtc::RowProblem p{B, H, L, Lp, T2, q, lengths, bias, vbias, row_ptr, row_key, row_tab, p_drop, seed, amax_p};
    seqtc::softmax_fwd_kernel<<<grid_for(rows + 1, seqtc::kWarps), seqtc::kWarps * 32, 0, (cudaStream_t)stream>>>(
        p, scores, lse, (__half*)p_split, o_extra);
    return check_launch("bl_seq_softmax_fwd");
}

extern "C" int bl_seq_softmax_bwd(const float* scores, const float* lse, const float* q, const int32_t* lengths, const float* bias,
                                  const float* vbias, const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab,
                                  int32_t B, int32_t H, int32_t L, int32_t Lp, int32_t T2, float p_drop, uint64_t seed,
                                  const float* amax_p, const float* out, const float* d_out, float* d_scores, void* p_split,
                                  float* dq_extra, float* d_entry_bias, float* d_entry_vbias, int32_t d_entry_dim,
                                  bl_stream_t stream) {
    if (!scores || !lse || !q || !lengths || !bias || !row_ptr || !amax_p || !out || !d_out || !d_scores || !p_split || !dq_extra ||
        !d_entry_bias)
        return BL_ERR_INVALID_ARGUMENT;
    if ((vbias != nullptr) != (d_entry_vbias != nullptr)) return BL_ERR_INVALID_ARGUMENT;
    if (!seqtc::shape_ok(B, H, L, Lp, T2) || d_entry_dim <= 0 || d_entry_dim > seqtc::kD) return BL_ERR_UNSUPPORTED;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return BL_ERR_INVALID_ARGUMENT;
    const int64_t rows = (int64_t)B * H * Lp;
    seqtc::RowProblem p{B, H, L, Lp, T2, q, lengths, bias, vbias, row_ptr, row_key, row_tab, p_drop, seed, amax_p};
    seqtc::softmax_bwd_kernel<<<grid_for(rows + 1, seqtc::kWarps), seqtc::kWarps * 32, 0, (cudaStream_t)stream>>>(
        p, scores, lse, out, d_out, d_scores, (__half*)p_split, dq_extra, d_entry_bias, d_entry_vbias, d_entry_dim);
    return check_launch("bl_seq_softmax_bwd");
}
