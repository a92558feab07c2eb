// Fused typed-edge message + max-aggregate, forward and backward (see include/buglab_b200.h).
//
// Replaces ptgnn MlpMessagePassingLayer's  GELU(cat_k Linear_k([h_src;h_tgt]))  ->  scatter_max
// (reference call site buglab/models/gnnlayerdefs.py:6-23).  HBM-bound by construction: per sorted
// edge the kernel reads one U row and one V row (2*M*4 bytes, coalesced 128-bit loads, several edges
// in flight per warp) plus 8 bytes of indices, and does 1 add + 4 compare/selects per channel.
// One warp owns one target node (its CSR segment), so the segmented max needs no atomics; the
// per-channel running extremes live in registers.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace bl {

// ---------------------------------------------------------------------------------------------
// selection of the winner between the segment's max-x and min-x candidates (GELU quasi-convex)
// torch_scatter CPU semantics: the first (lowest index) edge attaining the maximum message wins.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pick_winner(float xmax, int emax, float xmin, int emin,
                                            float& out_val, float& out_x, int& out_e) {
    if (emax < 0) {  // empty segment
        out_val = 0.f;
        out_x = 0.f;
        out_e = -1;
        return;
    }
    const float gmax = gelu_exact(xmax);
    const float gmin = gelu_exact(xmin);
    const bool take_min = (gmin > gmax) || (gmin == gmax && emin < emax);
    out_val = take_min ? gmin : gmax;
    out_x = take_min ? xmin : xmax;
    out_e = take_min ? emin : emax;
}

// ---------------------------------------------------------------------------------------------
// Fast path: M == 128*ITER, one warp per target node, lane owns float4 chunks {lane + 32*i}.
// UNR edges are fetched (U and V rows) before any of them is consumed -> UNR*2*ITER 16-byte
// loads in flight per lane.
// ---------------------------------------------------------------------------------------------
template <int ITER, int UNR>
__global__ void __launch_bounds__(256)
edge_segmax_fwd_warp(const float* __restrict__ U, const float* __restrict__ V,
                     const int* __restrict__ row_ptr, const int* __restrict__ urow,
                     const int* __restrict__ vrow, int num_nodes,
                     float* __restrict__ agg, float* __restrict__ xwin, int* __restrict__ ewin) {
    constexpr int M4 = 32 * ITER;  // float4 per row
    const int node = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (node >= num_nodes) return;
    const int beg = __ldg(row_ptr + node);
    const int end = __ldg(row_ptr + node + 1);

    float4 xmax[ITER], xmin[ITER];
    int4 emax[ITER], emin[ITER];
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        xmax[i] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        xmin[i] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
        emax[i] = make_int4(-1, -1, -1, -1);
        emin[i] = make_int4(-1, -1, -1, -1);
    }
    const float4* U4 = reinterpret_cast<const float4*>(U);
    const float4* V4 = reinterpret_cast<const float4*>(V);

    for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        const int lidx = base + min(lane, cnt - 1);
        const int my_u = __ldg(urow + lidx);
        const int my_v = __ldg(vrow + lidx);
        for (int t0 = 0; t0 < cnt; t0 += UNR) {
            float4 uu[UNR][ITER], vv[UNR][ITER];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const int t = min(t0 + k, cnt - 1);  // clamp: re-processing an edge is a no-op
                const int u = __shfl_sync(FULL_MASK, my_u, t);
                const int v = __shfl_sync(FULL_MASK, my_v, t);
#pragma unroll
                for (int i = 0; i < ITER; ++i) {
                    uu[k][i] = __ldg(U4 + (size_t)u * M4 + lane + 32 * i);
                    vv[k][i] = __ldg(V4 + (size_t)v * M4 + lane + 32 * i);
                }
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const int e = base + min(t0 + k, cnt - 1);
#pragma unroll
                for (int i = 0; i < ITER; ++i) {
                    const float x0 = uu[k][i].x + vv[k][i].x;
                    const float x1 = uu[k][i].y + vv[k][i].y;
                    const float x2 = uu[k][i].z + vv[k][i].z;
                    const float x3 = uu[k][i].w + vv[k][i].w;
                    if (x0 > xmax[i].x) { xmax[i].x = x0; emax[i].x = e; }
                    if (x1 > xmax[i].y) { xmax[i].y = x1; emax[i].y = e; }
                    if (x2 > xmax[i].z) { xmax[i].z = x2; emax[i].z = e; }
                    if (x3 > xmax[i].w) { xmax[i].w = x3; emax[i].w = e; }
                    if (x0 < xmin[i].x) { xmin[i].x = x0; emin[i].x = e; }
                    if (x1 < xmin[i].y) { xmin[i].y = x1; emin[i].y = e; }
                    if (x2 < xmin[i].z) { xmin[i].z = x2; emin[i].z = e; }
                    if (x3 < xmin[i].w) { xmin[i].w = x3; emin[i].w = e; }
                }
            }
        }
    }

    float4* agg4 = reinterpret_cast<float4*>(agg) + (size_t)node * M4;
    float4* xw4 = reinterpret_cast<float4*>(xwin) + (size_t)node * M4;
    int4* ew4 = reinterpret_cast<int4*>(ewin) + (size_t)node * M4;
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        float4 a, x;
        int4 e;
        pick_winner(xmax[i].x, emax[i].x, xmin[i].x, emin[i].x, a.x, x.x, e.x);
        pick_winner(xmax[i].y, emax[i].y, xmin[i].y, emin[i].y, a.y, x.y, e.y);
        pick_winner(xmax[i].z, emax[i].z, xmin[i].z, emin[i].z, a.z, x.z, e.z);
        pick_winner(xmax[i].w, emax[i].w, xmin[i].w, emin[i].w, a.w, x.w, e.w);
        agg4[lane + 32 * i] = a;
        xw4[lane + 32 * i] = x;
        ew4[lane + 32 * i] = e;
    }
}

// Generic path (any M % 4 == 0): one thread per (node, channel).  Used for small/odd widths.
__global__ void __launch_bounds__(256)
edge_segmax_fwd_generic(const float* __restrict__ U, const float* __restrict__ V,
                        const int* __restrict__ row_ptr, const int* __restrict__ urow,
                        const int* __restrict__ vrow, int64_t num_nodes, int M,
                        float* __restrict__ agg, float* __restrict__ xwin, int* __restrict__ ewin) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= num_nodes * M) return;
    const int node = (int)(gid / M);
    const int j = (int)(gid % M);
    const int beg = row_ptr[node], end = row_ptr[node + 1];
    float xmax = -CUDART_INF_F, xmin = CUDART_INF_F;
    int emax = -1, emin = -1;
    for (int e = beg; e < end; ++e) {
        const float x = U[(size_t)urow[e] * M + j] + V[(size_t)vrow[e] * M + j];
        if (x > xmax) { xmax = x; emax = e; }
        if (x < xmin) { xmin = x; emin = e; }
    }
    float a, x;
    int w;
    pick_winner(xmax, emax, xmin, emin, a, x, w);
    agg[gid] = a;
    xwin[gid] = x;
    ewin[gid] = w;
}

// ---------------------------------------------------------------------------------------------
// Backward.  g = d_agg * GELU'(xwin) goes to the winning edge's U row (REDs: several targets may
// share a (type,src) pair) and V row (plain stores: a (type,tgt) row belongs to exactly one target,
// so the owner warp writes every one of its V rows in full — winners get g, the rest 0).
// ---------------------------------------------------------------------------------------------
constexpr float kFanInHeadroom = 256.f;  // see the amax comment in edge_segmax_bwd_warp

template <int ITER>
__global__ void __launch_bounds__(256)
edge_segmax_bwd_warp(const float* __restrict__ d_agg, const float* __restrict__ xwin,
                     const int* __restrict__ ewin, const int* __restrict__ row_ptr,
                     const int* __restrict__ urow, const int* __restrict__ vrow, int num_nodes,
                     float* __restrict__ dU, float* __restrict__ dV, unsigned* __restrict__ amax_bits) {
    constexpr int M4 = 32 * ITER;
    constexpr int M = 128 * ITER;
    float local_amax = 0.f;
    const int node = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (node >= num_nodes) return;
    const int beg = __ldg(row_ptr + node);
    const int end = __ldg(row_ptr + node + 1);
    if (beg == end) return;

    float g[ITER][4];
    int wv[ITER][4];
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const size_t off = (size_t)node * M4 + lane + 32 * i;
        const float4 d = __ldg(reinterpret_cast<const float4*>(d_agg) + off);
        const float4 x = __ldg(reinterpret_cast<const float4*>(xwin) + off);
        const int4 e = __ldg(reinterpret_cast<const int4*>(ewin) + off);
        const float dd[4] = {d.x, d.y, d.z, d.w};
        const float xx[4] = {x.x, x.y, x.z, x.w};
        const int ee[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            // a non-empty segment always has a winner (ee >= 0)
            g[i][c] = dd[c] * gelu_grad(xx[c]);
            wv[i][c] = __ldg(vrow + ee[c]);
            const int wu = __ldg(urow + ee[c]);
            atomicAdd(dU + (size_t)wu * M + 4 * (lane + 32 * i) + c, g[i][c]);  // result unused -> RED
            local_amax = fmaxf(local_amax, fabsf(g[i][c]));
        }
    }
    if (amax_bits != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_amax = fmaxf(local_amax, __shfl_xor_sync(FULL_MASK, local_amax, o));
        // |dV| <= max|g| exactly; a dU entry sums the g of the targets that share its (type,src) pair.  The published
        // bound leaves room for sums of kFanInHeadroom maximal terms; the fp16 split additionally saturates (never inf).
        if (lane == 0 && local_amax > 0.f) atomicMax(amax_bits, __float_as_uint(local_amax * kFanInHeadroom));
    }
    // every V row of this segment is written exactly once (runs of equal vrow are contiguous)
    int prev_v = -1;
    for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        const int my_v = __ldg(vrow + base + min(lane, cnt - 1));
        for (int t = 0; t < cnt; ++t) {
            const int v = __shfl_sync(FULL_MASK, my_v, t);
            if (v == prev_v) continue;  // warp-uniform
            prev_v = v;
            float4* row = reinterpret_cast<float4*>(dV) + (size_t)v * M4;
#pragma unroll
            for (int i = 0; i < ITER; ++i) {
                float4 o;
                o.x = (wv[i][0] == v) ? g[i][0] : 0.f;
                o.y = (wv[i][1] == v) ? g[i][1] : 0.f;
                o.z = (wv[i][2] == v) ? g[i][2] : 0.f;
                o.w = (wv[i][3] == v) ? g[i][3] : 0.f;
                row[lane + 32 * i] = o;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
edge_segmax_bwd_generic(const float* __restrict__ d_agg, const float* __restrict__ xwin,
                        const int* __restrict__ ewin, const int* __restrict__ row_ptr,
                        const int* __restrict__ urow, const int* __restrict__ vrow,
                        int64_t num_nodes, int M, float* __restrict__ dU, float* __restrict__ dV,
                        unsigned* __restrict__ amax_bits) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= num_nodes * M) return;
    const int node = (int)(gid / M);
    const int j = (int)(gid % M);
    const int beg = row_ptr[node], end = row_ptr[node + 1];
    if (beg == end) return;
    const int e = ewin[gid];
    const float g = d_agg[gid] * gelu_grad(xwin[gid]);
    const int wv = vrow[e];
    atomicAdd(dU + (size_t)urow[e] * M + j, g);
    if (amax_bits != nullptr && g != 0.f) atomicMax(amax_bits, __float_as_uint(fabsf(g) * kFanInHeadroom));
    int prev_v = -1;
    for (int i = beg; i < end; ++i) {
        const int v = vrow[i];
        if (v == prev_v) continue;
        prev_v = v;
        dV[(size_t)v * M + j] = (v == wv) ? g : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// Second-generation backward (feeds csrc/gemm_tma.cu): gradient tables leave the kernels as fp16 hi/lo split tables
// [2][rows + 1][M] (pre-scaled by a power of two, last row of each part zero), every row written by exactly one warp.
//   targets kernel: warp per target node  -> g rows (fp32), dV split rows, per-type column sums of dV (d bias)
//   sources kernel: warp per S-pair row   -> dU split rows = sum over the pair's edges of g masked to the channels won
// ---------------------------------------------------------------------------------------------
constexpr float kGeluGradBound = 1.13f;  // max |GELU'(x)| = 1.1289...

__device__ __forceinline__ void split_store_f16x4(const float4 v, float scale, __half* __restrict__ hi_row, __half* __restrict__ lo_row,
                                                  int chunk) {
    const float x0 = fminf(fmaxf(v.x * scale, -65000.f), 65000.f), x1 = fminf(fmaxf(v.y * scale, -65000.f), 65000.f);
    const float x2 = fminf(fmaxf(v.z * scale, -65000.f), 65000.f), x3 = fminf(fmaxf(v.w * scale, -65000.f), 65000.f);
    const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
    const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(x0 - b01.x, x1 - b01.y), l23 = __floats2half2_rn(x2 - b23.x, x3 - b23.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    reinterpret_cast<uint2*>(hi_row)[chunk] = hv;
    reinterpret_cast<uint2*>(lo_row)[chunk] = lv;
}

template <int ITER>
__global__ void __launch_bounds__(256)
edge_bwd_targets_warp(const float* __restrict__ d_agg, const float* __restrict__ xwin, const int* __restrict__ ewin,
                      const int* __restrict__ row_ptr, const int* __restrict__ vrow, const int* __restrict__ e_type,
                      int num_nodes, int num_types, int64_t num_t_pairs, const float* __restrict__ amax_in,
                      float* __restrict__ amax_eff, float* __restrict__ g_rows, __half* __restrict__ dv_split,
                      float* __restrict__ d_bias) {
    constexpr int M4 = 32 * ITER;
    constexpr int M = 128 * ITER;
    extern __shared__ float bias_acc[];  // [num_types][M] when d_bias != nullptr
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const float amax = __ldg(amax_in) * (kGeluGradBound * kFanInHeadroom);
    const float scale = pow2_scale_for(amax);
    __half* const hi_base = dv_split;
    __half* const lo_base = dv_split + (size_t)(num_t_pairs + 1) * M;
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        if (lane == 0) *amax_eff = amax;
        // the zero (padding) row of both parts
        for (int i = lane; i < M4; i += 32) {
            reinterpret_cast<uint2*>(hi_base + (size_t)num_t_pairs * M)[i] = make_uint2(0u, 0u);
            reinterpret_cast<uint2*>(lo_base + (size_t)num_t_pairs * M)[i] = make_uint2(0u, 0u);
        }
    }
    if (d_bias != nullptr) {
        for (int i = threadIdx.x; i < num_types * M; i += blockDim.x) bias_acc[i] = 0.f;
        __syncthreads();
    }
    for (int node = blockIdx.x * warps_per_block + (threadIdx.x >> 5); node < num_nodes; node += gridDim.x * warps_per_block) {
        const int beg = __ldg(row_ptr + node);
        const int end = __ldg(row_ptr + node + 1);
        float4 g[ITER];
        int4 wv[ITER];
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const size_t off = (size_t)node * M4 + lane + 32 * i;
            g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            wv[i] = make_int4(-1, -1, -1, -1);
            if (beg != end) {  // a non-empty segment always has a winner (ewin >= 0)
                const float4 d = __ldg(reinterpret_cast<const float4*>(d_agg) + off);
                const float4 x = __ldg(reinterpret_cast<const float4*>(xwin) + off);
                const int4 e = __ldg(reinterpret_cast<const int4*>(ewin) + off);
                g[i] = make_float4(d.x * gelu_grad(x.x), d.y * gelu_grad(x.y), d.z * gelu_grad(x.z), d.w * gelu_grad(x.w));
                wv[i] = make_int4(__ldg(vrow + e.x), __ldg(vrow + e.y), __ldg(vrow + e.z), __ldg(vrow + e.w));
            }
            reinterpret_cast<float4*>(g_rows)[off] = g[i];  // isolated nodes: zeros
        }
        // every V row of this segment is written exactly once (runs of equal vrow are contiguous and share one type)
        int prev_v = -1;
        for (int base = beg; base < end; base += 32) {
            const int cnt = min(32, end - base);
            const int my_v = __ldg(vrow + base + min(lane, cnt - 1));
            const int my_t = __ldg(e_type + base + min(lane, cnt - 1));
            for (int t = 0; t < cnt; ++t) {
                const int v = __shfl_sync(FULL_MASK, my_v, t);
                const int type = __shfl_sync(FULL_MASK, my_t, t);
                if (v == prev_v) continue;  // warp-uniform
                prev_v = v;
                __half* hi_row = hi_base + (size_t)v * M;
                __half* lo_row = lo_base + (size_t)v * M;
#pragma unroll
                for (int i = 0; i < ITER; ++i) {
                    float4 o;
                    o.x = (wv[i].x == v) ? g[i].x : 0.f;
                    o.y = (wv[i].y == v) ? g[i].y : 0.f;
                    o.z = (wv[i].z == v) ? g[i].z : 0.f;
                    o.w = (wv[i].w == v) ? g[i].w : 0.f;
                    split_store_f16x4(o, scale, hi_row, lo_row, lane + 32 * i);
                    if (d_bias != nullptr) {
                        float* acc = bias_acc + (size_t)type * M + 4 * (lane + 32 * i);
                        if (o.x != 0.f) atomicAdd(acc + 0, o.x);
                        if (o.y != 0.f) atomicAdd(acc + 1, o.y);
                        if (o.z != 0.f) atomicAdd(acc + 2, o.z);
                        if (o.w != 0.f) atomicAdd(acc + 3, o.w);
                    }
                }
            }
        }
    }
    if (d_bias != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < num_types * M; i += blockDim.x) {
            const float v = bias_acc[i];
            if (v != 0.f) atomicAdd(d_bias + i, v);
        }
    }
}

// UNR S-pair rows per warp, walked in lock step: the chain  edge list -> ewin row of the target -> g row  is latency-bound
// (measured: DRAM 24 %, 18 warps stalled on the scoreboard per issue with one row per warp), so every level is issued for
// UNR rows at once; the per-edge target comes from the plan (s_edge_tgt) instead of a dependent e_tgt[e] load.
template <int ITER, int UNR>
__global__ void __launch_bounds__(256)
edge_bwd_sources_warp(const float* __restrict__ g_rows, const int* __restrict__ ewin, const int* __restrict__ s_edge_ptr,
                      const int* __restrict__ s_edge_idx, const int* __restrict__ s_edge_tgt, int64_t num_s_pairs,
                      const float* __restrict__ amax_eff, __half* __restrict__ du_split) {
    constexpr int M4 = 32 * ITER;
    constexpr int M = 128 * ITER;
    const int64_t row0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * UNR;
    const int lane = threadIdx.x & 31;
    if (row0 > num_s_pairs) return;
    float4 acc[UNR][ITER];
    int beg[UNR], len[UNR];
    int max_len = 0;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int i = 0; i < ITER; ++i) acc[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t row = row0 + u;
        beg[u] = 0;
        len[u] = 0;
        if (row < num_s_pairs) {  // row == num_s_pairs: the zero (padding) row
            beg[u] = __ldg(s_edge_ptr + row);
            len[u] = __ldg(s_edge_ptr + row + 1) - beg[u];
        }
        max_len = max(max_len, len[u]);
    }
    for (int j = 0; j < max_len; ++j) {  // ascending sorted-edge id inside a row: a fixed summation order
        int e[UNR], t[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            e[u] = -2;  // never equals an ewin entry (>= -1)
            t[u] = 0;
            if (j < len[u]) {
                e[u] = __ldg(s_edge_idx + beg[u] + j);
                t[u] = __ldg(s_edge_tgt + beg[u] + j);
            }
        }
        int4 w[UNR][ITER];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int i = 0; i < ITER; ++i)
                w[u][i] = (j < len[u]) ? __ldg(reinterpret_cast<const int4*>(ewin) + (size_t)t[u] * M4 + lane + 32 * i)
                                       : make_int4(-1, -1, -1, -1);
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int i = 0; i < ITER; ++i) {
                const int4 ww = w[u][i];
                if (ww.x == e[u] || ww.y == e[u] || ww.z == e[u] || ww.w == e[u]) {
                    const float4 gv = __ldg(reinterpret_cast<const float4*>(g_rows) + (size_t)t[u] * M4 + lane + 32 * i);
                    if (ww.x == e[u]) acc[u][i].x += gv.x;
                    if (ww.y == e[u]) acc[u][i].y += gv.y;
                    if (ww.z == e[u]) acc[u][i].z += gv.z;
                    if (ww.w == e[u]) acc[u][i].w += gv.w;
                }
            }
    }
    const float scale = pow2_scale_for(__ldg(amax_eff));
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const int64_t row = row0 + u;
        if (row > num_s_pairs) break;
        __half* hi_row = du_split + (size_t)row * M;
        __half* lo_row = du_split + (size_t)(num_s_pairs + 1 + row) * M;
#pragma unroll
        for (int i = 0; i < ITER; ++i) split_store_f16x4(acc[u][i], scale, hi_row, lo_row, lane + 32 * i);
    }
}

}  // namespace bl

using namespace bl;

extern "C" int bl_edge_segmax_fwd(const float* u_rows, const float* v_rows, const int32_t* row_ptr,
                                  const int32_t* urow, const int32_t* vrow, int64_t num_nodes,
                                  int32_t msg_dim, float* agg, float* xwin, int32_t* ewin,
                                  bl_stream_t stream_) {
    if (num_nodes < 0 || msg_dim <= 0 || (msg_dim & 3) || num_nodes > 0x7fffffffLL) return BL_ERR_INVALID_ARGUMENT;
    if (num_nodes == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int threads = 256;
    const unsigned warp_grid = grid_for(num_nodes * 32, threads);
    const int n = (int)num_nodes;
    if (msg_dim == 128) {
        edge_segmax_fwd_warp<1, 8><<<warp_grid, threads, 0, stream>>>(u_rows, v_rows, row_ptr, urow, vrow, n, agg, xwin, ewin);
    } else if (msg_dim == 256) {
        edge_segmax_fwd_warp<2, 4><<<warp_grid, threads, 0, stream>>>(u_rows, v_rows, row_ptr, urow, vrow, n, agg, xwin, ewin);
    } else if (msg_dim == 512) {
        edge_segmax_fwd_warp<4, 2><<<warp_grid, threads, 0, stream>>>(u_rows, v_rows, row_ptr, urow, vrow, n, agg, xwin, ewin);
    } else {
        edge_segmax_fwd_generic<<<grid_for(num_nodes * msg_dim, threads), threads, 0, stream>>>(
            u_rows, v_rows, row_ptr, urow, vrow, num_nodes, msg_dim, agg, xwin, ewin);
    }
    return check_launch("bl_edge_segmax_fwd");
}

extern "C" int bl_edge_segmax_bwd(const float* d_agg, const float* xwin, const int32_t* ewin,
                                  const int32_t* row_ptr, const int32_t* urow, const int32_t* vrow,
                                  int64_t num_nodes, int32_t msg_dim, int64_t num_s_pairs,
                                  int64_t num_t_pairs, float* d_u_rows, float* d_v_rows, float* amax,
                                  bl_stream_t stream_) {
    if (num_nodes < 0 || msg_dim <= 0 || (msg_dim & 3) || num_nodes > 0x7fffffffLL) return BL_ERR_INVALID_ARGUMENT;
    (void)num_t_pairs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_s_pairs > 0) {
        int rc = check_cuda(cudaMemsetAsync(d_u_rows, 0, (size_t)num_s_pairs * msg_dim * sizeof(float), stream),
                            "bl_edge_segmax_bwd memset");
        if (rc) return rc;
    }
    unsigned* amax_bits = reinterpret_cast<unsigned*>(amax);
    if (amax != nullptr) {
        int rc = check_cuda(cudaMemsetAsync(amax, 0, sizeof(float), stream), "bl_edge_segmax_bwd amax memset");
        if (rc) return rc;
    }
    if (num_nodes == 0) return BL_OK;
    const int threads = 256;
    const unsigned warp_grid = grid_for(num_nodes * 32, threads);
    const int n = (int)num_nodes;
    if (msg_dim == 128) {
        edge_segmax_bwd_warp<1><<<warp_grid, threads, 0, stream>>>(d_agg, xwin, ewin, row_ptr, urow, vrow, n, d_u_rows, d_v_rows, amax_bits);
    } else if (msg_dim == 256) {
        edge_segmax_bwd_warp<2><<<warp_grid, threads, 0, stream>>>(d_agg, xwin, ewin, row_ptr, urow, vrow, n, d_u_rows, d_v_rows, amax_bits);
    } else if (msg_dim == 512) {
        edge_segmax_bwd_warp<4><<<warp_grid, threads, 0, stream>>>(d_agg, xwin, ewin, row_ptr, urow, vrow, n, d_u_rows, d_v_rows, amax_bits);
    } else {
        edge_segmax_bwd_generic<<<grid_for(num_nodes * msg_dim, threads), threads, 0, stream>>>(
            d_agg, xwin, ewin, row_ptr, urow, vrow, num_nodes, msg_dim, d_u_rows, d_v_rows, amax_bits);
    }
    return check_launch("bl_edge_segmax_bwd");
}

extern "C" int bl_edge_bwd_targets(const float* d_agg, const float* xwin, const int32_t* ewin, const int32_t* row_ptr,
                                   const int32_t* vrow, const int32_t* e_type, int64_t num_nodes, int32_t msg_dim,
                                   int32_t num_edge_types, int64_t num_t_pairs, const float* amax_in, float* amax_eff,
                                   float* g_rows, void* dv_split, float* d_bias, bl_stream_t stream_) {
    if (num_nodes <= 0 || num_nodes > 0x7fffffffLL || num_edge_types <= 0 || num_t_pairs < 0 || amax_in == nullptr || amax_eff == nullptr)
        return BL_ERR_INVALID_ARGUMENT;
    if (msg_dim != 128 && msg_dim != 256 && msg_dim != 512) return BL_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t smem = d_bias ? (size_t)num_edge_types * msg_dim * sizeof(float) : 0;
    if (smem > 200 * 1024) return BL_ERR_UNSUPPORTED;
    const int threads = 256;
    const int grid = (int)std::min<int64_t>((num_nodes * 32 + threads - 1) / threads, (int64_t)num_sms() * 8);
    const int n = (int)num_nodes;
#define BL_LAUNCH_EBT(ITER)                                                                                            \
    do {                                                                                                               \
        if (smem > 48 * 1024) {                                                                                        \
            int rc = check_cuda(cudaFuncSetAttribute(edge_bwd_targets_warp<ITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                                "bl_edge_bwd_targets attribute");                                                      \
            if (rc) return rc;                                                                                         \
        }                                                                                                              \
        edge_bwd_targets_warp<ITER><<<grid, threads, smem, stream>>>(d_agg, xwin, ewin, row_ptr, vrow, e_type, n, num_edge_types, \
                                                                     num_t_pairs, amax_in, amax_eff, g_rows, (__half*)dv_split, d_bias); \
    } while (0)
    if (msg_dim == 128) BL_LAUNCH_EBT(1);
    else if (msg_dim == 256) BL_LAUNCH_EBT(2);
    else BL_LAUNCH_EBT(4);
#undef BL_LAUNCH_EBT
    return check_launch("bl_edge_bwd_targets");
}

extern "C" int bl_edge_bwd_sources(const float* g_rows, const int32_t* ewin, const int32_t* s_edge_ptr, const int32_t* s_edge_idx,
                                   const int32_t* s_edge_tgt, int64_t num_s_pairs, int32_t msg_dim, const float* amax_eff,
                                   void* du_split, bl_stream_t stream_) {
    if (num_s_pairs < 0 || amax_eff == nullptr) return BL_ERR_INVALID_ARGUMENT;
    if (msg_dim != 128 && msg_dim != 256 && msg_dim != 512) return BL_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int threads = 256;
    constexpr int UNR = 2;
    const unsigned grid = grid_for(((num_s_pairs + 1 + UNR - 1) / UNR) * 32, threads);
    if (msg_dim == 128)
        edge_bwd_sources_warp<1, UNR><<<grid, threads, 0, stream>>>(g_rows, ewin, s_edge_ptr, s_edge_idx, s_edge_tgt, num_s_pairs, amax_eff, (__half*)du_split);
    else if (msg_dim == 256)
        edge_bwd_sources_warp<2, UNR><<<grid, threads, 0, stream>>>(g_rows, ewin, s_edge_ptr, s_edge_idx, s_edge_tgt, num_s_pairs, amax_eff, (__half*)du_split);
    else
        edge_bwd_sources_warp<4, UNR><<<grid, threads, 0, stream>>>(g_rows, ewin, s_edge_ptr, s_edge_idx, s_edge_tgt, num_s_pairs, amax_eff, (__half*)du_split);
    return check_launch("bl_edge_bwd_sources");
}
