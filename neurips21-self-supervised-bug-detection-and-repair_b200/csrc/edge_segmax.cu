// Fused typed-edge message + max-aggregate, forward and backward (see include/buglab_b200.h).
//
// Replaces ptgnn MlpMessagePassingLayer's  GELU(cat_k Linear_k([h_src;h_tgt]))  ->  scatter_max
// (reference call site buglab/models/gnnlayerdefs.py:6-23).  HBM-bound by construction: per sorted
// edge the kernel reads one U row and one V row (2*M*4 bytes, coalesced 128-bit loads, several edges
// in flight per warp) plus 8 bytes of indices, and does 1 add + 4 compare/selects per channel.
// One warp owns one target node (its CSR segment), so the segmented max needs no atomics; the
// per-channel running extremes live in registers.
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
