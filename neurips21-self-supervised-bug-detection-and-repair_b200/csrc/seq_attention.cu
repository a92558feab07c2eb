// CUDA kernels + C ABI of the edge-biased attention of the sequence models: one thread per (sample, head, row), the row
// arithmetic itself lives in seq_attention_core.h (shared with the host emulation).  Consecutive lanes hold consecutive
// rows of one (sample, head), so the K / V / table rows every lane needs at the same loop step are warp-wide broadcasts.
// First correct path (fp32 CUDA cores); the two GEMM-shaped loops move to tcgen05 once this is parity-green on a B200.
#include "common.cuh"
#include "seq_attention_core.h"

namespace bl {

template <int D>
__global__ void __launch_bounds__(128) seq_attention_fwd_kernel(const seqatt::Problem p, float* __restrict__ out,
                                                                 float* __restrict__ lse) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)p.B * p.H * p.L) return;
    const int i = (int)(t % p.L);
    const int h = (int)((t / p.L) % p.H);
    const int b = (int)(t / ((int64_t)p.L * p.H));
    seqatt::forward_row<D>(p, b, h, i, out, lse);
}

template <int D>
__global__ void __launch_bounds__(128) seq_attention_bwd_row_kernel(const seqatt::Problem p, const float* __restrict__ out,
                                                                     const float* __restrict__ lse,
                                                                     const float* __restrict__ d_out, float* __restrict__ dq,
                                                                     float* __restrict__ d_entry_bias,
                                                                     float* __restrict__ d_entry_vbias,
                                                                     float* __restrict__ delta) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)p.B * p.H * p.L) return;
    const int i = (int)(t % p.L);
    const int h = (int)((t / p.L) % p.H);
    const int b = (int)(t / ((int64_t)p.L * p.H));
    seqatt::backward_row<D>(p, out, lse, d_out, b, h, i, dq, d_entry_bias, d_entry_vbias, delta);
}

template <int D>
__global__ void __launch_bounds__(128) seq_attention_bwd_col_kernel(const seqatt::Problem p, const float* __restrict__ lse,
                                                                     const float* __restrict__ delta,
                                                                     const float* __restrict__ d_out, float* __restrict__ dk,
                                                                     float* __restrict__ dv) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)p.B * p.H * p.L) return;
    const int j = (int)(t % p.L);
    const int h = (int)((t / p.L) % p.H);
    const int b = (int)(t / ((int64_t)p.L * p.H));
    seqatt::backward_col<D>(p, lse, delta, d_out, b, h, j, dk, dv);
}

static bool supported_head_dim(int D) { return D == 8 || D == 16 || D == 32 || D == 64; }

static seqatt::Problem make_problem(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                                    const float* vbias, const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab,
                                    const int32_t* col_ptr, const int32_t* col_query, const int32_t* col_tab, int B, int H, int L,
                                    int T2, float p_drop, uint64_t seed) {
    seqatt::Problem p;
    p.p_drop = p_drop; p.seed = seed;
    p.B = B; p.H = H; p.L = L; p.T2 = T2;
    p.q = q; p.k = k; p.v = v; p.lengths = lengths; p.bias = bias; p.vbias = vbias;
    p.row_ptr = row_ptr; p.row_key = row_key; p.row_tab = row_tab;
    p.col_ptr = col_ptr; p.col_query = col_query; p.col_tab = col_tab;
    return p;
}

}  // namespace bl

extern "C" int bl_seq_attention_supported(int32_t head_dim) { return bl::supported_head_dim(head_dim) ? 1 : 0; }

extern "C" int bl_seq_attention_fwd(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                                    const float* vbias, const int32_t* row_ptr, const int32_t* row_key,
                                    const int32_t* row_tab, int32_t B, int32_t H, int32_t L, int32_t D, int32_t T2, float p_drop,
                                    uint64_t seed, float* out, float* lse, bl_stream_t stream_) {
    using namespace bl;
    if (!q || !k || !v || !lengths || !bias || !row_ptr || !out || !lse) return BL_ERR_INVALID_ARGUMENT;
    if (B < 0 || H <= 0 || L <= 0 || T2 <= 0 || !supported_head_dim(D)) return BL_ERR_UNSUPPORTED;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return BL_ERR_INVALID_ARGUMENT;
    const int64_t rows = (int64_t)B * H * L;
    if (rows == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const seqatt::Problem p = make_problem(q, k, v, lengths, bias, vbias, row_ptr, row_key, row_tab, nullptr, nullptr, nullptr,
                                           B, H, L, T2, p_drop, seed);
    const unsigned grid = grid_for(rows, 128);
    switch (D) {
        case 8: seq_attention_fwd_kernel<8><<<grid, 128, 0, stream>>>(p, out, lse); break;
        case 16: seq_attention_fwd_kernel<16><<<grid, 128, 0, stream>>>(p, out, lse); break;
        case 32: seq_attention_fwd_kernel<32><<<grid, 128, 0, stream>>>(p, out, lse); break;
        default: seq_attention_fwd_kernel<64><<<grid, 128, 0, stream>>>(p, out, lse); break;
    }
    return check_launch("bl_seq_attention_fwd");
}

extern "C" int bl_seq_attention_bwd(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                                    const float* vbias, const int32_t* row_ptr, const int32_t* row_key,
                                    const int32_t* row_tab, const int32_t* col_ptr, const int32_t* col_query,
                                    const int32_t* col_tab, int32_t B, int32_t H, int32_t L, int32_t D, int32_t T2,
                                    float p_drop, uint64_t seed, const float* out, const float* lse, const float* d_out,
                                    float* dq, float* dk, float* dv, float* d_entry_bias, float* d_entry_vbias, float* delta,
                                    bl_stream_t stream_) {
    using namespace bl;
    if (!q || !k || !v || !lengths || !bias || !row_ptr || !col_ptr || !out || !lse || !d_out || !dq || !dk || !dv || !delta)
        return BL_ERR_INVALID_ARGUMENT;
    if ((vbias != nullptr) != (d_entry_vbias != nullptr)) return BL_ERR_INVALID_ARGUMENT;
    if (B < 0 || H <= 0 || L <= 0 || T2 <= 0 || !supported_head_dim(D)) return BL_ERR_UNSUPPORTED;
    const int64_t rows = (int64_t)B * H * L;
    if (rows == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return BL_ERR_INVALID_ARGUMENT;
    const seqatt::Problem p = make_problem(q, k, v, lengths, bias, vbias, row_ptr, row_key, row_tab, col_ptr, col_query, col_tab,
                                           B, H, L, T2, p_drop, seed);
    const unsigned grid = grid_for(rows, 128);
#define BL_LAUNCH_BWD(DD)                                                                                                   \
    seq_attention_bwd_row_kernel<DD><<<grid, 128, 0, stream>>>(p, out, lse, d_out, dq, d_entry_bias, d_entry_vbias, delta);  \
    seq_attention_bwd_col_kernel<DD><<<grid, 128, 0, stream>>>(p, lse, delta, d_out, dk, dv)
    switch (D) {
        case 8: BL_LAUNCH_BWD(8); break;
        case 16: BL_LAUNCH_BWD(16); break;
        case 32: BL_LAUNCH_BWD(32); break;
        default: BL_LAUNCH_BWD(64); break;
    }
#undef BL_LAUNCH_BWD
    return check_launch("bl_seq_attention_bwd");
}
