// Host emulation of the seq-attention kernels: the SAME row functions (csrc/seq_attention_core.h) driven by plain loops
// instead of one CUDA thread per row.  Test infrastructure only (tests/test_seq_attention_emul.py builds it with g++); it
// lets the arithmetic and index handling of the kernels be checked against the oracle on a machine without a GPU.
#include "../../neurips21-self-supervised-bug-detection-and-repair_b200/csrc/seq_attention_core.h"

namespace {

seqatt::Problem make(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                     const float* vbias, const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab,
                     const int32_t* col_ptr, const int32_t* col_query, const int32_t* col_tab, int B, int H, int L, int T2,
                     float p_drop, uint64_t seed) {
    seqatt::Problem p;
    p.p_drop = p_drop; p.seed = seed;
    p.B = B; p.H = H; p.L = L; p.T2 = T2;
    p.q = q; p.k = k; p.v = v; p.lengths = lengths; p.bias = bias; p.vbias = vbias;
    p.row_ptr = row_ptr; p.row_key = row_key; p.row_tab = row_tab;
    p.col_ptr = col_ptr; p.col_query = col_query; p.col_tab = col_tab;
    return p;
}

template <int D>
void fwd(const seqatt::Problem& p, float* out, float* lse) {
    for (int b = 0; b < p.B; ++b)
        for (int h = 0; h < p.H; ++h)
            for (int i = 0; i < p.L; ++i) seqatt::forward_row<D>(p, b, h, i, out, lse);
}

template <int D>
void bwd(const seqatt::Problem& p, const float* out, const float* lse, const float* d_out, float* dq, float* dk, float* dv,
         float* d_entry_bias, float* d_entry_vbias, float* delta) {
    for (int b = 0; b < p.B; ++b)
        for (int h = 0; h < p.H; ++h)
            for (int i = 0; i < p.L; ++i)
                seqatt::backward_row<D>(p, out, lse, d_out, b, h, i, dq, d_entry_bias, d_entry_vbias, delta);
    for (int b = 0; b < p.B; ++b)
        for (int h = 0; h < p.H; ++h)
            for (int j = 0; j < p.L; ++j) seqatt::backward_col<D>(p, lse, delta, d_out, b, h, j, dk, dv);
}

}  // namespace

extern "C" int emul_seq_attention_fwd(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                                      const float* vbias, const int32_t* row_ptr, const int32_t* row_key,
                                      const int32_t* row_tab, int32_t B, int32_t H, int32_t L, int32_t D, int32_t T2, float p_drop,
                                      uint64_t seed, float* out, float* lse) {
    const seqatt::Problem p = make(q, k, v, lengths, bias, vbias, row_ptr, row_key, row_tab, nullptr, nullptr, nullptr, B, H, L, T2,
                                   p_drop, seed);
    switch (D) {
        case 8: fwd<8>(p, out, lse); return 0;
        case 16: fwd<16>(p, out, lse); return 0;
        case 32: fwd<32>(p, out, lse); return 0;
        case 64: fwd<64>(p, out, lse); return 0;
        default: return -4;
    }
}

extern "C" int emul_seq_attention_bwd(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                                      const float* vbias, const int32_t* row_ptr, const int32_t* row_key,
                                      const int32_t* row_tab, const int32_t* col_ptr, const int32_t* col_query,
                                      const int32_t* col_tab, int32_t B, int32_t H, int32_t L, int32_t D, int32_t T2,
                                      float p_drop, uint64_t seed, const float* out, const float* lse, const float* d_out,
                                      float* dq, float* dk, float* dv, float* d_entry_bias, float* d_entry_vbias, float* delta) {
    const seqatt::Problem p = make(q, k, v, lengths, bias, vbias, row_ptr, row_key, row_tab, col_ptr, col_query, col_tab, B, H, L, T2,
                                   p_drop, seed);
    switch (D) {
        case 8: bwd<8>(p, out, lse, d_out, dq, dk, dv, d_entry_bias, d_entry_vbias, delta); return 0;
        case 16: bwd<16>(p, out, lse, d_out, dq, dk, dv, d_entry_bias, d_entry_vbias, delta); return 0;
        case 32: bwd<32>(p, out, lse, d_out, dq, dk, dv, d_entry_bias, d_entry_vbias, delta); return 0;
        case 64: bwd<64>(p, out, lse, d_out, dq, dk, dv, d_entry_bias, d_entry_vbias, delta); return 0;
        default: return -4;
    }
}
