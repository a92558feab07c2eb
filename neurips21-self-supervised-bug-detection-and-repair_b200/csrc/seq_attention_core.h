// Edge-biased multi-head attention of the sequence models (seq-great / seq-rat), one (sample, head, row) at a time.
//
// Reference semantics: buglab/models/layers/multihead_attention.py:45-85 and relational_multihead_attention.py:71-178 with
// edge_attention_bias_is_scalar == False (the only mode the registry builds, seqmodel.py:93-107):
//     s[i, j]  = <q_i, k_j> + sum over entries e of row i with key j of <q_i, bias[tab_e]>          (q is pre-scaled)
//     p[i, :]  = softmax over the unmasked keys j < len
//     p'[i, j] = p[i, j] * keep(seed, b, h, i, j) / (1 - p_drop)      (dropout on the probabilities, multihead_attention.py:77)
//     o_i      = sum_j p'[i, j] * (v_j + sum over entries e of (i, j) of vbias[tab_e])               (vbias: "rat" only)
// An "entry" is one direction of one typed edge: edge (b, s, t, type) gives (row s, key t, table type) and
// (row t, key s, table T + type); repeated entries add up (index_put_(accumulate=True), :105-109, :172-176).
//
// Every function below computes ONE output row from read-only inputs and writes only locations it alone owns — no shared
// memory, no warp primitives, no atomics — so that exactly the same source runs (a) inside the CUDA kernels of
// seq_attention.cu, one thread per row, and (b) in plain loops on the host for the CPU emulation that pins the arithmetic
// and the index handling against oracle/seq_ref.py without a GPU (tests/test_seq_attention_emul.py).  This is the first
// correct path of SURVEY.md §8(f) row 2; a tcgen05 version of the two GEMM-shaped loops replaces it later.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BL_HD __host__ __device__ __forceinline__
#else
#define BL_HD inline
#endif

// Compiler-only barrier: rows read for a dot product are read AGAIN (cache hits) when they are needed for an update a few
// lines later, instead of being kept in 64 more registers across the exp / dropout code — at head size 64 the three
// per-thread vectors already take 192 registers.
#define BL_REREAD_ROWS() asm volatile("" ::: "memory")

namespace seqatt {

struct Problem {
    int B, H, L, T2;            // samples, heads, padded length, 2 * relation kinds (rows of the bias tables)
    const float* q;             // [B, H, L, D] (already multiplied by D^-0.5)
    const float* k;             // [B, H, L, D]
    const float* v;             // [B, H, L, D]
    const int32_t* lengths;     // [B] keys j >= lengths[b] are padding
    const float* bias;          // [T2, H, D]
    const float* vbias;         // [T2, H, D] or nullptr
    // entries grouped by query row: row_ptr[b * L + i] .. row_ptr[b * L + i + 1], keys ascending
    const int32_t* row_ptr;
    const int32_t* row_key;
    const int32_t* row_tab;
    // the same entries grouped by key: col_ptr[b * L + j] .., queries ascending; col_entry = position in the row order
    const int32_t* col_ptr;
    const int32_t* col_query;
    const int32_t* col_tab;
    float p_drop;               // dropout probability on the attention probabilities (0 = off)
    uint64_t seed;              // counter-based mask: element (b, h, i, j) is kept iff u(seed, index) >= p_drop
};

// Same counter-based generator as the other kernels (common.cuh hash_u32 / keep_element): splitmix64 finaliser.
BL_HD float dropout_scale(const Problem& p, int b, int h, int i, int j) {
    if (p.p_drop <= 0.f) return 1.f;
    const uint64_t idx = (((uint64_t)b * p.H + h) * p.L + i) * p.L + j;
    uint64_t z = p.seed + idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const float u = (float)((uint32_t)(z >> 32) >> 8) * (1.0f / 16777216.0f);
    return u >= p.p_drop ? 1.f / (1.f - p.p_drop) : 0.f;
}

// Rows are D * 4 bytes with D a multiple of 4 and come from 256-byte aligned allocations, so every row can be read in
// 16-byte pieces (LDG.128 on the device; the row pointers are uniform across a warp -> one broadcast transaction).
struct alignas(16) F4 { float x, y, z, w; };
BL_HD F4 load4(const float* p) { return *reinterpret_cast<const F4*>(p); }

// <a, b> with `a` in registers (or anywhere) and `b` a row in memory
template <int D>
BL_HD float dot(const float* a, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
        const F4 v = load4(b + d);
        s = fmaf(a[d], v.x, s);
        s = fmaf(a[d + 1], v.y, s);
        s = fmaf(a[d + 2], v.z, s);
        s = fmaf(a[d + 3], v.w, s);
    }
    return s;
}

// <a, b> with both rows in memory
template <int D>
BL_HD float dot_rows(const float* a, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
        const F4 u = load4(a + d), v = load4(b + d);
        s = fmaf(u.x, v.x, s);
        s = fmaf(u.y, v.y, s);
        s = fmaf(u.z, v.z, s);
        s = fmaf(u.w, v.w, s);
    }
    return s;
}

// acc = acc * keep + w * row
template <int D>
BL_HD void scale_add(float* acc, float keep, float w, const float* row) {
#pragma unroll
    for (int d = 0; d < D; d += 4) {
        const F4 v = load4(row + d);
        acc[d] = fmaf(w, v.x, acc[d] * keep);
        acc[d + 1] = fmaf(w, v.y, acc[d + 1] * keep);
        acc[d + 2] = fmaf(w, v.z, acc[d + 2] * keep);
        acc[d + 3] = fmaf(w, v.w, acc[d + 3] * keep);
    }
}

// acc += w * row
template <int D>
BL_HD void axpy(float* acc, float w, const float* row) {
#pragma unroll
    for (int d = 0; d < D; d += 4) {
        const F4 v = load4(row + d);
        acc[d] = fmaf(w, v.x, acc[d]);
        acc[d + 1] = fmaf(w, v.y, acc[d + 1]);
        acc[d + 2] = fmaf(w, v.z, acc[d + 2]);
        acc[d + 3] = fmaf(w, v.w, acc[d + 3]);
    }
}

// ---- forward: out[b, h, i, :], lse[b, h, i] ------------------------------------------------------------------------
template <int D>
BL_HD void forward_row(const Problem& p, int b, int h, int i, float* out, float* lse) {
    const int len = p.lengths[b];
    const size_t head = ((size_t)b * p.H + h) * p.L;
    float* o = out + (head + i) * D;
    if (i >= len) {  // padding query: never read downstream
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = 0.f;
        lse[head + i] = 0.f;
        return;
    }
    float q[D], acc[D];
    const float* qrow = p.q + (head + i) * D;
#pragma unroll
    for (int d = 0; d < D; ++d) { q[d] = qrow[d]; acc[d] = 0.f; }
    int e = p.row_ptr[b * p.L + i];
    const int e_end = p.row_ptr[b * p.L + i + 1];
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < len; ++j) {
        float s = dot<D>(q, p.k + (head + j) * D);
        const int e_first = e;
        while (e < e_end && p.row_key[e] == j) {
            s += dot<D>(q, p.bias + ((size_t)p.row_tab[e] * p.H + h) * D);
            ++e;
        }
        const float m_new = fmaxf(m, s);
        const float rescale = expf(m - m_new);   // 0 on the first key (m = -inf)
        const float w = expf(s - m_new);
        l = l * rescale + w;
        const float wm = w * dropout_scale(p, b, h, i, j);   // the normaliser l sees the undropped weight
        scale_add<D>(acc, rescale, wm, p.v + (head + j) * D);
        if (p.vbias != nullptr) {
            for (int f = e_first; f < e; ++f) axpy<D>(acc, wm, p.vbias + ((size_t)p.row_tab[f] * p.H + h) * D);
        }
        m = m_new;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = acc[d] * inv;
    lse[head + i] = m + logf(l);
}

// ---- backward, query side: dq[b, h, i, :] and the per-entry table gradients ------------------------------------------
// d_entry_bias[e, h, :]  = dS[i, key_e] * q_i          (summed into the bias table by the caller: index_add over tab_e)
// d_entry_vbias[e, h, :] = p[i, key_e] * dO_i
template <int D>
BL_HD void backward_row(const Problem& p, const float* out, const float* lse, const float* d_out, int b, int h, int i,
                        float* dq, float* d_entry_bias, float* d_entry_vbias, float* delta_out) {
    const int len = p.lengths[b];
    const size_t head = ((size_t)b * p.H + h) * p.L;
    float* dqrow = dq + (head + i) * D;
    int e = p.row_ptr[b * p.L + i];
    const int e_end = p.row_ptr[b * p.L + i + 1];
    if (i >= len) {
        delta_out[head + i] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) dqrow[d] = 0.f;
        for (; e < e_end; ++e) {  // entries of padding rows cannot exist for well-formed inputs; keep their slots defined
            for (int d = 0; d < D; ++d) {
                d_entry_bias[((size_t)e * p.H + h) * D + d] = 0.f;
                if (d_entry_vbias != nullptr) d_entry_vbias[((size_t)e * p.H + h) * D + d] = 0.f;
            }
        }
        return;
    }
    float q[D], g[D], acc[D];
    const float* qrow = p.q + (head + i) * D;
    const float* grow = d_out + (head + i) * D;
    const float* orow = out + (head + i) * D;
    float delta = 0.f;  // sum_j p_ij dP_ij = <dO_i, O_i>
#pragma unroll
    for (int d = 0; d < D; ++d) { q[d] = qrow[d]; g[d] = grow[d]; acc[d] = 0.f; delta = fmaf(g[d], orow[d], delta); }
    delta_out[head + i] = delta;   // read by backward_col (launched after this kernel)
    const float row_lse = lse[head + i];
    for (int j = 0; j < len; ++j) {
        float s = dot<D>(q, p.k + (head + j) * D);
        float dp = dot<D>(g, p.v + (head + j) * D);
        const int e_first = e;
        while (e < e_end && p.row_key[e] == j) {
            s += dot<D>(q, p.bias + ((size_t)p.row_tab[e] * p.H + h) * D);
            if (p.vbias != nullptr) dp += dot<D>(g, p.vbias + ((size_t)p.row_tab[e] * p.H + h) * D);
            ++e;
        }
        const float prob = expf(s - row_lse);
        const float mask = dropout_scale(p, b, h, i, j);
        const float ds = prob * (mask * dp - delta);
        BL_REREAD_ROWS();
        axpy<D>(acc, ds, p.k + (head + j) * D);
        for (int f = e_first; f < e; ++f) {
            axpy<D>(acc, ds, p.bias + ((size_t)p.row_tab[f] * p.H + h) * D);
            float* db = d_entry_bias + ((size_t)f * p.H + h) * D;
#pragma unroll
            for (int d = 0; d < D; ++d) db[d] = ds * q[d];
            if (d_entry_vbias != nullptr) {
                float* dvb = d_entry_vbias + ((size_t)f * p.H + h) * D;
#pragma unroll
                for (int d = 0; d < D; ++d) dvb[d] = prob * mask * g[d];
            }
        }
    }
    // entries whose key is padding (>= len) get no probability mass
    for (; e < e_end; ++e) {
        for (int d = 0; d < D; ++d) {
            d_entry_bias[((size_t)e * p.H + h) * D + d] = 0.f;
            if (d_entry_vbias != nullptr) d_entry_vbias[((size_t)e * p.H + h) * D + d] = 0.f;
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) dqrow[d] = acc[d];
}

// ---- backward, key side: dk[b, h, j, :], dv[b, h, j, :]; delta[b, h, i] = <dO_i, O_i> comes from backward_row ----------
template <int D>
BL_HD void backward_col(const Problem& p, const float* lse, const float* delta, const float* d_out, int b, int h, int j,
                        float* dk, float* dv) {
    const int len = p.lengths[b];
    const size_t head = ((size_t)b * p.H + h) * p.L;
    float* dkrow = dk + (head + j) * D;
    float* dvrow = dv + (head + j) * D;
    if (j >= len) {
#pragma unroll
        for (int d = 0; d < D; ++d) { dkrow[d] = 0.f; dvrow[d] = 0.f; }
        return;
    }
    // Two sweeps over the queries, one per output, so that each holds only k_j and ONE accumulator in registers (at head
    // size 64 a single sweep with both accumulators spills); the score is recomputed in the second sweep.
    float kk[D], acc[D];
    const float* krow = p.k + (head + j) * D;
    const float* vrow = p.v + (head + j) * D;
#pragma unroll
    for (int d = 0; d < D; ++d) kk[d] = krow[d];
    const int e_begin = p.col_ptr[b * p.L + j];
    const int e_end = p.col_ptr[b * p.L + j + 1];

    // sweep 1: dK_j = sum_i dS_ij q_i
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    int e = e_begin;
    for (int i = 0; i < len; ++i) {
        const float* qrow = p.q + (head + i) * D;
        const float* grow = d_out + (head + i) * D;
        float s = dot<D>(kk, qrow);
        float dp = dot_rows<D>(vrow, grow);
        while (e < e_end && p.col_query[e] == i) {
            s += dot_rows<D>(p.bias + ((size_t)p.col_tab[e] * p.H + h) * D, qrow);
            if (p.vbias != nullptr) dp += dot_rows<D>(p.vbias + ((size_t)p.col_tab[e] * p.H + h) * D, grow);
            ++e;
        }
        const float prob = expf(s - lse[head + i]);
        const float ds = prob * (dropout_scale(p, b, h, i, j) * dp - delta[head + i]);
        BL_REREAD_ROWS();
        axpy<D>(acc, ds, qrow);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) dkrow[d] = acc[d];

    // sweep 2: dV_j = sum_i p'_ij dO_i
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    e = e_begin;
    for (int i = 0; i < len; ++i) {
        const float* qrow = p.q + (head + i) * D;
        float s = dot<D>(kk, qrow);
        while (e < e_end && p.col_query[e] == i) {
            s += dot_rows<D>(p.bias + ((size_t)p.col_tab[e] * p.H + h) * D, qrow);
            ++e;
        }
        const float weight = expf(s - lse[head + i]) * dropout_scale(p, b, h, i, j);
        axpy<D>(acc, weight, d_out + (head + i) * D);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) dvrow[d] = acc[d];
}

}  // namespace seqatt
