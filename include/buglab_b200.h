/*
 * buglab_b200.h — C ABI of the B200-native gnn-mlp message-passing hot path.
 *
 * The reference (microsoft/neurips21-self-supervised-bug-detection-and-repair) is 100 % Python and
 * has no FFI boundary of its own; the de-facto plugin boundary is the Python operator surface that
 * buglab/models imports from `ptgnn` and `torch_scatter`.  Every entry point below
 * names the reference call site whose arithmetic it replaces.  Conventions:
 *
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - row-major contiguous fp32 matrices, int32 indices (the reference's int64 index tensors,
 *     buglab/models/gnn.py:551-596, are narrowed once per minibatch by the caller);
 *   - the caller owns and pre-allocates every buffer (outputs and workspace); no allocation inside;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises;
 *   - return value 0 on success, negative BL_ERR_* otherwise (bl_error_string() explains);
 *     no exceptions cross the boundary.
 */
#ifndef BUGLAB_B200_H
#define BUGLAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BL_OK 0
#define BL_ERR_INVALID_ARGUMENT (-1)
#define BL_ERR_CUDA (-2)
#define BL_ERR_WORKSPACE_TOO_SMALL (-3)
#define BL_ERR_UNSUPPORTED (-4)

typedef void* bl_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch). */
int bl_version(void);
/* Human-readable text for a BL_ERR_* code; for BL_ERR_CUDA includes the last CUDA error seen. */
const char* bl_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * Typed-edge plan (CSR by target + (node,type) pair tables).
 *
 * Replaces the per-layer index handling of ptgnn's MlpMessagePassingLayer.forward
 * (`torch.cat([adj[1] for adj in adjacency_lists])`, per-type `index_select`; reference call site
 * buglab/models/gnnlayerdefs.py:6-23 via buglab/models/gnn.py:117) and the edge-typed adjacency
 * produced by buglab/representations/data.py:139-167.  Built ONCE per minibatch, reused by all 8
 * message-passing layers, forward and backward.
 *
 * Input: the type-major concatenation of the adjacency lists: src/tgt/etype[e], e in [0,E), where
 * edges of type 0 come first, then type 1, ... (exactly ptgnn's `cat` order, so "original edge
 * index" == position in this concatenation).
 *
 * Output (all int32, caller-allocated):
 *   e_perm[E]      original edge index of sorted edge i; sorted by (tgt, type, original index)
 *   e_src[E], e_type[E]   source node / type of sorted edge i
 *   row_ptr[N+1]   sorted edges of target n are [row_ptr[n], row_ptr[n+1])
 *   urow[E]        id of the unique (type, src) pair of sorted edge i   (row of the U table)
 *   vrow[E]        id of the unique (type, tgt) pair of sorted edge i   (row of the V table)
 *   s_node[E]      node of S-pair p (p < P_s); pairs are ordered by (type, node)
 *   s_type_ptr[K+1]  S-pairs of type k are [s_type_ptr[k], s_type_ptr[k+1])
 *   s_by_node_ptr[N+1], s_by_node_idx[E]   CSR node -> S-pair ids (ascending pair id per node)
 *   t_*            same four tables for the T-pairs (type, tgt)
 *   counts[2]      {P_s, P_t}
 *   block_nodes    0: pairs ordered by (type, node) — s_type_ptr / t_type_ptr have K+1 entries, one segment per type.
 *                  B > 0: pairs ordered by (node / B, type, node) — the two pointer arrays have ceil(N/B)*K + 1 entries,
 *                  segment s = (node block s / K, type s % K); only the segment-aware GEMMs (bl_tma_*) accept this
 *                  layout (they then sweep the node states once per layer instead of once per type).
 *   s_edge_ptr[E+1], s_edge_idx[E]  (nullable) CSR S-pair -> its sorted edges (ascending): the edges whose U row
 *                  is that pair; e_tgt[E] (nullable) target node of every sorted edge; s_edge_tgt[E] (nullable) =
 *                  e_tgt[s_edge_idx[.]], the same targets in S-pair order.  They feed the by-source half
 *                  of the edge kernel's backward (bl_edge_bwd_sources).
 * ------------------------------------------------------------------------------------------------ */
size_t bl_plan_workspace_bytes(int64_t num_edges, int64_t num_nodes, int32_t num_edge_types);

int bl_plan_build(const int32_t* src, const int32_t* tgt, const int32_t* etype,
                  int64_t num_edges, int64_t num_nodes, int32_t num_edge_types,
                  int32_t* e_perm, int32_t* e_src, int32_t* e_type, int32_t* row_ptr,
                  int32_t* urow, int32_t* vrow,
                  int32_t* s_node, int32_t* s_type_ptr, int32_t* s_by_node_ptr, int32_t* s_by_node_idx,
                  int32_t* t_node, int32_t* t_type_ptr, int32_t* t_by_node_ptr, int32_t* t_by_node_idx,
                  int32_t* counts, int32_t* s_edge_ptr, int32_t* s_edge_idx, int32_t* e_tgt, int32_t* s_edge_tgt,
                  int32_t block_nodes,
                  void* workspace, size_t workspace_bytes, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row gather / segmented row sum — the `index_select h[src]`, `h[tgt]` of MlpMessagePassingLayer
 * (P4 in SURVEY.md §8a) restricted to unique (type,node) pairs, and its backward (the scatter-add
 * of the gathered-row gradients), done as a CSR segmented sum: deterministic, no atomics.
 * ------------------------------------------------------------------------------------------------ */
/* out[p,:] = table[idx[p],:]   (p < num_rows; dim % 4 == 0) */
int bl_rows_gather(const float* table, const int32_t* idx, int64_t num_rows, int32_t dim,
                   float* out, bl_stream_t stream);

/* out[n,:] (+)= sum_{q in [a_ptr[n],a_ptr[n+1])} a_rows[a_idx[q],:] + sum_{q in b range} b_rows[b_idx[q],:]
 * b_* may be NULL.  accumulate != 0 adds to the existing contents of out.  amax (device scalar, may be NULL): the rows
 * come from a table that was pre-scaled by the power of two derived from *amax (see bl_rows_split3_f16); the sum is
 * multiplied by the exact inverse. */
int bl_rows_segment_sum(const float* a_rows, const int32_t* a_ptr, const int32_t* a_idx,
                        const float* b_rows, const int32_t* b_ptr, const int32_t* b_idx,
                        int64_t num_nodes, int32_t dim, int32_t accumulate, const float* amax, float* out,
                        bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pair projections on the tensor cores (split-fp16, fp32-class accuracy) — the hoisted per-type affine
 *     U[p] = A_k h[s_node[p]],   V[p] = B_k h[t_node[p]] + b_k      with  Linear_k.weight = [A_k | B_k]
 * of ptgnn MlpMessagePassingLayer (reference call site buglab/models/gnnlayerdefs.py:6-23) and its backward.
 * fp32 operands are split x = x1 + x2 (fp16 parts, 22 bits together) and x.w is evaluated as x1.w1 + x1.w2 + x2.w1 by
 * ONE fp16 GEMM over the concatenated reduction [x1|x1|x2].[w1|w2|w1] with fp32 accumulation (max error ~2e-6 at D=256,
 * like fp32 SGEMM; a single TF32/BF16 pass and even a bf16 split miss the 1e-4 parity budget over 8 layers).
 * Row layout of every split table: 3*dim + 8 fp16 per row (the 8 trailing columns carry [1 1 1 0 0 0 0 0] so that a
 * bias b = b1+b2+b3 folds into the same reduction).  Gradient tables are pre-scaled by an exact power of two derived
 * from a device-side absolute maximum (`amax`) so that fp16's narrow exponent range costs no accuracy.
 * type_ptr_host[K+1] are HOST arrays (they size the per-type GEMMs).
 * ------------------------------------------------------------------------------------------------ */
/* out[r,:] = [hi | hi | lo | 1 1 1 0 0 0 0 0] of s*table[idx[r],:] (idx may be NULL: r itself); out is fp16
 * [num_rows, 3*dim+8]; s = 2^(12-ceil(log2(*amax))) if amax != NULL (device scalar), else 1. */
int bl_rows_split3_f16(const float* table, const int32_t* idx, int64_t num_rows, int32_t dim, const float* amax,
                       void* out, bl_stream_t stream);
/* x[i] /= s  with the same s as above (undoes the pre-scale on a small tensor, e.g. the weight gradient) */
int bl_unscale_pow2(float* x, int64_t n, const float* amax, bl_stream_t stream);
/* From weight[K, out_dim, ld] columns [col0, col0+in_dim) (+ bias[K,out_dim] or NULL):
 *   w3_fwd [K, out_dim, 3*in_dim+8] = [w1 | w2 | w1 | b1 b2 b3 0...]      (forward, "NT" operand)
 *   b3_bwd [K, 3*out_dim, in_dim]   = [w1 ; w2 ; w1] stacked along rows     (backward w.r.t. the input rows)
 * either output may be NULL. */
int bl_weights_split3_f16(const float* weight, const float* bias, int32_t num_types, int32_t out_dim, int32_t in_dim,
                           int32_t ld, int32_t col0, void* w3_fwd, void* b3_bwd, bl_stream_t stream);
/* out[rows of type k, 0:out_dim] = a3[rows] . w3_fwd[k]^T   (fp32 out) */
int bl_pair_project_fwd(const void* a3, const void* w3_fwd, const int32_t* type_ptr_host, int32_t num_types,
                        int32_t out_dim, int32_t in_dim, float* out, bl_stream_t stream);
/* d_rows[rows of type k, 0:in_dim] = g3[rows, 0:3*out_dim] . b3_bwd[k]   (g3 = bl_rows_split3_f16 of the table gradient) */
int bl_pair_project_bwd_input(const void* g3, const void* b3_bwd, const int32_t* type_ptr_host, int32_t num_types,
                              int32_t out_dim, int32_t in_dim, float* d_rows, bl_stream_t stream);
/* out[r,:] = [hi | lo] of s*table[idx[r],:]  (fp16 [num_rows, 2*dim]) — the operand of the weight-gradient GEMM */
int bl_rows_split2_f16(const float* table, const int32_t* idx, int64_t num_rows, int32_t dim, const float* amax,
                       void* out, bl_stream_t stream);
/* d_weight[k, 0:out_dim, col0:col0+in_dim] = (1/s) * sum over rows of type k of g^T h.  One fp16 GEMM per type forms all
 * four hi/lo cross products T_k[2*out_dim, 2*in_dim] = [g1|g2]^T.[h1|h2] into tmp (fp32, [num_types, 2*out_dim,
 * 2*in_dim]); a fold kernel adds the blocks and undoes the pow2 pre-scale of g (amax may be NULL: s = 1).
 * g holds [g1|g2] at columns [g_col0, g_col0 + 2*out_dim) of a split table with row stride g_stride (three-part table:
 * g_stride = 3*out_dim+8, g_col0 = out_dim; two-part table: g_stride = 2*out_dim, g_col0 = 0);
 * a2 = bl_rows_split2_f16 of the gathered inputs. */
int bl_pair_project_bwd_weight(const void* g, int32_t g_stride, int32_t g_col0, const void* a2,
                               const int32_t* type_ptr_host, int32_t num_types, int32_t out_dim, int32_t in_dim,
                               const float* amax, float* tmp, float* d_weight, int32_t ld, int32_t col0,
                               bl_stream_t stream);
/* out[k, 0:dim] = sum of rows[type_ptr[k] : type_ptr[k+1], :]   (type_ptr on the DEVICE; the bias gradient; dim <= 1024) */
int bl_grouped_colsum(const float* rows, const int32_t* type_ptr, int32_t num_types, int32_t dim, float* out,
                      bl_stream_t stream);

/* amax[0] = max_i |x[i]|  (n % 4 == 0) — the device-side scale source for an fp16 split of a gradient tensor */
int bl_absmax(const float* x, int64_t n, float* amax, bl_stream_t stream);

/* Hand-written tcgen05/TMEM version of the pair projection: gathers + splits the fp32 rows inside the GEMM loader, so
 * the split table never exists in HBM.   out[p, 0:n_out] = s * src[idx[p], 0:k_in] . W_k^T (+ bias_k),  p in type k.
 *   parts  = bl_weight_parts_f16 output [num_types, 2 (hi,lo), n_out, k_in] fp16
 *   idx    may be NULL (identity), amax may be NULL (s = 1), bias may be NULL; type_ptr is a DEVICE array.
 * Supported shapes: k_in % 64 == 0 and n_out in {128, 256, 512, 768, 1024} (bl_pair_project_tc_supported).
 * bl_weight_parts_f16: amax (nullable) = device scalar max|weight|: the parts are pre-scaled by the power of two derived
 * from it (only the bl_tma_* kernels undo it, through their amax_b argument; pass NULL for bl_pair_project_tc). */
int bl_weight_parts_f16(const float* weight, int32_t num_types, int32_t n_out, int32_t k_in, int32_t ld, int32_t col0,
                        int32_t transposed, const float* amax, void* parts, bl_stream_t stream);
int bl_pair_project_tc_supported(int32_t n_out, int32_t k_in);
int bl_pair_project_tc(const float* src, const int32_t* idx, const float* amax, const void* parts, const float* bias,
                       const int32_t* type_ptr, int32_t num_types, int64_t num_rows, int32_t n_out, int32_t k_in,
                       float* out, bl_stream_t stream);

/* Hand-written tcgen05/TMEM weight gradient of the pair projection (no split tables in HBM):
 *   d_weight[k, 0:m_out, col0:col0+n_in] = (1/s) * sum over pair rows p of type k of (s*g[p,:])^T x[idx[p],:]
 * g = the table gradient [P, m_out] (dU or dV), x = node states [*, n_in], s from *amax (NULL: 1).  The destination
 * block is zero-filled here and accumulated with fp32 REDs (one partial per 4096-row slab).
 * Supported: m_out % 128 == 0, n_in % 256 == 0, both <= 1024 (bl_pair_weight_grad_tc_supported). */
int bl_pair_weight_grad_tc_supported(int32_t m_out, int32_t n_in);
int bl_pair_weight_grad_tc(const float* g, const float* x, const int32_t* idx, const float* amax,
                           const int32_t* type_ptr, int32_t num_types, int64_t num_rows, int32_t m_out, int32_t n_in,
                           float* d_weight, int32_t ld, int32_t col0, bl_stream_t stream);

/* Backward of the fused edge kernel in the form the second-generation GEMMs consume (deterministic: every output row
 * has one writer, no memset, no atomics on the tables).  g[n,c] = d_agg[n,c] * GELU'(xwin[n,c]) belongs to the winning
 * edge ewin[n,c] only (arg-routed backward of scatter_max).
 *   bl_edge_bwd_targets: one warp per target node.  Writes g_rows[N, M] (fp32) and dV as an fp16 hi/lo split table
 *     [2][P_t + 1][M] (pre-scaled by the power of two derived from amax_eff, last row of each part zero); adds the
 *     per-type column sums of dV (= d bias_k) into d_bias[K, M] (zero-initialised by the caller; NULL: skipped).
 *     amax_in = device scalar max|d_agg| (bl_absmax); the kernel publishes amax_eff = amax_in * 1.13 (bound of GELU') *
 *     256 (fan-in headroom of the dU sums) for the consumers' 1/scale.
 *   bl_edge_bwd_sources: one warp per S-pair row.  dU[p] = sum over the pair's edges e (ascending) of g masked to the
 *     channels e won, written as the split table [2][P_s + 1][M] with the same scale. */
int bl_edge_bwd_targets(const float* d_agg, const float* xwin, const int32_t* ewin, const int32_t* row_ptr,
                        const int32_t* vrow, const int32_t* e_type, int64_t num_nodes, int32_t msg_dim,
                        int32_t num_edge_types, int64_t num_t_pairs, const float* amax_in, float* amax_eff, float* g_rows,
                        void* dv_split, float* d_bias, bl_stream_t stream);
int bl_edge_bwd_sources(const float* g_rows, const int32_t* ewin, const int32_t* s_edge_ptr, const int32_t* s_edge_idx,
                        const int32_t* s_edge_tgt, int64_t num_s_pairs, int32_t msg_dim, const float* amax_eff, void* du_split,
                        bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Second-generation projection GEMMs: TMA-fed tcgen05 (csrc/gemm_tma.cu).  Same reference call sites as
 * bl_pair_project_tc / bl_pair_weight_grad_tc (per-type Linear_k of ptgnn's MlpMessagePassingLayer hoisted
 * to unique (type, node) pairs, buglab/models/gnnlayerdefs.py:6-23; its two backward products; the layer's
 * node-update Linear(M -> D_out)).  Operands are split into fp16 hi/lo parts ONCE per table
 * (bl_rows_split_f16) and moved by the TMA engine (row gathers: tile::gather4); CTA pairs run
 * tcgen05.mma.cta_group::2 on 256-row tiles unless BUGLAB_B200_TMA_CG=1.
 *
 * Split table layout: fp16 [2][rows + 1][dim] — part 0 = fp16(s*x), part 1 = fp16(s*x - part 0); the extra
 * last row of each part is zero (the padding row the kernels read for rows past a segment end).
 * s = pow2 pre-scale derived from *amax (NULL: 1).
 *
 * Segments and work units: pair rows are grouped in num_segs ranges [seg_ptr[s], seg_ptr[s+1]) that share weight matrix
 * seg_type[s] (NULL: s).  The GEMMs walk a device-side table of work units built once per plan by bl_segment_units:
 * units[u] = int32 {first row, end row, weight matrix, segment} for every run of <= `unit` consecutive pair rows of one
 * segment (unit = bl_tma_tile_rows() for the projections, bl_tma_slab_rows() for the weight gradient and the
 * weight-stationary projection); *count = number of units.  max_units (>= num_rows / unit + num_segs) sizes the table and
 * the persistent grid without a device->host copy. */
int bl_rows_split_f16(const float* x, const int32_t* idx, int64_t rows, int32_t dim, const float* amax, void* out,
                      bl_stream_t stream);
int bl_segment_unit_prefix(const int32_t* seg_ptr, int32_t num_segs, int32_t unit, int32_t* prefix, bl_stream_t stream);
int bl_segment_units(const int32_t* seg_ptr, const int32_t* seg_type, int32_t num_segs, int32_t unit, int64_t max_units,
                     int32_t* prefix_ws, void* units, int32_t* count, bl_stream_t stream);
int bl_tma_tile_rows(void);
int bl_tma_slab_rows(void);
int bl_tma_gemm_supported(int32_t n_out, int32_t k_in);
/* out[p, 0:n_out] = (1/(s_a*s_b)) * A[row(p), :] . W_type[0:n_out, 0:k_in]^T (+ bias_type);  row(p) = idx[p] or p (idx
 * NULL); a_rows = rows per part of a_split (incl. the zero row); wparts as written by bl_weight_parts_f16; s_a / s_b =
 * the power-of-two pre-scales of the A table / the weight parts, from *amax / *amax_b (NULL: 1).  Pre-scaling BOTH
 * operands keeps their fp16 lo parts out of the subnormal range.  tiles / num_tiles: bl_segment_units
 * with unit bl_tma_tile_rows(). */
int bl_tma_project(const void* a_split, int64_t a_rows, const int32_t* idx, const void* wparts, const float* bias,
                   const float* amax, const float* amax_b, const void* tiles, const int32_t* num_tiles,
                   int32_t num_types, int64_t num_rows, int64_t max_tiles, int32_t n_out, int32_t k_in,
                   float* out, bl_stream_t stream);
/* Weight-stationary variant for the 256 x 256 products (CTA pairs): the pair keeps a slab's weight matrix (hi and lo, its
 * 128 output columns per CTA: 128 KB) in shared memory and streams only A.  slabs / num_slabs: bl_segment_units with unit
 * bl_tma_slab_rows().  Opt-in (BUGLAB_B200_TMA_BSTAT=1): the streaming kernel measured faster on B200. */
int bl_tma_project_stationary_supported(int32_t n_out, int32_t k_in);
int bl_tma_project_stationary(const void* a_split, int64_t a_rows, const int32_t* idx, const void* wparts, const float* bias,
                              const float* amax, const float* amax_b, const void* slabs, const int32_t* num_slabs,
                              int32_t num_types, int64_t num_rows, int64_t max_slabs, int32_t n_out,
                              int32_t k_in, float* out, bl_stream_t stream);
int bl_tma_weight_grad_supported(int32_t m_out, int32_t n_in);
/* d_weight[type, 0:m_out, col0:col0+n_in] = (1/(s_g*s_x)) * sum over pair rows of G[p, :]^T X[idx[p], :]  (block zeroed
 * first); s_g / s_x from *amax / *amax_x (NULL: 1). */
int bl_tma_weight_grad(const void* g_split, int64_t g_rows, const void* x_split, int64_t x_rows, const int32_t* idx,
                       const float* amax, const float* amax_x, const void* slabs, const int32_t* num_slabs,
                       int32_t num_types, int64_t num_rows, int64_t max_slabs, int32_t m_out,
                       int32_t n_in, float* d_weight, int32_t ld, int32_t col0, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused typed-edge message + aggregate (THE hot kernel).
 *
 * Replaces, per layer, ptgnn MlpMessagePassingLayer.forward's
 *     m = GELU(cat_k Linear_k(cat[h[src_k], h[tgt_k]]));  agg = torch_scatter.scatter_max(m, tgt, dim_size=N)[0]
 * (reference call sites buglab/models/gnnlayerdefs.py:6-23, aggregation "max" at :11,:20;
 *  torch_scatter semantics: empty segment -> 0, first maximum wins).
 *
 * With Linear_k([h_s;h_t]) = A_k h_s + B_k h_t + b_k hoisted to the unique (type,node) pairs,
 *     U[urow] = A_k h_src,  V[vrow] = B_k h_tgt + b_k   ([P_s,M] and [P_t,M] tables)
 * the pre-activation of sorted edge i is x_i = U[urow[i]] + V[vrow[i]].  GELU is quasi-convex
 * (decreasing left of x0 ~ -0.7518, increasing right of it), so max_i GELU(x_i) is attained at the
 * minimum or the maximum x_i of the segment; the kernel tracks both extremes per channel and
 * evaluates erf twice per (node, channel) instead of once per (edge, channel).
 *
 * Outputs: agg[N,M] aggregated messages (0 for nodes without in-edges), xwin[N,M] the winning
 * pre-activation, ewin[N,M] the winning SORTED edge index (-1 for empty) — both kept for backward.
 * M % 4 == 0 required.
 * ------------------------------------------------------------------------------------------------ */
int bl_edge_segmax_fwd(const float* u_rows, const float* v_rows,
                       const int32_t* row_ptr, const int32_t* urow, const int32_t* vrow,
                       int64_t num_nodes, int32_t msg_dim,
                       float* agg, float* xwin, int32_t* ewin, bl_stream_t stream);

/* Backward of the above: g = d_agg * GELU'(xwin) routed to the winning edge only (the arg-routing
 * backward of torch_scatter.scatter_max).  d_v_rows[P_t,M] is fully written (each (type,tgt) row has
 * exactly one owner node); d_u_rows[P_s,M] is zero-filled here and accumulated with fp32 REDs.  amax (device scalar,
 * may be NULL) receives 256 * max|g|: the scale source of the fp16 split of both tables (|d_v_rows| <= max|g|; a
 * d_u_rows entry sums the g of the targets sharing its pair — 256x headroom, and the split saturates, never infs). */
int bl_edge_segmax_bwd(const float* d_agg, const float* xwin, const int32_t* ewin,
                       const int32_t* row_ptr, const int32_t* urow, const int32_t* vrow,
                       int64_t num_nodes, int32_t msg_dim, int64_t num_s_pairs, int64_t num_t_pairs,
                       float* d_u_rows, float* d_v_rows, float* amax, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Node update  LayerNorm(M) -> [Linear(M->D_out) is a library/tensor-core GEMM] -> Tanh -> Dropout
 * (ptgnn MlpMessagePassingLayer state update).
 * ------------------------------------------------------------------------------------------------ */
/* y = (x-mean)*rstd*gamma+beta per row; saves mean[rows], rstd[rows]. dim % 4 == 0, eps as torch. */
int bl_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int32_t dim,
                     float eps, float* y, float* mean, float* rstd, bl_stream_t stream);
/* dx; d_gamma/d_beta accumulated into partial[2, num_partials, dim] then reduced by the same call
 * into d_gamma[dim], d_beta[dim].  partial must hold 2*BL_LN_PARTIALS*dim floats. */
#define BL_LN_PARTIALS 1184 /* 8 x 148 SMs */
int bl_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                     const float* rstd, int64_t rows, int32_t dim,
                     float* dx, float* d_gamma, float* d_beta, float* partial, bl_stream_t stream);
/* y = dropout(tanh(x)) elementwise; keep-mask is regenerated from (seed, element index):
 * p_drop == 0 disables dropout.  t = tanh(x) is stored for backward. */
int bl_tanh_dropout_fwd(const float* x, int64_t n, float p_drop, uint64_t seed,
                        float* y, float* t_out, bl_stream_t stream);
int bl_tanh_dropout_bwd(const float* dy, const float* t, int64_t n, float p_drop, uint64_t seed,
                        float* dx, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Segment primitives of the heads — replace torch_scatter.scatter_{max,min,sum} and
 * buglab/models/utils.py:15-48 (scatter_log_softmax and the thin wrappers), used by
 * buglab/models/layers/localizationmodule.py:59,75,105 and buglab/models/gnn.py:299,305.
 * src is [L,F] row-major, index[L] in [0,S) in ANY order.
 * ------------------------------------------------------------------------------------------------ */
/* out[s,f] = max/min_l src[l,f] over index[l]==s (0 if none); arg[s,f] = first l attaining it (L if none).
 * is_min != 0 selects minimum. */
int bl_segment_minmax(const float* src, const int32_t* index, int64_t L, int32_t F, int64_t S,
                      int32_t is_min, float* out, int32_t* arg, bl_stream_t stream);
/* d_src[l,f] = (arg[index[l],f]==l) ? d_out[index[l],f] : 0 */
int bl_segment_minmax_bwd(const float* d_out, const int32_t* arg, const int32_t* index,
                          int64_t L, int32_t F, float* d_src, bl_stream_t stream);
/* out[s,f] = sum_l src[l,f] (fp32 REDs; out zero-filled here) */
int bl_segment_sum(const float* src, const int32_t* index, int64_t L, int32_t F, int64_t S,
                   float* out, bl_stream_t stream);
/* scatter_log_softmax over 1-D scores (utils.py:15-28, eps=1e-12 inside the log):
 * seg_max[S], seg_sum[S] are outputs kept for backward. */
int bl_segment_log_softmax_fwd(const float* src, const int32_t* index, int64_t L, int64_t S, float eps,
                               float* out, float* seg_max, float* seg_sum, bl_stream_t stream);
/* d_src[l] = d_out[l] - exp(out[l]) * sum_{index==index[l]} d_out ; seg_tmp[S] is scratch. */
int bl_segment_log_softmax_bwd(const float* d_out, const float* out, const int32_t* index,
                               int64_t L, int64_t S, float* d_src, float* seg_tmp, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Subtoken embedding + masked max-pool (ptgnn StrElementRepresentationModel with
 * token_splitting="subtoken", subtoken_combination="max"; reference wiring
 * buglab/models/modelregistry.py:57-66,79-82).  ids[N,T] (padded), lens[N] in [1,T].
 * Dropout (p_drop>0) is applied to the embedded subtokens before the max, mask from (seed, n, t, j).
 * ------------------------------------------------------------------------------------------------ */
int bl_subtoken_maxpool_fwd(const float* emb, const int32_t* ids, const int32_t* lens,
                            int64_t N, int32_t T, int32_t H, float p_drop, uint64_t seed,
                            float* out, int32_t* arg, bl_stream_t stream);
/* d_emb[V,H] must be zero-filled (or hold a running gradient) by the caller; accumulated with REDs. */
int bl_subtoken_maxpool_bwd(const float* d_out, const int32_t* ids, const int32_t* arg,
                            int64_t N, int32_t T, int32_t H, float p_drop, uint64_t seed,
                            float* d_emb, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step over ONE flat fp32 buffer — Adam(lr, torch defaults) with global-norm clipping,
 * reference buglab/models/utils.py:51-52 and train.py:104 (clip_gradient_norm=0.5).
 * The flat gradient buffer is also the NCCL all-reduce bucket.
 * ------------------------------------------------------------------------------------------------ */
/* sqnorm[0] = sum g^2  (sqnorm zeroed here; partial[>=1024] scratch) */
int bl_grad_sqnorm(const float* grad, int64_t n, float* sqnorm, float* partial, bl_stream_t stream);
/* clip coefficient c = min(1, max_norm / (sqrt(sqnorm[0]) * grad_scale + 1e-6)) (max_norm<=0: no clip);
 * g' = g*grad_scale*c ; Adam update with bias correction for step `step` (1-based). */
int bl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, int64_t step,
                 float max_norm, const float* sqnorm, float grad_scale, bl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Edge-biased multi-head attention of the sequence models (seq-great / seq-rat) — replaces
 * RelationalMultiheadAttention._compute_attention_scores + _add_edge_attention_scores + _compute_attention_probs +
 * _compute_weighted_sum + _add_edge_value_biases (reference buglab/models/layers/multihead_attention.py:57-79,
 * relational_multihead_attention.py:89-178) in the vector-bias mode the registry builds (seqmodel.py:93-107).
 *   q (pre-scaled), k, v, out: fp32 [B, H, L, D], D in {8, 16, 32, 64};  lse: [B, H, L];  lengths[B]: keys >= length masked.
 *   bias / vbias: [T2, H, D] with rows [0, T) = forward tables, [T, 2T) = reverse tables; vbias NULL for seq-great.
 *   "entries" = the two directions of every typed edge (b, s, t, type): (row s, key t, table type), (row t, key s, table
 *   T + type).  row_ptr[B*L+1] / row_key / row_tab list them grouped by (b, query row) with ascending keys; col_ptr /
 *   col_query / col_tab list the same entries grouped by (b, key) with ascending query rows.  Duplicates add up.
 *   p_drop / seed: dropout on the attention probabilities (multihead_attention.py:77), counter-based mask over
 *   (seed, b, h, query, key) as in the other kernels; pass the same pair to the backward call.
 * This is the fp32 CUDA-core path (one thread per row), kept for shapes the tensor-core path below does not cover (padded
 * length > 512) and as its referee; its arithmetic is pinned against the oracle through the host emulation of the same
 * source (csrc/seq_attention_core.h) and on B200 by tests/test_seq_attention_gpu.py.
 * ------------------------------------------------------------------------------------------------ */
int bl_seq_attention_supported(int32_t head_dim);
int bl_seq_attention_fwd(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                         const float* vbias, const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab,
                         int32_t B, int32_t H, int32_t L, int32_t D, int32_t T2, float p_drop, uint64_t seed,
                         float* out, float* lse, bl_stream_t stream);
/* dq, dk, dv: [B, H, L, D];  d_entry_bias (and d_entry_vbias when vbias != NULL): [entries, H, D] in row order — the caller
 * sums them into the tables by row_tab;  delta[B, H, L] is scratch. */
int bl_seq_attention_bwd(const float* q, const float* k, const float* v, const int32_t* lengths, const float* bias,
                         const float* vbias, const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab,
                         const int32_t* col_ptr, const int32_t* col_query, const int32_t* col_tab,
                         int32_t B, int32_t H, int32_t L, int32_t D, int32_t T2, float p_drop, uint64_t seed,
                         const float* out, const float* lse, const float* d_out, float* dq, float* dk, float* dv,
                         float* d_entry_bias, float* d_entry_vbias, float* delta, bl_stream_t stream);

/* Tensor-core path of the same attention (csrc/seq_attention_tc.cu; head size <= 64 zero-padded to 64, padded length
 * Lp in {128, 256, 512} >= L).  The GEMM-shaped products run on the TMA-fed tcgen05 kernels above, one segment per
 * (sample, head) whose "weight matrix" is that head's K, V, Q or dO:
 *     forward   S  = Q K^T    bl_tma_project (n_out = Lp, k_in = 64)      O  = P' V     bl_tma_project (n_out = 64, k_in = Lp)
 *     backward  dP = dO V^T   bl_tma_project                              dQ = dS K     bl_tma_project
 *               dK = dS^T Q,  dV = P'^T dO                                bl_tma_weight_grad (m_out = Lp, n_in = 64)
 * and the two row kernels below do everything in between, one warp per (sample, head, query) row of the score tile.
 *   bl_seq_softmax_fwd: scores [B*H*Lp, Lp] (in: Q K^T; out: + the typed-edge terms <q_i, bias[tab_e]>, kept for backward);
 *     q [B*H*Lp, 64]; lse [B*H*Lp]; p_split: fp16 hi/lo split table [2][B*H*Lp + 1][Lp] of P' = dropout(softmax) scaled by
 *     the power of two derived from *amax_p (pass 1 / (1 - p_drop)); o_extra [B*H*Lp, 64] = the value-bias terms of "rat"
 *     (NULL iff vbias is NULL).  row_ptr / row_key / row_tab as above (indexed with the unpadded L).
 *   bl_seq_softmax_bwd: d_scores [B*H*Lp, Lp] (in: dP = dO V^T; out: dS); out / d_out [B*H*Lp, 64]; p_split recomputed;
 *     dq_extra [B*H*Lp, 64] = entry terms of dQ; d_entry_bias / d_entry_vbias [entries, H, d_entry_dim] as in
 *     bl_seq_attention_bwd (d_entry_dim = the caller's unpadded head size). */
int bl_seq_attention_tc_supported(int32_t head_dim, int32_t max_len);
int bl_seq_softmax_fwd(float* scores, const float* q, const int32_t* lengths, const float* bias, const float* vbias,
                       const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab, int32_t B, int32_t H, int32_t L,
                       int32_t Lp, int32_t T2, float p_drop, uint64_t seed, const float* amax_p, float* lse, void* p_split,
                       float* o_extra, bl_stream_t stream);
int bl_seq_softmax_bwd(const float* scores, const float* lse, const float* q, const int32_t* lengths, const float* bias,
                       const float* vbias, const int32_t* row_ptr, const int32_t* row_key, const int32_t* row_tab, int32_t B,
                       int32_t H, int32_t L, int32_t Lp, int32_t T2, float p_drop, uint64_t seed, const float* amax_p,
                       const float* out, const float* d_out, float* d_scores, void* p_split, float* dq_extra,
                       float* d_entry_bias, float* d_entry_vbias, int32_t d_entry_dim, bl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BUGLAB_B200_H */
