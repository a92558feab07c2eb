// Typed-edge plan: CSR by target + unique (type,node) pair tables, built on device once per
// minibatch (see include/buglab_b200.h).  Integer bookkeeping only — must be bit-exact against the
// host restatement in oracle/plan_ref.py.  Replaces the per-layer `torch.cat` of adjacency targets
// and per-type `index_select`s of ptgnn's MlpMessagePassingLayer (reference call site
// buglab/models/gnnlayerdefs.py:6-23) and consumes the edge-typed adjacency of
// buglab/representations/data.py:139-167.
//
// All sorts are stable LSD radix sorts (cub::DeviceRadixSort), so the result is deterministic:
//   1. sort edges by tgt; the input is type-major, stability => order (tgt, type, original index)
//   2. S-pairs: sort sorted-edge ids by the pair key of (type, src) — type-major, or node-blocked (block, type, node) —
//      flag key changes, scan -> urow, s_node, segment pointers
//   3. T-pairs: same with (type, tgt) -> vrow, t_node
//   4. node -> pair CSRs (for the segmented-sum backward of the row gathers)
// Pair counts are only known on device; tables are sized by the upper bound E and padded entries
// carry sentinel keys, so no host synchronisation happens here.
#include <cub/cub.cuh>

#include "common.cuh"

namespace bl {

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static inline int bits_for(uint64_t max_value) {  // number of bits needed to represent max_value
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

__global__ void iota_kernel(int* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)i;
}

__global__ void gather_sorted_edges(const int* __restrict__ perm, const int* __restrict__ src,
                                    const int* __restrict__ etype, int64_t E, int* __restrict__ e_src,
                                    int* __restrict__ e_type) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const int p = perm[i];
    e_src[i] = src[p];
    e_type[i] = etype[p];
}

// ptr[t] = first position i with keys[i] >= t, for t in [0, num_segments]; keys sorted ascending,
// entries with key >= num_segments are padding.
__global__ void fill_ptr_kernel(const unsigned* __restrict__ keys, int64_t n_items, int64_t num_segments,
                                int* __restrict__ ptr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_items) return;
    int64_t lo = (i == 0) ? 0 : (int64_t)keys[i - 1] + 1;
    int64_t hi = (i == n_items) ? num_segments : (int64_t)keys[i];
    if (hi > num_segments) hi = num_segments;
    for (int64_t t = lo; t <= hi; ++t) ptr[t] = (int)i;
}

// Pair key.  Type-major layout (block == 0): key = type*N + node, pairs ordered by (type, node), one segment per type.
// Node-blocked layout (block = B > 0): key = ((node / B)*K + type)*B + node % B, pairs ordered by (node block, type, node):
// one segment per (block, type), so that the GEMMs and the by-source edge backward sweep the node states ONCE (all
// types of a block while its rows are L2-resident) instead of once per type.
__device__ __forceinline__ unsigned long long pair_key(int type, int node, int64_t N, int K, int block) {
    if (block <= 0) return (unsigned long long)type * (unsigned long long)N + (unsigned long long)node;
    const unsigned long long seg = (unsigned long long)(node / block) * (unsigned long long)K + (unsigned long long)type;
    return seg * (unsigned long long)block + (unsigned long long)(node % block);
}
__device__ __forceinline__ int node_of_key(unsigned long long key, int64_t N, int K, int block) {
    if (block <= 0) return (int)(key % (unsigned long long)N);
    const unsigned long long seg = key / (unsigned long long)block;
    return (int)((seg / (unsigned long long)K) * (unsigned long long)block + key % (unsigned long long)block);
}
__global__ void make_pair_keys(const int* __restrict__ e_type, const int* __restrict__ node, int64_t E,
                               int64_t N, int K, int block, unsigned long long* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) keys[i] = pair_key(e_type[i], node[i], N, K, block);
}

__global__ void flag_changes(const unsigned long long* __restrict__ keys_sorted, int64_t E,
                             int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) flags[i] = (i == 0 || keys_sorted[i] != keys_sorted[i - 1]) ? 1 : 0;
}

// pid_incl = inclusive scan of flags.  Writes row id of each sorted edge, node of each pair, the
// unique keys, and (last thread) the pair count.
__global__ void scatter_pairs(const unsigned long long* __restrict__ keys_sorted,
                              const int* __restrict__ edge_of, const int* __restrict__ flags,
                              const int* __restrict__ pid_incl, int64_t E, int64_t N, int K, int block,
                              int* __restrict__ row_of_edge, int* __restrict__ pair_node,
                              unsigned* __restrict__ pair_node_key, unsigned long long* __restrict__ ukeys,
                              int* __restrict__ count_out, int* __restrict__ edge_ptr, int* __restrict__ edge_idx,
                              const int* __restrict__ e_tgt_sorted, int* __restrict__ edge_tgt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const int pid = pid_incl[i] - 1;
    row_of_edge[edge_of[i]] = pid;
    if (edge_idx != nullptr) edge_idx[i] = edge_of[i];  // sorted-edge ids grouped by pair, ascending inside a pair (stable sort)
    if (edge_tgt != nullptr) edge_tgt[i] = e_tgt_sorted[edge_of[i]];  // ... and their target nodes (saves a dependent load)
    if (flags[i]) {
        const int node = node_of_key(keys_sorted[i], N, K, block);
        pair_node[pid] = node;
        pair_node_key[pid] = (unsigned)node;
        ukeys[pid] = keys_sorted[i];
        if (edge_ptr != nullptr) edge_ptr[pid] = (int)i;
    }
    if (i == E - 1) {
        *count_out = pid + 1;
        if (edge_ptr != nullptr) edge_ptr[pid + 1] = (int)E;
    }
}

// pad pair tables beyond the pair count with sentinels (node key N sorts last)
__global__ void pad_pairs(const int* __restrict__ count, int64_t E, int64_t N, int* __restrict__ pair_node,
                          unsigned* __restrict__ pair_node_key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    if (i >= *count) {
        pair_node[i] = 0;
        pair_node_key[i] = (unsigned)N;
    }
}

// seg_ptr[s] = lower_bound(ukeys[0:P], s*unit), s in [0, S]   (type-major: unit = N, S = K; blocked: unit = B, S = blocks*K)
__global__ void type_ptr_kernel(const unsigned long long* __restrict__ ukeys, const int* __restrict__ count,
                                int64_t unit, int S, int* __restrict__ type_ptr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > S) return;
    const unsigned long long target = (unsigned long long)k * (unsigned long long)unit;
    int lo = 0, hi = *count;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
    }
    type_ptr[k] = lo;
}

struct Workspace {
    // carved from the caller's buffer; sizes depend only on (E, N)
    unsigned* k32_a;            // E
    unsigned* k32_b;            // E
    int* v32_a;                 // E
    int* v32_b;                 // E
    unsigned long long* k64_a;  // E
    unsigned long long* k64_b;  // E
    unsigned long long* ukeys;  // E
    int* flags;                 // E
    int* scan;                  // E
    void* cub_temp;
    size_t cub_bytes;
};

static size_t cub_temp_bytes(int64_t E) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)E, 0, 32);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned long long*)nullptr,
                                    (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr,
                                    (int)E, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, c, (const int*)nullptr, (int*)nullptr, (int)E);
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}

static size_t carve(Workspace* ws, void* base, int64_t E) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> void* {
        void* p = base ? (void*)((char*)base + off) : nullptr;
        off += align_up(bytes);
        return p;
    };
    const size_t e = (size_t)(E > 0 ? E : 1);
    ws->k32_a = (unsigned*)take(e * 4);
    ws->k32_b = (unsigned*)take(e * 4);
    ws->v32_a = (int*)take(e * 4);
    ws->v32_b = (int*)take(e * 4);
    ws->k64_a = (unsigned long long*)take(e * 8);
    ws->k64_b = (unsigned long long*)take(e * 8);
    ws->ukeys = (unsigned long long*)take(e * 8);
    ws->flags = (int*)take(e * 4);
    ws->scan = (int*)take(e * 4);
    ws->cub_bytes = cub_temp_bytes((int64_t)e);
    ws->cub_temp = take(ws->cub_bytes);
    return off;
}

// Builds one (type,node) pair table + node->pair CSR.  `node_of_edge[i]` is the node (src or tgt)
// of sorted edge i.
static int build_pairs(const Workspace& ws, const int* e_type, const int* node_of_edge, int64_t E,
                       int64_t N, int K, int block, int* row_of_edge, int* pair_node, int* type_ptr,
                       int* by_node_ptr, int* by_node_idx, int* count_out, int* edge_ptr, int* edge_idx,
                       const int* e_tgt_sorted, int* edge_tgt, cudaStream_t stream) {
    const int T = 256;
    const unsigned g = grid_for(E, T);
    size_t tb = ws.cub_bytes;
    make_pair_keys<<<g, T, 0, stream>>>(e_type, node_of_edge, E, N, K, block, ws.k64_a);
    iota_kernel<<<g, T, 0, stream>>>(ws.v32_a, E);
    const int64_t unit = block > 0 ? block : N;
    const int num_segs = block > 0 ? (int)((N + block - 1) / block) * K : K;
    const int key_bits = bits_for((uint64_t)num_segs * (uint64_t)unit);
    cub::DeviceRadixSort::SortPairs(ws.cub_temp, tb, ws.k64_a, ws.k64_b, ws.v32_a, ws.v32_b, (int)E, 0,
                                    key_bits, stream);
    flag_changes<<<g, T, 0, stream>>>(ws.k64_b, E, ws.flags);
    tb = ws.cub_bytes;
    cub::DeviceScan::InclusiveSum(ws.cub_temp, tb, ws.flags, ws.scan, (int)E, stream);
    scatter_pairs<<<g, T, 0, stream>>>(ws.k64_b, ws.v32_b, ws.flags, ws.scan, E, N, K, block, row_of_edge, pair_node,
                                       ws.k32_a, ws.ukeys, count_out, edge_ptr, edge_idx, e_tgt_sorted, edge_tgt);
    pad_pairs<<<g, T, 0, stream>>>(count_out, E, N, pair_node, ws.k32_a);
    type_ptr_kernel<<<grid_for(num_segs + 1, 64), 64, 0, stream>>>(ws.ukeys, count_out, unit, num_segs, type_ptr);
    // node -> pairs CSR: stable sort of pair ids by node (padding has key N and sorts last)
    iota_kernel<<<g, T, 0, stream>>>(ws.v32_a, E);
    tb = ws.cub_bytes;
    cub::DeviceRadixSort::SortPairs(ws.cub_temp, tb, ws.k32_a, ws.k32_b, ws.v32_a, by_node_idx, (int)E, 0,
                                    bits_for((uint64_t)N), stream);
    fill_ptr_kernel<<<grid_for(E + 1, T), T, 0, stream>>>(ws.k32_b, E, N, by_node_ptr);
    return check_launch("bl_plan_build/build_pairs");
}

}  // namespace bl

using namespace bl;

extern "C" size_t bl_plan_workspace_bytes(int64_t num_edges, int64_t num_nodes, int32_t num_edge_types) {
    (void)num_nodes;
    (void)num_edge_types;
    Workspace ws;
    return carve(&ws, nullptr, num_edges);
}

extern "C" int bl_plan_build(const int32_t* src, const int32_t* tgt, const int32_t* etype, int64_t E,
                             int64_t N, int32_t K, int32_t* e_perm, int32_t* e_src, int32_t* e_type,
                             int32_t* row_ptr, int32_t* urow, int32_t* vrow, int32_t* s_node,
                             int32_t* s_type_ptr, int32_t* s_by_node_ptr, int32_t* s_by_node_idx,
                             int32_t* t_node, int32_t* t_type_ptr, int32_t* t_by_node_ptr,
                             int32_t* t_by_node_idx, int32_t* counts, int32_t* s_edge_ptr, int32_t* s_edge_idx,
                             int32_t* e_tgt, int32_t* s_edge_tgt, int32_t block_nodes, void* workspace, size_t workspace_bytes,
                             bl_stream_t stream_) {
    if (E < 0 || N <= 0 || K <= 0 || E > 0x7ffffff0LL || N > 0x7ffffff0LL || block_nodes < 0) return BL_ERR_INVALID_ARGUMENT;
    const int num_segs = block_nodes > 0 ? (int)((N + block_nodes - 1) / block_nodes) * K : K;
    cudaStream_t stream = (cudaStream_t)stream_;
    Workspace ws;
    const size_t need = carve(&ws, workspace, E);
    if (workspace_bytes < need) return BL_ERR_WORKSPACE_TOO_SMALL;
    const int T = 256;
    if (E == 0) {
        // no edges: every table is empty
        int rc = check_cuda(cudaMemsetAsync(row_ptr, 0, (size_t)(N + 1) * 4, stream), "plan memset");
        if (rc) return rc;
        cudaMemsetAsync(s_by_node_ptr, 0, (size_t)(N + 1) * 4, stream);
        cudaMemsetAsync(t_by_node_ptr, 0, (size_t)(N + 1) * 4, stream);
        cudaMemsetAsync(s_type_ptr, 0, (size_t)(num_segs + 1) * 4, stream);
        cudaMemsetAsync(t_type_ptr, 0, (size_t)(num_segs + 1) * 4, stream);
        if (s_edge_ptr != nullptr) cudaMemsetAsync(s_edge_ptr, 0, 4, stream);
        return check_cuda(cudaMemsetAsync(counts, 0, 8, stream), "plan memset");
    }
    const unsigned g = grid_for(E, T);
    // 1. stable sort by target
    iota_kernel<<<g, T, 0, stream>>>(ws.v32_a, E);
    size_t tb = ws.cub_bytes;
    cub::DeviceRadixSort::SortPairs(ws.cub_temp, tb, (const unsigned*)tgt, ws.k32_b, ws.v32_a, e_perm, (int)E,
                                    0, bits_for((uint64_t)N), stream);
    gather_sorted_edges<<<g, T, 0, stream>>>(e_perm, src, etype, E, e_src, e_type);
    fill_ptr_kernel<<<grid_for(E + 1, T), T, 0, stream>>>(ws.k32_b, E, N, row_ptr);
    int rc = check_launch("bl_plan_build/sort");
    if (rc) return rc;
    // sorted targets are needed as the node array of the T-pairs; keep them in scan-free storage:
    // ws.k32_b is reused by build_pairs' CSR step only AFTER its keys were consumed, so copy first.
    int* e_tgt_sorted = t_by_node_idx;  // scratch until the T-pair CSR overwrites it at the very end
    rc = check_cuda(cudaMemcpyAsync(e_tgt_sorted, ws.k32_b, (size_t)E * 4, cudaMemcpyDeviceToDevice, stream),
                    "plan copy");
    if (rc) return rc;
    if (e_tgt != nullptr) {  // target node of every sorted edge (for the by-source backward of the edge kernel)
        rc = check_cuda(cudaMemcpyAsync(e_tgt, ws.k32_b, (size_t)E * 4, cudaMemcpyDeviceToDevice, stream), "plan copy");
        if (rc) return rc;
    }
    // 2. S-pairs (type, src)
    rc = build_pairs(ws, e_type, e_src, E, N, K, block_nodes, urow, s_node, s_type_ptr, s_by_node_ptr, s_by_node_idx,
                     counts + 0, s_edge_ptr, s_edge_idx, e_tgt_sorted, (s_edge_ptr != nullptr) ? s_edge_tgt : nullptr, stream);
    if (rc) return rc;
    // 3. T-pairs (type, tgt).  e_tgt_sorted aliases t_by_node_idx, which build_pairs writes only in
    // its final sort, after every read of node_of_edge (make_pair_keys) has been issued in stream order.
    rc = build_pairs(ws, e_type, e_tgt_sorted, E, N, K, block_nodes, vrow, t_node, t_type_ptr, t_by_node_ptr,
                     t_by_node_idx, counts + 1, nullptr, nullptr, nullptr, nullptr, stream);
    return rc;
}
