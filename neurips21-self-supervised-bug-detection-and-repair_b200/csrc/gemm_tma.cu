// Split-fp16 grouped GEMMs of the typed-edge message layer on TMA + tcgen05 (sm_100a only), second generation.
//
// What changed against pair_project_tc.cu (round 1): operands are split into fp16 hi/lo parts ONCE, at node (or table
// row) granularity, by bl_rows_split_f16 — not per pair row inside the GEMM loader — and every operand tile is moved
// by the TMA engine straight into the 128-byte-swizzled shared-memory layout the UMMA descriptors read:
//   * gathered A rows (h[s_node[p]], h[t_node[p]]):  cp.async.bulk.tensor.2d ... tile::gather4  (4 rows per request)
//   * contiguous A rows (gradient tables) and the weight parts:  ordinary tiled boxes
// so the SM issues ~70 asynchronous copies per 128x64 chunk instead of ~30k convert/store instructions.  The CTA is
// warp specialised: warps 0-3 TMA producers (a gather4 takes per-lane row coordinates but is a uniform-datapath
// instruction, so a warp issues its 32 requests one lane at a time; four warps share the work: warp w serves the
// chunks of parity w >> 1 and the hi (w & 1 == 0) or lo part), warp 4 MMA issuer and TMEM owner, warps 5-8 epilogue,
// with a STAGES-deep smem ring and two TMEM accumulators.  With CG == 2 two CTAs of a cluster (one TPC) run
// tcgen05.mma.cta_group::2 on a 256-row tile: each CTA stages only its 128 rows of A and HALF of the weight columns,
// which halves the L2->SM weight traffic that bounded the first-generation kernel.
//
// Arithmetic (unchanged): x.w ~= x_hi.w_hi + x_hi.w_lo + x_lo.w_hi with fp32 accumulation in TMEM.
//
// Reference semantics replaced: the per-type Linear_k of ptgnn's MlpMessagePassingLayer hoisted to unique
// (type, node) pairs (buglab/models/gnnlayerdefs.py:6-23), its backward w.r.t. inputs and weights, and the
// node-update Linear(M -> D_out) of the same layer.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace bl {
namespace tg {

constexpr int TILE_M = 128;   // rows per CTA = TMEM lanes
constexpr int CHUNK_K = 64;   // fp16 elements per smem row = 128 bytes = one swizzle atom
constexpr int THREADS = 288;  // warps 0-3 producers, 4 MMA + TMEM alloc, 5-8 epilogue
constexpr int MMA_WARP = 4, FIRST_EPI_WARP = 5;
// slab = unit of the weight gradient's pair-row reduction (also bounds the truncating tensor-core accumulation chain:
// 768 MMAs) and of the weight-stationary projection (one weight load per slab)
constexpr int SLAB_ROWS = 4096;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (pair leader)

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __noinline__ void mbar_timeout(uint32_t addr, uint32_t parity, int role) {
    printf("buglab_b200 gemm_tma: mbarrier wait timed out (block %d thread %d role %d smem 0x%x parity %u)\n", blockIdx.x,
           threadIdx.x, role, addr, parity);
    __trap();
}
// Bounded wait (test_wait never suspends): a protocol bug traps within ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int role) {
    const uint32_t addr = smem_u32(bar);
    uint64_t t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if ((spin & 1023u) == 1023u) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) mbar_timeout(addr, parity, role);
        }
    }
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    if (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    } else {  // the barrier at this offset in BOTH CTAs of the pair
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// TMA: tiled 2-D box -> smem, completion on an mbarrier (of the pair leader when CG == 2)
template <int CG>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    if (CG == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
    }
}
// TMA gather4: four rows r0..r3 of a 2-D tensor (box = 1 row x 64 columns) -> four consecutive 128-byte smem rows
template <int CG>
__device__ __forceinline__ void tma_gather4(const CUtensorMap* map, uint64_t* bar, void* dst, int col, int r0, int r1, int r2, int r3) {
    if (CG == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::2 [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    }
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, SWIZZLE_128B smem matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 at [0,14), SBO = 1024 B (8 rows
// of 128 B) at [32,46), version 1 at [46,48), layout SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, SWIZZLE_128B: K rows of 128 bytes (64 fp16 along M/N), 8-row groups 1024 B apart (SBO), further 64-wide
// M/N blocks `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 at [4,6)), A = B = F16 (0), a/b major at 15/16
// (0 = K-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_f16_f32(int m, int n, bool mn_major) {
    return (1u << 4) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// segment containing `tile` in the prefix array tile_ptr[0..num_segs]
__device__ __forceinline__ int find_segment(const int* __restrict__ tile_ptr, int num_segs, int tile) {
    int lo = 0, hi = num_segs;  // invariant: tile_ptr[lo] <= tile < tile_ptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(tile_ptr + mid) <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------ coalesced epilogue
// After tcgen05.ld a thread holds 32 consecutive columns of ITS row: storing them directly makes every warp-level store
// touch 32 different 128-byte lines (measured: L1TEX 66-74 % busy, the tensor pipe waiting on the epilogue).  Each
// epilogue warp therefore transposes a 32 x 32 fp32 block through a private 4 KB shared-memory tile (16-byte chunks
// XOR-swizzled by row & 7: conflict-free both ways) and writes 4 full 128-byte row segments per instruction.
constexpr int EPI_STAGE_FLOATS = 32 * 32;

__device__ __forceinline__ void epi_stage_write(float* stage, const float* v, int lane, float scale) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stage + lane * 32 + ((q ^ (lane & 7)) << 2)) =
            make_float4(v[4 * q] * scale, v[4 * q + 1] * scale, v[4 * q + 2] * scale, v[4 * q + 3] * scale);
    __syncwarp();
}
// rows [0, rows_valid) of the staged block -> out[r * ld + 0..31]  (+ bias[0..31])
__device__ __forceinline__ void epi_store_rows(const float* stage, int lane, float* out, int64_t ld, int rows_valid,
                                               const float* bias) {
    const int sub = lane >> 3, ch = lane & 7;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(bias) + ch);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((ch ^ (r & 7)) << 2));
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        if (r < rows_valid) *reinterpret_cast<float4*>(out + (int64_t)r * ld + (ch << 2)) = o;
    }
    __syncwarp();
}
// same, added to the destination with one 16-byte vector RED per lane (weight-gradient partial tiles)
__device__ __forceinline__ void epi_red_rows(const float* stage, int lane, float* out, int64_t ld) {
    const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        const float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((ch ^ (r & 7)) << 2));
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (int64_t)r * ld + (ch << 2)), "f"(o.x), "f"(o.y),
                     "f"(o.z), "f"(o.w) : "memory");
    }
    __syncwarp();
}

// =====================================================================================================================
// Projection:  out[p, 0:N] = inv_scale * A[row(p), 0:Kin] . W_type(seg(p))[0:N, 0:Kin]^T (+ bias_type)
//   A is a split table  [2 parts][a_rows][Kin] fp16  (part 0 = hi, part 1 = lo; its LAST row of each part is zero and
//   serves as the padding row);  row(p) = idx[p] (GATHER) or p.
//   Pair rows are grouped in segments that share one weight matrix; the kernels walk a device-side table of work units
//   (row range + matrix id per tile / slab, built once per plan by bl_segment_units: one 16-byte load locates a tile —
//   a binary search over ~1 200 (node block, type) segments per tile cost the producers ~1.4 us of dependent L2 latency
//   per 3 us tile).
// =====================================================================================================================
struct ProjParams {
    const int* idx;       // [P] rows of the split table, or nullptr (contiguous: row(p) = p)
    const float* bias;    // [num_types, N] or nullptr
    const float* amax;    // nullable: out is multiplied by 1 / pow2_scale_for(*amax)  (undoes the pre-scale of A)
    const float* amax_b;  // nullable: the same for the pre-scale of the weight parts
    const int4* units;    // work units {first pair row, end pair row, weight matrix, -}: tiles of <= 128*CG rows of one
                          // segment (proj_kernel) or slabs of <= SLAB_ROWS rows (proj_bs_kernel); bl_segment_units
    const int* num_units; // device scalar: number of valid entries of `units`
    float* out;           // [P, N]
    int N, Kin, a_rows;
};

template <int NT, int CG>
struct ProjCfg {
    static constexpr int NTL = NT / CG;                       // weight rows (= output columns) staged per CTA
    static constexpr uint32_t A_BYTES = TILE_M * 128;         // one part of a 128 x 64 fp16 tile
    static constexpr uint32_t B_BYTES = NTL * 128;
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (STAGE_BYTES <= 48 * 1024) ? 4 : (STAGE_BYTES <= 64 * 1024 ? 3 : 2);
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
};

template <int NT, int CG, bool GATHER>
__global__ void __launch_bounds__(THREADS, 1)
proj_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const ProjParams p) {
    using Cfg = ProjCfg<NT, CG>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float epi_stage[4][EPI_STAGE_FLOATS];  // one 32 x 32 transposition tile per epilogue warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int num_clusters = gridDim.x / CG, cluster_id = blockIdx.x / CG;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], CG);   // one producer arrival per CTA of the pair (+ transaction bytes)
            mbar_init(&empty[i], 1);   // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);        // tcgen05.commit
            mbar_init(&acc_empty[i], CG * 4);  // one arrival per epilogue warp of every CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
    }
    if (CG == 2) cluster_sync_all();  // both CTAs alive before the pair-wide TMEM allocation
    if (warp == MMA_WARP) tmem_alloc<CG>(&tmem_base_smem, 2 * NT);
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int n_splits = p.N / NT;
    const int total_work = __ldg(p.num_units) * n_splits;
    const int num_chunks = p.Kin / CHUNK_K;

    if (warp < 4) {
        // ============================== TMA PRODUCERS ==============================
        // warp w: chunks whose running index has parity (w >> 1); part (w & 1): 0 = hi (also owns the barrier arrival), 1 = lo
        const int part = warp & 1, parity = warp >> 1;
        int stage = 0;
        uint32_t phase = 0, chunk_counter = 0;
        for (int work = cluster_id; work < total_work; work += num_clusters) {
            const int tile = work / n_splits, split = work - tile * n_splits;
            const int4 unit = __ldg(p.units + tile);
            const int row0 = unit.x + (int)cta_rank * TILE_M, row_end = unit.y, type = unit.z;
            int rows[4];
            if (GATHER) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = row0 + lane * 4 + j;
                    rows[j] = ((r < row_end) ? __ldg(p.idx + r) : (p.a_rows - 1)) + part * p.a_rows;  // padding -> the zero row
                }
            }
            const int b_row = (type * 2 + part) * p.N + split * NT + (int)cta_rank * Cfg::NTL;
            for (int c = 0; c < num_chunks; ++c, ++chunk_counter) {
                if ((int)(chunk_counter & 1u) == parity) {
                    mbar_wait(&empty[stage], phase ^ 1u, 0);
                    uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
                    if (lane == 0) {
                        if (part == 0) {
                            if (leader) mbar_arrive_expect_tx(&full[stage], STAGE_BYTES * CG);
                            else mbar_arrive_cluster(&full[stage], 0);
                        }
                        tma_load_2d<CG>(&map_b, &full[stage], st + 2 * A_BYTES + part * B_BYTES, c * CHUNK_K, b_row);
                        if (!GATHER) {
                            // rows past the segment end are loaded too (next rows of the table, zero beyond its end): every
                            // output row depends on its own A row only and rows >= row_end are never stored
                            tma_load_2d<CG>(&map_a, &full[stage], st + part * A_BYTES, c * CHUNK_K, part * p.a_rows + row0);
                        }
                    }
                    __syncwarp();
                    if (GATHER)
                        tma_gather4<CG>(&map_a, &full[stage], st + part * A_BYTES + lane * 512, c * CHUNK_K, rows[0], rows[1],
                                        rows[2], rows[3]);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == MMA_WARP) {
        // ============================== MMA ISSUER (pair leader only) ==============================
        if (leader) {
            const uint32_t idesc = idesc_f16_f32(TILE_M * CG, NT, false);
            int stage = 0;
            uint32_t phase = 0, tile_counter = 0;
            for (int work = cluster_id; work < total_work; work += num_clusters, ++tile_counter) {
                const uint32_t a = tile_counter & 1u, ause = tile_counter >> 1;
                mbar_wait(&acc_empty[a], (ause & 1u) ^ 1u, 1);  // the epilogues drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + a * NT;
                for (int c = 0; c < num_chunks; ++c) {
                    mbar_wait(&full[stage], phase, 2);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_hi = smem_u32(smem + (size_t)stage * STAGE_BYTES), a_lo = a_hi + A_BYTES;
                        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                        for (int kk = 0; kk < CHUNK_K / 16; ++kk) {
                            const uint32_t koff = kk * 32;  // 16 halfs = 32 bytes along K inside the swizzle atom
                            umma_f16<CG>(tmem_acc, desc_k_sw128(a_hi + koff), desc_k_sw128(b_hi + koff), idesc, (c | kk) ? 1u : 0u);
                            umma_f16<CG>(tmem_acc, desc_k_sw128(a_hi + koff), desc_k_sw128(b_lo + koff), idesc, 1u);
                            umma_f16<CG>(tmem_acc, desc_k_sw128(a_lo + koff), desc_k_sw128(b_hi + koff), idesc, 1u);
                        }
                        tc_commit<CG>(&empty[stage]);                        // stage reusable once these MMAs have read it
                        if (c == num_chunks - 1) tc_commit<CG>(&acc_full[a]);  // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ============================== EPILOGUE (warps 5-8: TMEM lane quarter = warp % 4) ==============================
        const float inv_scale = ((p.amax != nullptr) ? 1.0f / pow2_scale_for(__ldg(p.amax)) : 1.0f) *
                                ((p.amax_b != nullptr) ? 1.0f / pow2_scale_for(__ldg(p.amax_b)) : 1.0f);
        const int lane_base = (warp & 3) * 32;
        float* stage_tile = epi_stage[warp & 3];
        uint32_t tile_counter = 0;
        for (int work = cluster_id; work < total_work; work += num_clusters, ++tile_counter) {
            const int tile = work / n_splits, split = work - tile * n_splits;
            const int4 unit = __ldg(p.units + tile);
            const int row0 = unit.x + (int)cta_rank * TILE_M, row_end = unit.y, type = unit.z;
            const int col0 = split * NT;
            const uint32_t a = tile_counter & 1u, ause = tile_counter >> 1;
            mbar_wait(&acc_full[a], ause & 1u, 3);
            tc_fence_after();
            const int rows_valid = max(0, min(32, row_end - (row0 + lane_base)));
            float* oblock = p.out + (size_t)(row0 + lane_base) * p.N + col0;
            const float* brow = p.bias ? p.bias + (size_t)type * p.N + col0 : nullptr;
#pragma unroll 1
            for (int j = 0; j < NT / 32; ++j) {
                float v[32];
                const int col = j * 32;
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(a * NT + col), v);
                epi_stage_write(stage_tile, v, lane, inv_scale);
                epi_store_rows(stage_tile, lane, oblock + col, p.N, rows_valid, brow ? brow + col : nullptr);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive_local(&acc_empty[a]);
                else mbar_arrive_cluster(&acc_empty[a], 0);
            }
        }
    }
    // every role has left its loop: all MMAs were waited for by the epilogues, all copies by the MMA warp
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<CG>(tmem_base, 2 * NT);
}

// =====================================================================================================================
// Projection, weight-stationary variant (CTA pairs, N = 256, Kin = 256: the H -> H layers' forward and backward-input).
// Work unit = one slab (<= SLAB_ROWS consecutive pair rows of one segment): the pair loads the slab's weight matrix ONCE
// (hi and lo parts, each CTA its 128 output columns: 128 KB of shared memory) and streams only the A tiles (three
// 32 KB stages) for the slab's <= 16 row tiles.  Streaming the weights per tile costs as many L2->SM bytes as A itself
// (at the tensor peak the two together meet the chip's L2 throughput); here they are read once per 4 096 rows.
// =====================================================================================================================
struct ProjBsCfg {
    static constexpr int NT = 256, CG = 2, NTL = 128, KIN = 256, CHUNKS = KIN / CHUNK_K;
    static constexpr uint32_t A_BYTES = TILE_M * 128;
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES;               // A hi + lo of one 64-wide chunk
    static constexpr uint32_t B_CHUNK_BYTES = NTL * 128;               // one part of one chunk of the weights
    static constexpr uint32_t B_BYTES = CHUNKS * 2 * B_CHUNK_BYTES;    // 128 KB
    static constexpr int STAGES = 2;  // 128 KB of weights + 2 x 32 KB of A + 16 KB of epilogue staging (static)
    static constexpr uint32_t SMEM_BYTES = B_BYTES + STAGES * STAGE_BYTES + 1024;
};

template <bool GATHER>
__global__ void __launch_bounds__(THREADS, 1)
proj_bs_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const ProjParams p) {
    using Cfg = ProjBsCfg;
    constexpr int CG = Cfg::CG, NT = Cfg::NT, STAGES = Cfg::STAGES, CHUNKS = Cfg::CHUNKS;
    constexpr uint32_t A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES, B_CHUNK_BYTES = Cfg::B_CHUNK_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem_b = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // resident weights
    uint8_t* smem_a = smem_b + Cfg::B_BYTES;                                        // A stage ring

    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2], b_full, b_empty;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float epi_stage[4][EPI_STAGE_FLOATS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int num_clusters = gridDim.x / CG, cluster_id = blockIdx.x / CG;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], CG);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], CG * 4);
        }
        mbar_init(&b_full, CG);
        mbar_init(&b_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
    }
    cluster_sync_all();
    if (warp == MMA_WARP) tmem_alloc<CG>(&tmem_base_smem, 2 * NT);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // p.units holds SLABS here (<= SLAB_ROWS rows); a slab has <= SLAB_ROWS / 256 row tiles
    const int total_slabs = __ldg(p.num_units);
    auto locate = [&](int slab, int& type, int& row_begin, int& row_end) {
        const int4 unit = __ldg(p.units + slab);
        row_begin = unit.x; row_end = unit.y; type = unit.z;
    };

    if (warp < 4) {
        // ============================== TMA PRODUCERS ==============================
        const int part = warp & 1, parity = warp >> 1;
        int stage = 0;
        uint32_t phase = 0, chunk_counter = 0, slab_counter = 0;
        for (int slab = cluster_id; slab < total_slabs; slab += num_clusters, ++slab_counter) {
            int type, row_begin, row_end;
            locate(slab, type, row_begin, row_end);
            if (parity == 0) {  // warps 0 (hi) and 1 (lo) bring the slab's weights in once the previous slab's MMAs are done
                mbar_wait(&b_empty, (slab_counter & 1u) ^ 1u, 4);
                if (lane == 0) {
                    if (part == 0) {
                        if (leader) mbar_arrive_expect_tx(&b_full, Cfg::B_BYTES * CG);
                        else mbar_arrive_cluster(&b_full, 0);
                    }
                    const int b_row = (type * 2 + part) * p.N + (int)cta_rank * Cfg::NTL;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c)
                        tma_load_2d<CG>(&map_b, &b_full, smem_b + (c * 2 + part) * B_CHUNK_BYTES, c * CHUNK_K, b_row);
                }
                __syncwarp();
            }
            const int num_tiles = (row_end - row_begin + TILE_M * CG - 1) / (TILE_M * CG);
            for (int t = 0; t < num_tiles; ++t) {
                const int row0 = row_begin + t * (TILE_M * CG) + (int)cta_rank * TILE_M;
                int rows[4];
                if (GATHER) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = row0 + lane * 4 + j;
                        rows[j] = ((r < row_end) ? __ldg(p.idx + r) : (p.a_rows - 1)) + part * p.a_rows;
                    }
                }
                for (int c = 0; c < CHUNKS; ++c, ++chunk_counter) {
                    if ((int)(chunk_counter & 1u) == parity) {
                        mbar_wait(&empty[stage], phase ^ 1u, 0);
                        uint8_t* st = smem_a + (size_t)stage * STAGE_BYTES;
                        if (lane == 0) {
                            if (part == 0) {
                                if (leader) mbar_arrive_expect_tx(&full[stage], STAGE_BYTES * CG);
                                else mbar_arrive_cluster(&full[stage], 0);
                            }
                            if (!GATHER) tma_load_2d<CG>(&map_a, &full[stage], st + part * A_BYTES, c * CHUNK_K, part * p.a_rows + row0);
                        }
                        __syncwarp();
                        if (GATHER)
                            tma_gather4<CG>(&map_a, &full[stage], st + part * A_BYTES + lane * 512, c * CHUNK_K, rows[0], rows[1],
                                            rows[2], rows[3]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ============================== MMA ISSUER (pair leader only) ==============================
        if (leader) {
            const uint32_t idesc = idesc_f16_f32(TILE_M * CG, NT, false);
            const uint32_t b_base = smem_u32(smem_b);
            int stage = 0;
            uint32_t phase = 0, tile_counter = 0, slab_counter = 0;
            for (int slab = cluster_id; slab < total_slabs; slab += num_clusters, ++slab_counter) {
                int type, row_begin, row_end;
                locate(slab, type, row_begin, row_end);
                const int num_tiles = (row_end - row_begin + TILE_M * CG - 1) / (TILE_M * CG);
                mbar_wait(&b_full, slab_counter & 1u, 5);
                tc_fence_after();
                for (int t = 0; t < num_tiles; ++t, ++tile_counter) {
                    const uint32_t a = tile_counter & 1u, ause = tile_counter >> 1;
                    mbar_wait(&acc_empty[a], (ause & 1u) ^ 1u, 1);
                    tc_fence_after();
                    const uint32_t tmem_acc = tmem_base + a * NT;
                    for (int c = 0; c < CHUNKS; ++c) {
                        mbar_wait(&full[stage], phase, 2);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t a_hi = smem_u32(smem_a + (size_t)stage * STAGE_BYTES), a_lo = a_hi + A_BYTES;
                            const uint32_t b_hi = b_base + (c * 2) * B_CHUNK_BYTES, b_lo = b_hi + B_CHUNK_BYTES;
#pragma unroll
                            for (int kk = 0; kk < CHUNK_K / 16; ++kk) {
                                const uint32_t koff = kk * 32;
                                umma_f16<CG>(tmem_acc, desc_k_sw128(a_hi + koff), desc_k_sw128(b_hi + koff), idesc, (c | kk) ? 1u : 0u);
                                umma_f16<CG>(tmem_acc, desc_k_sw128(a_hi + koff), desc_k_sw128(b_lo + koff), idesc, 1u);
                                umma_f16<CG>(tmem_acc, desc_k_sw128(a_lo + koff), desc_k_sw128(b_hi + koff), idesc, 1u);
                            }
                            tc_commit<CG>(&empty[stage]);
                            if (c == CHUNKS - 1) {
                                tc_commit<CG>(&acc_full[a]);
                                if (t == num_tiles - 1) tc_commit<CG>(&b_empty);  // the weights may be replaced
                            }
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ============================== EPILOGUE ==============================
        const float inv_scale = ((p.amax != nullptr) ? 1.0f / pow2_scale_for(__ldg(p.amax)) : 1.0f) *
                                ((p.amax_b != nullptr) ? 1.0f / pow2_scale_for(__ldg(p.amax_b)) : 1.0f);
        const int lane_base = (warp & 3) * 32;
        float* stage_tile = epi_stage[warp & 3];
        uint32_t tile_counter = 0;
        for (int slab = cluster_id; slab < total_slabs; slab += num_clusters) {
            int type, row_begin, row_end;
            locate(slab, type, row_begin, row_end);
            const int num_tiles = (row_end - row_begin + TILE_M * CG - 1) / (TILE_M * CG);
            const float* brow = p.bias ? p.bias + (size_t)type * p.N : nullptr;
            for (int t = 0; t < num_tiles; ++t, ++tile_counter) {
                const int row0 = row_begin + t * (TILE_M * CG) + (int)cta_rank * TILE_M;
                const uint32_t a = tile_counter & 1u, ause = tile_counter >> 1;
                mbar_wait(&acc_full[a], ause & 1u, 3);
                tc_fence_after();
                const int rows_valid = max(0, min(32, row_end - (row0 + lane_base)));
                float* oblock = p.out + (size_t)(row0 + lane_base) * p.N;
#pragma unroll 1
                for (int j = 0; j < NT / 32; ++j) {
                    float v[32];
                    const int col = j * 32;
                    tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(a * NT + col), v);
                    epi_stage_write(stage_tile, v, lane, inv_scale);
                    epi_store_rows(stage_tile, lane, oblock + col, p.N, rows_valid, brow ? brow + col : nullptr);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive_local(&acc_empty[a]);
                    else mbar_arrive_cluster(&acc_empty[a], 0);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == MMA_WARP) tmem_dealloc<CG>(tmem_base, 2 * NT);
}

// =====================================================================================================================
// Weight gradient:  dW_type[m, col0 + n] += inv_scale * sum over the pair rows p of the segments of that type of
//                   G[p, m] * X[idx[p], n]
//   G = split table of the (pre-scaled) table gradient [2][g_rows][M], X = split table of the node states [2][x_rows][Nin].
//   The reduction runs over pair rows, so both operands are MN-major: a chunk is 64 pair rows, staged as blocks of
//   [64 rows x 64 outputs (128 B)]; G by tiled boxes, X by gather4 (padding rows -> the zero row of X, which makes the
//   rows of G past a slab end harmless).  Work item = (slab of <= SLAB_ROWS pair rows, 128*CG-row m tile, 256-col n tile);
//   partial tiles are added to dW with fp32 REDs (dW pre-zeroed by the caller side of the C ABI).
// =====================================================================================================================

struct WgParams {
    const int* idx;        // [P]
    const float* amax;     // pre-scale source of G (nullable)
    const float* amax_x;   // pre-scale source of X (nullable)
    const int4* units;     // slabs {first pair row, end pair row, weight matrix, -} (bl_segment_units, unit SLAB_ROWS)
    const int* num_units;  // device scalar
    float* d_weight;       // [num_types, M, ld]; this call accumulates into columns [col0, col0 + Nin)
    int M, Nin, ld, col0, g_rows, x_rows;
};

template <int CG, int NT_>
struct WgCfg {
    static constexpr int NT = NT_;                                          // columns of dW per work item (256, or 64 for
                                                                            // the narrow products of the attention backward)
    static constexpr int NTL = NT / CG;
    static constexpr uint32_t BLOCK_BYTES = CHUNK_K * 128;                  // 64 pair rows x 64 outputs, fp16
    static constexpr uint32_t A_BYTES = (TILE_M / 64) * BLOCK_BYTES;        // one part: 128 outputs of G
    static constexpr uint32_t B_BYTES = (NTL / 64) * BLOCK_BYTES;           // one part: NTL outputs of X
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (STAGE_BYTES <= 64 * 1024) ? 3 : 2;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

template <int CG, int NT_>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, const WgParams p) {
    static_assert(NT_ / CG >= 64 && (NT_ / CG) % 64 == 0, "X is staged in 64-column blocks");
    using Cfg = WgCfg<CG, NT_>;
    constexpr int NT = Cfg::NT, NTL = Cfg::NTL, STAGES = Cfg::STAGES;
    constexpr uint32_t BLOCK_BYTES = Cfg::BLOCK_BYTES, A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES,
                       STAGE_BYTES = Cfg::STAGE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float epi_stage[4][EPI_STAGE_FLOATS];  // one 32 x 32 transposition tile per epilogue warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int num_clusters = gridDim.x / CG, cluster_id = blockIdx.x / CG;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], CG);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], CG * 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_g);
        prefetch_tensormap(&map_x);
    }
    if (CG == 2) cluster_sync_all();
    if (warp == MMA_WARP) tmem_alloc<CG>(&tmem_base_smem, 2 * NT);
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int m_tiles = p.M / (TILE_M * CG), n_tiles = p.Nin / NT;
    const int per_slab = m_tiles * n_tiles;
    const int total_work = __ldg(p.num_units) * per_slab;

    // (slab, m tile, n tile) -> pair-row range and output offsets
    auto locate = [&](int work, int& type, int& row_begin, int& row_end, int& m0, int& n0) {
        const int slab = work / per_slab, mn = work - slab * per_slab;
        const int mt = mn / n_tiles, nt = mn - mt * n_tiles;
        const int4 unit = __ldg(p.units + slab);
        type = unit.z; row_begin = unit.x; row_end = unit.y;
        m0 = mt * (TILE_M * CG) + (int)cta_rank * TILE_M;  // this CTA's 128 rows of dW
        n0 = nt * NT;                                      // the pair's 256 columns; this CTA stages NTL of them
    };

    if (warp < 4) {
        // ============================== TMA PRODUCERS (same division of labour as in proj_kernel) ==============================
        const int part = warp & 1, parity = warp >> 1;
        int stage = 0;
        uint32_t phase = 0, chunk_counter = 0;
        const int sub = lane >> 4, grp = lane & 15;  // lane -> (64-column block parity, 4-row group)
        for (int work = cluster_id; work < total_work; work += num_clusters) {
            int type, row_begin, row_end, m0, n0;
            locate(work, type, row_begin, row_end, m0, n0);
            const int num_chunks = (row_end - row_begin + CHUNK_K - 1) / CHUNK_K;
            const int x_col0 = n0 + (int)cta_rank * NTL;
            for (int c = 0; c < num_chunks; ++c, ++chunk_counter) {
                if ((int)(chunk_counter & 1u) == parity) {
                    const int p0 = row_begin + c * CHUNK_K;
                    int rows[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = p0 + grp * 4 + j;
                        rows[j] = ((r < row_end) ? __ldg(p.idx + r) : (p.x_rows - 1)) + part * p.x_rows;  // padding: zero row of X
                    }
                    mbar_wait(&empty[stage], phase ^ 1u, 0);
                    uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
                    if (lane == 0) {
                        if (part == 0) {
                            if (leader) mbar_arrive_expect_tx(&full[stage], STAGE_BYTES * CG);
                            else mbar_arrive_cluster(&full[stage], 0);
                        }
#pragma unroll
                        for (int blk = 0; blk < TILE_M / 64; ++blk)  // G: boxes of 64 outputs x 64 pair rows
                            tma_load_2d<CG>(&map_g, &full[stage], st + part * A_BYTES + blk * BLOCK_BYTES, m0 + blk * 64,
                                            part * p.g_rows + p0);
                    }
                    __syncwarp();
                    if constexpr (NTL >= 128) {
#pragma unroll
                        for (int pass = 0; pass < NTL / 128; ++pass) {  // X: 16 gather4 per 64-column block
                            const int blk = pass * 2 + sub;
                            tma_gather4<CG>(&map_x, &full[stage], st + 2 * A_BYTES + part * B_BYTES + blk * BLOCK_BYTES + grp * 512,
                                            x_col0 + blk * 64, rows[0], rows[1], rows[2], rows[3]);
                        }
                    } else if (sub == 0) {  // a single 64-column block: half the warp issues
                        tma_gather4<CG>(&map_x, &full[stage], st + 2 * A_BYTES + part * B_BYTES + grp * 512, x_col0, rows[0], rows[1],
                                        rows[2], rows[3]);
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == MMA_WARP) {
        // ============================== MMA ISSUER ==============================
        if (leader) {
            const uint32_t idesc = idesc_f16_f32(TILE_M * CG, NT, true);
            int stage = 0;
            uint32_t phase = 0, tile_counter = 0;
            for (int work = cluster_id; work < total_work; work += num_clusters, ++tile_counter) {
                int type, row_begin, row_end, m0, n0;
                locate(work, type, row_begin, row_end, m0, n0);
                const int num_chunks = (row_end - row_begin + CHUNK_K - 1) / CHUNK_K;
                const uint32_t a = tile_counter & 1u, ause = tile_counter >> 1;
                mbar_wait(&acc_empty[a], (ause & 1u) ^ 1u, 1);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + a * NT;
                for (int c = 0; c < num_chunks; ++c) {
                    mbar_wait(&full[stage], phase, 2);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_hi = smem_u32(smem + (size_t)stage * STAGE_BYTES), a_lo = a_hi + A_BYTES;
                        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                        for (int kk = 0; kk < CHUNK_K / 16; ++kk) {
                            const uint32_t koff = kk * 16 * 128;  // 16 pair rows = two 8-row groups of 1024 B
                            const uint64_t dah = desc_mn_sw128(a_hi + koff, BLOCK_BYTES), dal = desc_mn_sw128(a_lo + koff, BLOCK_BYTES);
                            const uint64_t dbh = desc_mn_sw128(b_hi + koff, BLOCK_BYTES), dbl = desc_mn_sw128(b_lo + koff, BLOCK_BYTES);
                            umma_f16<CG>(tmem_acc, dah, dbh, idesc, (c | kk) ? 1u : 0u);
                            umma_f16<CG>(tmem_acc, dah, dbl, idesc, 1u);
                            umma_f16<CG>(tmem_acc, dal, dbh, idesc, 1u);
                        }
                        tc_commit<CG>(&empty[stage]);
                        if (c == num_chunks - 1) tc_commit<CG>(&acc_full[a]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ============================== EPILOGUE: fp32 REDs into dW ==============================
        const float inv_scale = ((p.amax != nullptr) ? 1.0f / pow2_scale_for(__ldg(p.amax)) : 1.0f) *
                                ((p.amax_x != nullptr) ? 1.0f / pow2_scale_for(__ldg(p.amax_x)) : 1.0f);
        const int lane_base = (warp & 3) * 32;
        float* stage_tile = epi_stage[warp & 3];
        uint32_t tile_counter = 0;
        for (int work = cluster_id; work < total_work; work += num_clusters, ++tile_counter) {
            int type, row_begin, row_end, m0, n0;
            locate(work, type, row_begin, row_end, m0, n0);
            const uint32_t a = tile_counter & 1u, ause = tile_counter >> 1;
            mbar_wait(&acc_full[a], ause & 1u, 3);
            tc_fence_after();
            float* wblock = p.d_weight + ((size_t)type * p.M + m0 + lane_base) * p.ld + p.col0 + n0;
#pragma unroll 1
            for (int j = 0; j < NT / 32; ++j) {
                float v[32];
                const int col = j * 32;
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(a * NT + col), v);
                epi_stage_write(stage_tile, v, lane, inv_scale);
                epi_red_rows(stage_tile, lane, wblock + col, p.ld);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive_local(&acc_empty[a]);
                else mbar_arrive_cluster(&acc_empty[a], 0);
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<CG>(tmem_base, 2 * NT);
}

// ------------------------------------------------------------------------------------------------ small kernels
// out[part][r][0:dim]: part 0 = fp16(scale * x), part 1 = fp16(scale * x - part 0); row `rows` (the last) of both parts = 0.
// idx (nullable) gathers rows of x first.  One thread per float4.
__global__ void rows_split_kernel(const float* __restrict__ x, const int* __restrict__ idx, int64_t rows, int dim,
                                  const float* __restrict__ amax, __half* __restrict__ out) {
    const int64_t vec_per_row = dim / 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (rows + 1) * vec_per_row) return;
    const int64_t r = gid / vec_per_row, c4 = gid - r * vec_per_row;
    uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
    if (r < rows) {
        const int64_t src = idx ? (int64_t)__ldg(idx + r) : r;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + src * dim) + c4);
        if (amax != nullptr) {
            const float s = pow2_scale_for(__ldg(amax));
            v.x = fminf(fmaxf(v.x * s, -65000.f), 65000.f); v.y = fminf(fmaxf(v.y * s, -65000.f), 65000.f);
            v.z = fminf(fmaxf(v.z * s, -65000.f), 65000.f); v.w = fminf(fmaxf(v.w * s, -65000.f), 65000.f);
        }
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - b01.x, v.y - b01.y), l23 = __floats2half2_rn(v.z - b23.x, v.w - b23.y);
        hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
        lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
    }
    uint2* o = reinterpret_cast<uint2*>(out);
    o[r * vec_per_row + c4] = hi;
    o[(rows + 1 + r) * vec_per_row + c4] = lo;
}

// prefix[s] = sum over s' < s of ceil((seg_ptr[s'+1] - seg_ptr[s']) / unit); one block, num_segs + 1 outputs
__global__ void unit_prefix_kernel(const int* __restrict__ seg_ptr, int num_segs, int unit, int* __restrict__ prefix) {
    __shared__ int carry;
    __shared__ int warp_sums[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < num_segs; base += blockDim.x) {
        const int s = base + tid;
        int v = 0;
        if (s < num_segs) v = (seg_ptr[s + 1] - seg_ptr[s] + unit - 1) / unit;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL_MASK, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;  // inclusive
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - v;
        if (s < num_segs) prefix[s] = before;
        __syncthreads();
        if (tid == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) prefix[num_segs] = carry;
}

// units[u] = {first row, end row, matrix id, segment} of work unit u (<= `unit` consecutive rows of one segment), u < prefix[num_segs]
__global__ void unit_table_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ seg_type, const int* __restrict__ prefix,
                                  int num_segs, int unit, int64_t max_units, int4* __restrict__ units, int* __restrict__ count) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int total = prefix[num_segs];
    if (u == 0) *count = (int)min((int64_t)total, max_units);
    if (u >= total || u >= max_units) return;
    const int s = find_segment(prefix, num_segs, (int)u);
    const int row_begin = seg_ptr[s] + ((int)u - prefix[s]) * unit;
    units[u] = make_int4(row_begin, min(row_begin + unit, seg_ptr[s + 1]), seg_type ? seg_type[s] : s, s);
}

// ------------------------------------------------------------------------------------------------ host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// fp16 matrix [rows, cols] (cols contiguous) -> tiled tensor map with a box of `box_cols` x `box_rows`, SWIZZLE_128B
static int make_map_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return BL_ERR_UNSUPPORTED;
    if (((uintptr_t)base & 15u) != 0 || (cols * 2) % 16 != 0 || box_cols * 2 > 128 || box_rows > 256) return BL_ERR_INVALID_ARGUMENT;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
        return BL_ERR_CUDA;
    }
    return BL_OK;
}

static int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : fallback;
}
// 2 = CTA pairs (tcgen05 cta_group::2), 1 = single-CTA kernels; BUGLAB_B200_TMA_CG overrides (diagnostics)
static int default_cg() {
    static const int cg = []() {
        const int v = env_int("BUGLAB_B200_TMA_CG", 2);
        return (v == 1) ? 1 : 2;
    }();
    return cg;
}

template <typename Kernel, typename... Args>
static int launch(Kernel kernel, int cg, int grid, uint32_t smem, cudaStream_t stream, const char* what, Args... args) {
    int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
    if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    rc = check_cuda(cudaLaunchKernelEx(&cfg, kernel, args...), what);
    if (rc) return rc;
    return check_launch(what);
}

}  // namespace tg
}  // namespace bl

using namespace bl;

/* fp32 rows -> fp16 hi/lo split table [2][rows + 1][dim] (the extra last row of each part is zero). */
extern "C" int bl_rows_split_f16(const float* x, const int32_t* idx, int64_t rows, int32_t dim, const float* amax, void* out,
                                 bl_stream_t stream) {
    if (rows < 0 || dim <= 0 || dim % 4 != 0) return BL_ERR_INVALID_ARGUMENT;
    const int64_t work = (rows + 1) * (dim / 4);
    tg::rows_split_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(x, idx, rows, dim, amax, (__half*)out);
    return check_launch("bl_rows_split_f16");
}

extern "C" int bl_tma_gemm_supported(int32_t n_out, int32_t k_in) {
    return (k_in % tg::CHUNK_K == 0) && (n_out == 64 || n_out == 128 || (n_out % 256 == 0 && n_out <= 4096)) && k_in <= 4096 &&
           tg::encode_fn() != nullptr;
}

/* prefix[s] = sum_{s' < s} ceil(rows(s') / unit): the tile / slab tables the GEMMs walk (device to device, no sync). */
extern "C" int bl_segment_unit_prefix(const int32_t* seg_ptr, int32_t num_segs, int32_t unit, int32_t* prefix, bl_stream_t stream) {
    if (num_segs <= 0 || unit <= 0) return BL_ERR_INVALID_ARGUMENT;
    tg::unit_prefix_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(seg_ptr, num_segs, unit, prefix);
    return check_launch("bl_segment_unit_prefix");
}

/* Work-unit table of the TMA GEMMs: units[u] = {first row, end row, weight matrix (seg_type[s] or s), segment} for every
 * run of <= `unit` consecutive pair rows of one segment, in row order; *count = number of units (<= max_units, which must
 * be >= num_rows / unit + num_segs).  prefix_ws: scratch of num_segs + 1 ints.  Device to device, no sync. */
extern "C" int bl_segment_units(const int32_t* seg_ptr, const int32_t* seg_type, int32_t num_segs, int32_t unit, int64_t max_units,
                                int32_t* prefix_ws, void* units, int32_t* count, bl_stream_t stream_) {
    if (num_segs <= 0 || unit <= 0 || max_units <= 0) return BL_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = (cudaStream_t)stream_;
    tg::unit_prefix_kernel<<<1, 1024, 0, stream>>>(seg_ptr, num_segs, unit, prefix_ws);
    tg::unit_table_kernel<<<grid_for(max_units, 256), 256, 0, stream>>>(seg_ptr, seg_type, prefix_ws, num_segs, unit, max_units,
                                                                        (int4*)units, count);
    return check_launch("bl_segment_units");
}

extern "C" int bl_tma_tile_rows(void) { return tg::TILE_M * tg::default_cg(); }
extern "C" int bl_tma_slab_rows(void) { return tg::SLAB_ROWS; }

/* 1 when bl_tma_project_stationary covers the shape (CTA pairs, 256 x 256) and BUGLAB_B200_TMA_BSTAT=1 (default 0: on
 * B200 the streaming kernel measured faster — 2.51 vs 2.57 ms forward, 2.52 vs 2.86 ms backward-input at 5 M rows — because
 * the kernels were bound by the epilogue's store pattern, not by L2->SM weight traffic; profiles/r2_gemm_microbench*.jsonl) */
extern "C" int bl_tma_project_stationary_supported(int32_t n_out, int32_t k_in) {
    static const int enabled = tg::env_int("BUGLAB_B200_TMA_BSTAT", 0);
    return enabled && tg::default_cg() == 2 && n_out == 256 && k_in == 256 && tg::encode_fn() != nullptr;
}

/* Same product as bl_tma_project with the weights held in shared memory per slab of <= bl_tma_slab_rows() pair rows
 * (slab_ptr = bl_segment_unit_prefix(seg_ptr, bl_tma_slab_rows()), max_slabs = an upper bound of its last entry). */
extern "C" int bl_tma_project_stationary(const void* a_split, int64_t a_rows, const int32_t* idx, const void* wparts, const float* bias,
                                         const float* amax, const float* amax_b, const void* slabs, const int32_t* num_slabs,
                                         int32_t num_types, int64_t num_rows, int64_t max_slabs, int32_t n_out,
                                         int32_t k_in, float* out, bl_stream_t stream_) {
    if (slabs == nullptr || num_slabs == nullptr || num_types <= 0 || num_rows < 0 || a_rows <= 0 || a_rows > (1ll << 30)) return BL_ERR_INVALID_ARGUMENT;
    if (!bl_tma_project_stationary_supported(n_out, k_in)) return BL_ERR_UNSUPPORTED;
    if (num_rows == 0 || max_slabs <= 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    CUtensorMap map_a, map_b;
    int rc = tg::make_map_f16(&map_a, a_split, (uint64_t)(2 * a_rows), (uint64_t)k_in, tg::CHUNK_K, idx ? 1u : (uint32_t)tg::TILE_M);
    if (rc) return rc;
    rc = tg::make_map_f16(&map_b, wparts, (uint64_t)num_types * 2 * n_out, (uint64_t)k_in, tg::CHUNK_K, (uint32_t)tg::ProjBsCfg::NTL);
    if (rc) return rc;
    tg::ProjParams p{idx, bias, amax, amax_b, (const int4*)slabs, num_slabs, out, n_out, k_in, (int)a_rows};
    int grid = (int)std::min<int64_t>((int64_t)(num_sms() / 2), max_slabs) * 2;
    if (grid < 2) grid = 2;
    if (idx) return tg::launch(tg::proj_bs_kernel<true>, 2, grid, tg::ProjBsCfg::SMEM_BYTES, stream, "bl_tma_project_stationary", map_a, map_b, p);
    return tg::launch(tg::proj_bs_kernel<false>, 2, grid, tg::ProjBsCfg::SMEM_BYTES, stream, "bl_tma_project_stationary", map_a, map_b, p);
}

extern "C" int bl_tma_project(const void* a_split, int64_t a_rows, const int32_t* idx, const void* wparts, const float* bias,
                              const float* amax, const float* amax_b, const void* tiles, const int32_t* num_tiles,
                              int32_t num_types, int64_t num_rows, int64_t max_tiles, int32_t n_out,
                              int32_t k_in, float* out, bl_stream_t stream_) {
    if (tiles == nullptr || num_tiles == nullptr || num_types <= 0 || num_rows < 0 || a_rows <= 0 || a_rows > (1ll << 30)) return BL_ERR_INVALID_ARGUMENT;
    if (!bl_tma_gemm_supported(n_out, k_in)) return BL_ERR_UNSUPPORTED;
    if (num_rows == 0 || max_tiles <= 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cg = tg::default_cg();
    const int nt = (n_out < 256) ? n_out : 256;
    const int ntl = nt / cg;
    CUtensorMap map_a, map_b;
    int rc = tg::make_map_f16(&map_a, a_split, (uint64_t)(2 * a_rows), (uint64_t)k_in, tg::CHUNK_K, idx ? 1u : (uint32_t)tg::TILE_M);
    if (rc) return rc;
    rc = tg::make_map_f16(&map_b, wparts, (uint64_t)num_types * 2 * n_out, (uint64_t)k_in, tg::CHUNK_K, (uint32_t)ntl);
    if (rc) return rc;
    tg::ProjParams p{idx, bias, amax, amax_b, (const int4*)tiles, num_tiles, out, n_out, k_in, (int)a_rows};
    int sms = num_sms();
    int grid = (int)std::min<int64_t>((int64_t)(sms / cg), max_tiles * (n_out / nt)) * cg;
    if (grid < cg) grid = cg;
#define BL_LAUNCH_PROJ(NT_, CG_)                                                                                         \
    (idx ? tg::launch(tg::proj_kernel<NT_, CG_, true>, CG_, grid, tg::ProjCfg<NT_, CG_>::SMEM_BYTES, stream, "bl_tma_project", \
                      map_a, map_b, p)                                                                                   \
         : tg::launch(tg::proj_kernel<NT_, CG_, false>, CG_, grid, tg::ProjCfg<NT_, CG_>::SMEM_BYTES, stream, "bl_tma_project", \
                      map_a, map_b, p))
    if (nt == 64) return cg == 2 ? BL_LAUNCH_PROJ(64, 2) : BL_LAUNCH_PROJ(64, 1);
    if (nt == 128) return cg == 2 ? BL_LAUNCH_PROJ(128, 2) : BL_LAUNCH_PROJ(128, 1);
    return cg == 2 ? BL_LAUNCH_PROJ(256, 2) : BL_LAUNCH_PROJ(256, 1);
#undef BL_LAUNCH_PROJ
}

extern "C" int bl_tma_weight_grad_supported(int32_t m_out, int32_t n_in) {
    if (tg::encode_fn() == nullptr) return 0;
    if (n_in == 64) return m_out % tg::TILE_M == 0 && m_out <= 1024;  // narrow products (attention backward): single CTAs
    return (m_out % 256 == 0) && (n_in % 256 == 0) && m_out <= 4096 && n_in <= 4096;
}

/* d_weight[type, 0:m_out, col0:col0+n_in] = (1/scale) sum over pair rows of G[p,:]^T X[idx[p],:]  (zeroes that block first) */
extern "C" int bl_tma_weight_grad(const void* g_split, int64_t g_rows, const void* x_split, int64_t x_rows, const int32_t* idx,
                                  const float* amax, const float* amax_x, const void* slabs, const int32_t* num_slabs,
                                  int32_t num_types, int64_t num_rows, int64_t max_slabs, int32_t m_out,
                                  int32_t n_in, float* d_weight, int32_t ld, int32_t col0, bl_stream_t stream_) {
    if (slabs == nullptr || num_slabs == nullptr || num_types <= 0 || num_rows < 0 || idx == nullptr || g_rows <= 0 || x_rows <= 0) return BL_ERR_INVALID_ARGUMENT;
    if (!bl_tma_weight_grad_supported(m_out, n_in)) return BL_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_cuda(cudaMemset2DAsync(d_weight + col0, (size_t)ld * sizeof(float), 0, (size_t)n_in * sizeof(float),
                                          (size_t)num_types * m_out, stream), "bl_tma_weight_grad memset");
    if (rc) return rc;
    if (num_rows == 0 || max_slabs <= 0) return BL_OK;
    const int cg = (n_in == 64) ? 1 : tg::default_cg();
    const int nt = (n_in == 64) ? 64 : 256;
    CUtensorMap map_g, map_x;
    rc = tg::make_map_f16(&map_g, g_split, (uint64_t)(2 * g_rows), (uint64_t)m_out, 64, 64);
    if (rc) return rc;
    rc = tg::make_map_f16(&map_x, x_split, (uint64_t)(2 * x_rows), (uint64_t)n_in, 64, 1);
    if (rc) return rc;
    tg::WgParams p{idx, amax, amax_x, (const int4*)slabs, num_slabs, d_weight, m_out, n_in, ld, col0, (int)g_rows, (int)x_rows};
    const int64_t items = max_slabs * (m_out / (tg::TILE_M * cg)) * (n_in / nt);
    int grid = (int)std::min<int64_t>((int64_t)(num_sms() / cg), items) * cg;
    if (grid < cg) grid = cg;
    if (nt == 64) return tg::launch(tg::wgrad_kernel<1, 64>, 1, grid, tg::WgCfg<1, 64>::SMEM_BYTES, stream, "bl_tma_weight_grad", map_g, map_x, p);
    if (cg == 2) return tg::launch(tg::wgrad_kernel<2, 256>, 2, grid, tg::WgCfg<2, 256>::SMEM_BYTES, stream, "bl_tma_weight_grad", map_g, map_x, p);
    return tg::launch(tg::wgrad_kernel<1, 256>, 1, grid, tg::WgCfg<1, 256>::SMEM_BYTES, stream, "bl_tma_weight_grad", map_g, map_x, p);
}
