// Fused gather + fp16-split + grouped GEMM on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   out[p, 0:N] = scale * src[idx[p], 0:Kin] . W_k[0:N, 0:Kin]^T (+ bias_k)      for the pair rows p of type k
//
// This is the hoisted per-type affine of ptgnn's MlpMessagePassingLayer (U = A_k h_src, V = B_k h_tgt + b_k; reference
// call site buglab/models/gnnlayerdefs.py:6-23) and, with src = the table gradient and W transposed, its backward with
// respect to the gathered rows.  Compared with bl_rows_split3_f16 + cublasGemmEx it never materialises the split table
// A3[P, 3*Kin+8] in HBM: the loader gathers the fp32 rows, splits them into fp16 hi/lo parts in registers and writes the
// 128-byte-swizzled K-major smem tiles the UMMA descriptors expect; one gathered 128x64 fp32 chunk feeds 12 MMAs
// (hi.w1, hi.w2, lo.w1 for 4 k-steps of 16), all accumulating into one fp32 TMEM tile of 128 lanes x N columns.
//
// First-generation kernels (round 1).  Since round 2 the projections run on csrc/gemm_tma.cu (operands split once per
// table, TMA-fed, CTA pairs); these remain the path for BUGLAB_B200_TMA=0 and the referee the TMA kernels are checked
// against.  Persistent warp-specialised CTA (see pair_project_tc_v2_kernel); epilogue: tcgen05.ld (32 lanes x 32 columns
// per warp and step) -> + bias -> 128-byte row segments to global.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace bl {
namespace tc {

constexpr int TILE_M = 128;     // pair rows per tile = TMEM lanes
constexpr int CHUNK_K = 64;     // fp16 elements per smem row = 128 bytes = one swizzle atom
constexpr int THREADS = 256;
constexpr int MAX_TYPES = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait: a protocol bug must trap (and fail the launch) within ~2 s instead of hanging the GPU.  test_wait never
// suspends the thread, so the wall-clock bound below is real.
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __noinline__ void mbar_timeout(uint32_t addr, uint32_t parity) {
    printf("buglab_b200: mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n", blockIdx.x, threadIdx.x, addr, parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
            else if (now - t0 > 2000000000ull) mbar_timeout(addr, parity);
        }
    }
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in
// [0,14), leading byte offset (unused for one 128-byte atom along K) in [16,30), stride byte offset = 8 rows * 128 B
// = 1024 >> 4 in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, SWIZZLE_128B: the tile is stored as K rows of 128 bytes (64 fp16 along M/N), 8-row groups 1024 B apart
// (stride byte offset), further 64-wide M/N blocks `lbo_bytes` apart (leading byte offset) — canonical layout
// Swizzle<3,4,3> o ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) of cute::UMMA::make_umma_desc<Major::MN>.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1 at [4,6)), A=B=F16 (0), both K-major, N>>3 at [17,23),
// M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(int m, int n, bool mn_major = false) {
    return (1u << 4) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
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
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// byte offset of (row, 16-byte unit j) inside a [rows x 128 B] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int row, int unit) { return (uint32_t)(row * 128 + ((unit ^ (row & 7)) << 4)); }

// fp32x4 -> fp16 hi/lo parts with packed conversions (cvt.rn.f16x2.f32): ~14 instructions per float4 instead of ~40 —
// the loader, not the tensor pipe, is what bounds these kernels.  SCALED adds the pow2 pre-scale + fp16 saturation.
template <bool SCALED>
__device__ __forceinline__ void split_f32x4(float4 v, float scale, uint2& hi, uint2& lo) {
    if (SCALED) {
        v.x = fminf(fmaxf(v.x * scale, -65000.f), 65000.f); v.y = fminf(fmaxf(v.y * scale, -65000.f), 65000.f);
        v.z = fminf(fmaxf(v.z * scale, -65000.f), 65000.f); v.w = fminf(fmaxf(v.w * scale, -65000.f), 65000.f);
    }
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - b01.x, v.y - b01.y), l23 = __floats2half2_rn(v.z - b23.x, v.w - b23.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}

struct Params {
    const float* src;       // [*, Kin] fp32
    const int* idx;         // [P] rows of src, or nullptr (identity)
    const float* amax;      // device scalar for the pow2 pre-scale, or nullptr
    const __half* wparts;   // [num_types, 2 (hi,lo), N, Kin] fp16, Kin contiguous
    const float* bias;      // [num_types, N] or nullptr
    const int* type_ptr;    // [num_types+1] on the device
    float* out;             // [P, N]
    int num_types, N, Kin;
};

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised pipeline ("v2"; the single-role first version was removed in round 2):
//   warps 0-7   PRODUCERS  two groups of 4 warps; group g owns smem stage g and loads every chunk with
//                          (running chunk index & 1) == g: gather + split A, cp.async B, then arrive on full[g] (128
//                          arrivals) — while one group waits for its loads the other converts and stores
//   warp  12    MMA        waits full[s], issues the 12 MMAs of the chunk, tcgen05.commit -> empty[s];
//                          after a tile's last chunk tcgen05.commit -> acc_full[a]
//   warps 8-11  EPILOGUE   waits acc_full[a], tcgen05.ld -> +bias -> global, arrives on acc_empty[a]    (128 arrivals)
// Two smem stages and TWO TMEM accumulators (2*NT columns): the epilogue of tile i overlaps the main loop of tile i+1.
// ---------------------------------------------------------------------------------------------------------------
constexpr int THREADS_V2 = 416;  // 8 producer warps + 4 epilogue warps + 1 MMA warp

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(THREADS_V2, 1) pair_project_tc_v2_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t A_BYTES = TILE_M * 128;
    constexpr uint32_t B_BYTES = NT * 128;
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    __shared__ uint64_t full[2], empty[2], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int tile_prefix[MAX_TYPES + 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_splits = p.N / NT;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 128);
            mbar_init(&empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int acc = 0;
        for (int k = 0; k < p.num_types; ++k) {
            tile_prefix[k] = acc;
            acc += (p.type_ptr[k + 1] - p.type_ptr[k] + TILE_M - 1) / TILE_M;
        }
        tile_prefix[p.num_types] = acc;
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(2 * NT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int total_tiles = tile_prefix[p.num_types] * n_splits;
    const int num_chunks = p.Kin / CHUNK_K;

    auto locate = [&](int work, int& k, int& row0, int& row_end, int& col0) {
        const int tile = work / n_splits, split = work - tile * n_splits;
        k = 0;
        while (tile >= tile_prefix[k + 1]) ++k;
        row0 = p.type_ptr[k] + (tile - tile_prefix[k]) * TILE_M;
        row_end = p.type_ptr[k + 1];
        col0 = split * NT;
    };

    if (warp < 8) {
        // ======================= PRODUCERS (2 groups x 128 threads) =======================
        const int group = warp >> 2, t = tid & 127;
        const float scale = (p.amax != nullptr) ? pow2_scale_for(__ldg(p.amax)) : 1.0f;
        constexpr int A_ITERS = (TILE_M * CHUNK_K / 4) / 128;  // 16 float4 per thread and chunk, in two batches of 8
        uint8_t* stage = smem + (size_t)group * STAGE_BYTES;
        uint32_t chunk_counter = 0;
        for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
            int k, row0, row_end, col0;
            locate(work, k, row0, row_end, col0);
            int my_rows[A_ITERS];
#pragma unroll
            for (int i = 0; i < A_ITERS; ++i) {
                const int r = row0 + ((i * 128 + t) >> 4);
                my_rows[i] = (r < row_end) ? (p.idx ? __ldg(p.idx + r) : r) : -1;
            }
            const __half* w_hi = p.wparts + ((size_t)(k * 2 + 0) * p.N + col0) * p.Kin;
            const __half* w_lo = p.wparts + ((size_t)(k * 2 + 1) * p.N + col0) * p.Kin;
            for (int c = 0; c < num_chunks; ++c, ++chunk_counter) {
                if ((int)(chunk_counter & 1) != group) continue;  // the other group's chunk
                const uint32_t use = chunk_counter >> 1;
                if (use > 0) mbar_wait(&empty[group], (use - 1) & 1);
#pragma unroll
                for (int i = 0; i < (NT * 8) / 128; ++i) {
                    const int f = i * 128 + t;
                    const int r = f >> 3, u = f & 7;
                    const size_t goff = (size_t)r * p.Kin + c * CHUNK_K + u * 8;
                    cp_async16(smem_u32(stage + 2 * A_BYTES + sw128(r, u)), w_hi + goff);
                    cp_async16(smem_u32(stage + 2 * A_BYTES + B_BYTES + sw128(r, u)), w_lo + goff);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 av[A_ITERS / 2];
#pragma unroll
                    for (int j = 0; j < A_ITERS / 2; ++j) {
                        const int i = half * (A_ITERS / 2) + j;
                        const int f = i * 128 + t;
                        av[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (my_rows[i] >= 0)
                            av[j] = __ldg(reinterpret_cast<const float4*>(p.src + (size_t)my_rows[i] * p.Kin + c * CHUNK_K) + (f & 15));
                    }
#pragma unroll
                    for (int j = 0; j < A_ITERS / 2; ++j) {
                        const int f = (half * (A_ITERS / 2) + j) * 128 + t;
                        const int r = f >> 4, c4 = f & 15;
                        uint2 hp, lp;
                        if (p.amax != nullptr) split_f32x4<true>(av[j], scale, hp, lp); else split_f32x4<false>(av[j], 1.f, hp, lp);
                        const uint32_t off = sw128(r, c4 >> 1) + ((c4 & 1) << 3);
                        *reinterpret_cast<uint2*>(stage + off) = hp;
                        *reinterpret_cast<uint2*>(stage + A_BYTES + off) = lp;
                    }
                }
                cp_async_wait_all();
                fence_async_proxy();
                mbar_arrive(&full[group]);
            }
        }
    } else if (warp == 12) {
        // ======================= MMA ISSUER =======================
        // The whole warp walks the schedule and waits on the barriers together (a lone lane racing ahead of its warp to
        // the final bar.sync is undefined behaviour); only lane 0 issues the single-thread tcgen05 instructions.
        const uint32_t idesc = umma_idesc_f16_f32(TILE_M, NT);
        uint32_t chunk_counter = 0, tile_counter = 0;
        for (int work = blockIdx.x; work < total_tiles; work += gridDim.x, ++tile_counter) {
            const int a = tile_counter & 1;
            const uint32_t ause = tile_counter >> 1;
            if (ause > 0) mbar_wait(&acc_empty[a], (ause - 1) & 1);  // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(a * NT);
            for (int c = 0; c < num_chunks; ++c, ++chunk_counter) {
                const int s = chunk_counter & 1;
                const uint32_t use = chunk_counter >> 1;
                mbar_wait(&full[s], use & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                    for (int kk = 0; kk < CHUNK_K / 16; ++kk) {
                        const uint32_t koff = kk * 32;
                        umma_f16(tmem_acc, umma_desc_sw128(a_hi + koff), umma_desc_sw128(b_hi + koff), idesc, (c | kk) ? 1u : 0u);
                        umma_f16(tmem_acc, umma_desc_sw128(a_hi + koff), umma_desc_sw128(b_lo + koff), idesc, 1u);
                        umma_f16(tmem_acc, umma_desc_sw128(a_lo + koff), umma_desc_sw128(b_hi + koff), idesc, 1u);
                    }
                    tc_commit(&empty[s]);
                    if (c == num_chunks - 1) tc_commit(&acc_full[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ======================= EPILOGUE (warps 8-11, 128 threads) =======================
        uint32_t tile_counter = 0;
        const int lane_base = (warp & 3) * 32;
        for (int work = blockIdx.x; work < total_tiles; work += gridDim.x, ++tile_counter) {
            int k, row0, row_end, col0;
            locate(work, k, row0, row_end, col0);
            const int a = tile_counter & 1;
            const uint32_t ause = tile_counter >> 1;
            mbar_wait(&acc_full[a], ause & 1);
            tc_fence_after();
            const int r = lane_base + lane;
            const bool valid = (row0 + r) < row_end;
            float* orow = p.out + (size_t)(row0 + r) * p.N + col0;
            const float* brow = p.bias ? p.bias + (size_t)k * p.N + col0 : nullptr;
#pragma unroll 1
            for (int j = 0; j < NT / 32; ++j) {
                float v[32];
                const int col = j * 32;
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(a * NT + col), v);
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        if (brow) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(brow + col) + q);
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        reinterpret_cast<float4*>(orow + col)[q] = o;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[a]);
        }
    }
    __syncthreads();
    if (warp == 12) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * NT));
    }
}

// wparts[k, part, n, kin] = hi/lo of W[k, n, col0 + kin]              (transposed == 0, W is [K, N, ld])
//                         = hi/lo of W[k, kin, col0 + n]              (transposed != 0, W is [K, Kin, ld])
__global__ void weight_parts_kernel(const float* __restrict__ W, int num_types, int N, int Kin, int ld, int col0,
                                    int transposed, const float* __restrict__ amax, __half* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)num_types * N * Kin) return;
    const int kin = (int)(gid % Kin);
    const int n = (int)((gid / Kin) % N);
    const int64_t k = gid / ((int64_t)Kin * N);
    // optional power-of-two pre-scale: without it the lo parts of typical weights (|w| ~ 0.05 -> lo ~ 1e-5) are fp16
    // SUBNORMALS (spacing 6e-8), i.e. the pair hi+lo carries ~17 significant bits instead of 22
    const float scale = (amax != nullptr) ? pow2_scale_for(__ldg(amax)) : 1.0f;
    const float w = scale * (transposed ? W[(k * Kin + kin) * ld + col0 + n] : W[(k * N + n) * ld + col0 + kin]);
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    out[((k * 2 + 0) * N + n) * Kin + kin] = hi;
    out[((k * 2 + 1) * N + n) * Kin + kin] = lo;
}


// ---------------------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  dW_k[m, n] (+)= (1/s) * sum_{p in type k} (s*g[p, m]) * x[idx[p], n]
// (dA_k = dU^T h_src, dB_k = dV^T h_tgt of the hoisted affine).  The reduction runs over PAIR ROWS, i.e. both operands
// are naturally "M/N-major" (a row of g or x is contiguous along the OUTPUT dimension), so the loader copies rows
// exactly like the projection kernel (128-bit gathers, fp16 hi/lo split, 8-byte swizzled stores) and the MMAs use
// MN-major descriptors.  Work item = (type, slab of WG_ROWS_PER_ITEM pair rows, 128-row m tile, 256-column n tile);
// the fp32 TMEM partial is added to dW with REDs (dW pre-zeroed).  The slab length also bounds the accumulation chain:
// tensor-core accumulation truncates, so very long chains drift (measured 3e-5 relative after 1500 MMAs).
constexpr int WG_ROWS_PER_ITEM = 4096;

struct WgParams {
    const float* g;        // [P, M]   table gradient (dU or dV), fp32
    const float* x;        // [*, Nin] node states, gathered through idx
    const int* idx;        // [P]
    const float* amax;     // pow2 pre-scale source for g (nullable)
    const int* type_ptr;   // [num_types + 1], device
    float* d_weight;       // [num_types, M, ld]; this call fills columns [col0, col0 + Nin)
    int num_types, M, Nin, ld, col0;
};

__global__ void __launch_bounds__(THREADS, 1) pair_weight_grad_tc_kernel(const WgParams p) {
    constexpr int NT = 256;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // stage: A_hi | A_lo  (each 2 blocks of [64 rows x 128 B])  |  B_hi | B_lo  (each 4 blocks of [64 rows x 128 B])
    constexpr uint32_t BLOCK_BYTES = CHUNK_K * 128;           // one 64(pair rows) x 64(outputs) fp16 block
    constexpr uint32_t A_BYTES = (TILE_M / 64) * BLOCK_BYTES;  // 16 KB
    constexpr uint32_t B_BYTES = (NT / 64) * BLOCK_BYTES;      // 32 KB
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    __shared__ uint64_t mbar[2];
    __shared__ uint64_t mbar_acc;
    __shared__ uint32_t tmem_base_smem;
    __shared__ int slab_prefix[MAX_TYPES + 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_tiles = p.M / TILE_M, n_tiles = p.Nin / NT;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int acc = 0;
        for (int k = 0; k < p.num_types; ++k) {
            slab_prefix[k] = acc;
            acc += (p.type_ptr[k + 1] - p.type_ptr[k] + WG_ROWS_PER_ITEM - 1) / WG_ROWS_PER_ITEM;
        }
        slab_prefix[p.num_types] = acc;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(NT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int total_items = slab_prefix[p.num_types] * m_tiles * n_tiles;
    const float scale = (p.amax != nullptr) ? pow2_scale_for(__ldg(p.amax)) : 1.0f;
    const float inv_scale = 1.0f / scale;
    const uint32_t idesc = umma_idesc_f16_f32(TILE_M, NT, /*mn_major=*/true);

    uint32_t commits[2] = {0, 0};
    uint32_t acc_commits = 0;

    for (int work = blockIdx.x; work < total_items; work += gridDim.x) {
        const int slab = work / (m_tiles * n_tiles);
        const int mn = work - slab * (m_tiles * n_tiles);
        const int mt = mn / n_tiles, nt = mn - mt * n_tiles;
        int k = 0;
        while (slab >= slab_prefix[k + 1]) ++k;
        const int row_begin = p.type_ptr[k] + (slab - slab_prefix[k]) * WG_ROWS_PER_ITEM;
        const int row_end = min(row_begin + WG_ROWS_PER_ITEM, p.type_ptr[k + 1]);
        const int num_chunks = (row_end - row_begin + CHUNK_K - 1) / CHUNK_K;
        const int m0 = mt * TILE_M, n0 = nt * NT;

        for (int c = 0; c < num_chunks; ++c) {
            const int s = c & 1;
            uint8_t* stage = smem + (size_t)s * STAGE_BYTES;
            const int p0 = row_begin + c * CHUNK_K;
            if (commits[s] > 0) mbar_wait(&mbar[s], (commits[s] - 1) & 1);
            // ---- loads first (all in flight), then split + swizzled stores ----
            // A: 64 pair rows x 128 outputs of g  = 2048 float4, 8 per thread;  f -> (row = f / 32, float4 col = f % 32)
            // B: 64 pair rows x 256 outputs of x  = 4096 float4, 16 per thread; f -> (row = f / 64, float4 col = f % 64)
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int f = i * THREADS + tid;
                const int r = p0 + (f >> 5);
                av[i] = (r < row_end) ? __ldg(reinterpret_cast<const float4*>(p.g + (size_t)r * p.M + m0) + (f & 31))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 bv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int f = i * THREADS + tid;
                const int r = p0 + (f >> 6);
                bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < row_end) {
                    const int src_row = __ldg(p.idx + r);
                    bv[i] = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)src_row * p.Nin + n0) + (f & 63));
                }
            }
            auto store_split = [&](const float4 v, float sc, uint8_t* hi_base, uint8_t* lo_base, int row, int col4) {
                // col4 = float4 index along the output dimension; 16 float4 per 64-wide block
                uint2 hp, lp;
                if (sc != 1.0f) split_f32x4<true>(v, sc, hp, lp); else split_f32x4<false>(v, 1.f, hp, lp);
                const int block = col4 >> 4, c4 = col4 & 15;
                const uint32_t off = block * BLOCK_BYTES + sw128(row, c4 >> 1) + ((c4 & 1) << 3);
                *reinterpret_cast<uint2*>(hi_base + off) = hp;
                *reinterpret_cast<uint2*>(lo_base + off) = lp;
            };
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int f = i * THREADS + tid;
                store_split(av[i], scale, stage, stage + A_BYTES, f >> 5, f & 31);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int f = i * THREADS + tid;
                store_split(bv[i], 1.0f, stage + 2 * A_BYTES, stage + 2 * A_BYTES + B_BYTES, f >> 6, f & 63);
            }
            fence_async_proxy();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = smem_u32(stage), a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int kk = 0; kk < CHUNK_K / 16; ++kk) {
                    const uint32_t koff = kk * 16 * 128;  // 16 pair rows = two 8-row groups of 1024 B
                    const uint64_t dah = umma_desc_mn_sw128(a_hi + koff, BLOCK_BYTES), dal = umma_desc_mn_sw128(a_lo + koff, BLOCK_BYTES);
                    const uint64_t dbh = umma_desc_mn_sw128(b_hi + koff, BLOCK_BYTES), dbl = umma_desc_mn_sw128(b_lo + koff, BLOCK_BYTES);
                    umma_f16(tmem_base, dah, dbh, idesc, (c | kk) ? 1u : 0u);
                    umma_f16(tmem_base, dah, dbl, idesc, 1u);
                    umma_f16(tmem_base, dal, dbh, idesc, 1u);
                }
                tc_commit(&mbar[s]);
                if (c == num_chunks - 1) tc_commit(&mbar_acc);
            }
            commits[s] += 1;
        }
        // ---- epilogue: TMEM partial [128 m x 256 n] -> scaled fp32 REDs into dW_k ----
        if (num_chunks > 0) {
            mbar_wait(&mbar_acc, acc_commits & 1);
            acc_commits += 1;
            tc_fence_after();
            const int lane_base = (warp & 3) * 32;
            const int m = m0 + lane_base + lane;
            float* wrow = p.d_weight + ((size_t)k * p.M + m) * p.ld + p.col0 + n0;
            const int cbase = (warp >> 2) * (NT / 2);
#pragma unroll 1
            for (int j = 0; j < (NT / 2) / 32; ++j) {
                float v[32];
                const int col = cbase + j * 32;
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)col, v);
#pragma unroll
                for (int q = 0; q < 32; ++q) atomicAdd(wrow + col + q, v[q] * inv_scale);
            }
            tc_fence_before();
        }
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NT));
    }
}

}  // namespace tc
}  // namespace bl

using namespace bl;

extern "C" int bl_weight_parts_f16(const float* weight, int32_t num_types, int32_t n_out, int32_t k_in, int32_t ld,
                                   int32_t col0, int32_t transposed, const float* amax, void* parts, bl_stream_t stream) {
    if (num_types <= 0 || n_out <= 0 || k_in <= 0) return BL_ERR_INVALID_ARGUMENT;
    tc::weight_parts_kernel<<<grid_for((int64_t)num_types * n_out * k_in, 256), 256, 0, (cudaStream_t)stream>>>(
        weight, num_types, n_out, k_in, ld, col0, transposed, amax, (__half*)parts);
    return check_launch("bl_weight_parts_f16");
}

extern "C" int bl_pair_project_tc_supported(int32_t n_out, int32_t k_in) {
    return (k_in % tc::CHUNK_K == 0) && (n_out == 128 || (n_out % 256 == 0 && n_out <= 1024));
}

extern "C" int bl_pair_project_tc(const float* src, const int32_t* idx, const float* amax, const void* parts,
                                  const float* bias, const int32_t* type_ptr, int32_t num_types, int64_t num_rows,
                                  int32_t n_out, int32_t k_in, float* out, bl_stream_t stream_) {
    if (num_types <= 0 || num_types > tc::MAX_TYPES || num_rows < 0) return BL_ERR_INVALID_ARGUMENT;
    if (!bl_pair_project_tc_supported(n_out, k_in)) return BL_ERR_UNSUPPORTED;
    if (num_rows == 0) return BL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    tc::Params p{src, idx, amax, (const __half*)parts, bias, type_ptr, out, num_types, n_out, k_in};
    const int64_t max_tiles = (num_rows + tc::TILE_M - 1) / tc::TILE_M + num_types;
    // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: set it on every call (cheap) and check it
    if (n_out == 128) {
        constexpr uint32_t smem = 2 * (2 * 128 * 128 + 2 * 128 * 128) + 1024;
        int rc = check_cuda(cudaFuncSetAttribute(tc::pair_project_tc_v2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                            "bl_pair_project_tc attribute");
        if (rc) return rc;
        tc::pair_project_tc_v2_kernel<128><<<(int)std::min<int64_t>(num_sms(), max_tiles), tc::THREADS_V2, smem, stream>>>(p);
    } else {
        constexpr uint32_t smem = 2 * (2 * 128 * 128 + 2 * 256 * 128) + 1024;
        int rc = check_cuda(cudaFuncSetAttribute(tc::pair_project_tc_v2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                            "bl_pair_project_tc attribute");
        if (rc) return rc;
        tc::pair_project_tc_v2_kernel<256><<<(int)std::min<int64_t>(num_sms(), max_tiles * (n_out / 256)), tc::THREADS_V2, smem, stream>>>(p);
    }
    return check_launch("bl_pair_project_tc");
}

extern "C" int bl_pair_weight_grad_tc_supported(int32_t m_out, int32_t n_in) {
    return (m_out % tc::TILE_M == 0) && (n_in % 256 == 0) && m_out <= 1024 && n_in <= 1024;
}

/* d_weight[k, 0:m_out, col0:col0+n_in] = sum over the pair rows p of type k of g[p, :]^T x[idx[p], :] */
extern "C" int bl_pair_weight_grad_tc(const float* g, const float* x, const int32_t* idx, const float* amax,
                                      const int32_t* type_ptr, int32_t num_types, int64_t num_rows, int32_t m_out,
                                      int32_t n_in, float* d_weight, int32_t ld, int32_t col0, bl_stream_t stream_) {
    if (num_types <= 0 || num_types > tc::MAX_TYPES || num_rows < 0 || idx == nullptr) return BL_ERR_INVALID_ARGUMENT;
    if (!bl_pair_weight_grad_tc_supported(m_out, n_in)) return BL_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    // zero the destination block (columns [col0, col0+n_in) of every [m_out, ld] matrix): partials are added with REDs
    int rc = check_cuda(cudaMemset2DAsync(d_weight + col0, (size_t)ld * sizeof(float), 0, (size_t)n_in * sizeof(float),
                                          (size_t)num_types * m_out, stream), "bl_pair_weight_grad_tc memset");
    if (rc) return rc;
    if (num_rows == 0) return BL_OK;
    tc::WgParams p{g, x, idx, amax, type_ptr, d_weight, num_types, m_out, n_in, ld, col0};
    constexpr uint32_t smem = 2 * (2 * 128 * 128 + 2 * 256 * 128) + 1024;
    rc = check_cuda(cudaFuncSetAttribute(tc::pair_weight_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                    "bl_pair_weight_grad_tc attribute");
    if (rc) return rc;
    const int64_t items = ((num_rows + tc::WG_ROWS_PER_ITEM - 1) / tc::WG_ROWS_PER_ITEM + num_types) * (m_out / tc::TILE_M) * (n_in / 256);
    const int grid = (int)std::min<int64_t>(num_sms(), items);
    tc::pair_weight_grad_tc_kernel<<<grid, tc::THREADS, smem, stream>>>(p);
    return check_launch("bl_pair_weight_grad_tc");
}
