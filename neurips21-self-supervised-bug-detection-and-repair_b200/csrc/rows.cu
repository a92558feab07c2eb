// Row gather and its deterministic backward (CSR segmented row sum).  See include/buglab_b200.h.
// The gather is the `index_select h[src] / h[tgt]` of ptgnn's MlpMessagePassingLayer.forward
// (reference call site buglab/models/gnnlayerdefs.py:6-23) applied to unique (type,node) pairs.
#include "common.cuh"

namespace bl {

__global__ void __launch_bounds__(256)
rows_gather_kernel(const float4* __restrict__ table, const int* __restrict__ idx, int64_t num_rows,
                   int dim4, float4* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= num_rows * dim4) return;
    const int64_t r = gid / dim4;
    const int c = (int)(gid - r * dim4);
    out[gid] = __ldg(table + (size_t)__ldg(idx + r) * dim4 + c);
}

__global__ void __launch_bounds__(256)
rows_segment_sum_kernel(const float4* __restrict__ a_rows, const int* __restrict__ a_ptr,
                        const int* __restrict__ a_idx, const float4* __restrict__ b_rows,
                        const int* __restrict__ b_ptr, const int* __restrict__ b_idx,
                        int64_t num_nodes, int dim4, int accumulate, const float* __restrict__ amax,
                        float4* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= num_nodes * dim4) return;
    const int64_t n = gid / dim4;
    const int c = (int)(gid - n * dim4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const int beg = __ldg(a_ptr + n), end = __ldg(a_ptr + n + 1);
        for (int q = beg; q < end; ++q) {
            const float4 v = __ldg(a_rows + (size_t)__ldg(a_idx + q) * dim4 + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (b_rows != nullptr) {
        const int beg = __ldg(b_ptr + n), end = __ldg(b_ptr + n + 1);
        for (int q = beg; q < end; ++q) {
            const float4 v = __ldg(b_rows + (size_t)__ldg(b_idx + q) * dim4 + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (amax != nullptr) {  // rows were produced from a pow2-scaled table (fp16 split): undo exactly
        const float inv = 1.0f / pow2_scale_for(__ldg(amax));
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    if (accumulate) {
        const float4 o = out[gid];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    out[gid] = acc;
}

}  // namespace bl

using namespace bl;

extern "C" int bl_rows_gather(const float* table, const int32_t* idx, int64_t num_rows, int32_t dim,
                              float* out, bl_stream_t stream) {
    if (num_rows < 0 || dim <= 0 || (dim & 3)) return BL_ERR_INVALID_ARGUMENT;
    if (num_rows == 0) return BL_OK;
    const int dim4 = dim / 4;
    rows_gather_kernel<<<grid_for(num_rows * dim4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)table, idx, num_rows, dim4, (float4*)out);
    return check_launch("bl_rows_gather");
}

extern "C" int bl_rows_segment_sum(const float* a_rows, const int32_t* a_ptr, const int32_t* a_idx,
                                   const float* b_rows, const int32_t* b_ptr, const int32_t* b_idx,
                                   int64_t num_nodes, int32_t dim, int32_t accumulate, const float* amax,
                                   float* out, bl_stream_t stream) {
    if (num_nodes < 0 || dim <= 0 || (dim & 3)) return BL_ERR_INVALID_ARGUMENT;
    if (num_nodes == 0) return BL_OK;
    const int dim4 = dim / 4;
    rows_segment_sum_kernel<<<grid_for(num_nodes * dim4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)a_rows, a_ptr, a_idx, (const float4*)b_rows, b_ptr, b_idx, num_nodes, dim4,
        accumulate, amax, (float4*)out);
    return check_launch("bl_rows_segment_sum");
}
