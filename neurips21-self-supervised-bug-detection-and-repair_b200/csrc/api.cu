// Version and error reporting of the buglab_b200 C ABI.
#include <string.h>

#include "common.cuh"

namespace bl {
static char g_cuda_error[512] = "no CUDA error recorded";
void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_error, sizeof(g_cuda_error), "CUDA error in %s: %s (%s)", where, cudaGetErrorName(e),
             cudaGetErrorString(e));
}
}  // namespace bl

extern "C" int bl_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* bl_error_string(int code) {
    switch (code) {
        case BL_OK: return "ok";
        case BL_ERR_INVALID_ARGUMENT: return "invalid argument (size/alignment/range)";
        case BL_ERR_CUDA: return bl::g_cuda_error;
        case BL_ERR_WORKSPACE_TOO_SMALL: return "workspace too small (see bl_plan_workspace_bytes)";
        case BL_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error code";
    }
}
