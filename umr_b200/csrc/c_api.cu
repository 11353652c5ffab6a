// c_api.cu -- error strings / version for the C ABI (include/umr_b200.h).
#include <cuda_runtime.h>

#include "umr_b200.h"

extern "C" const char* umr_error_string(int code) {
    switch (code) {
        case UMR_OK: return "ok";
        case UMR_ERR_UNSUPPORTED: return "mode not supported by the sm_100a kernels";
        case UMR_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or misaligned buffer)";
        case UMR_ERR_TOO_LARGE: return "size exceeds a compiled limit";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" int umr_version(void) { return 100; }
