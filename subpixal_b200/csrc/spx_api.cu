// ABI version and error strings.
#include <cuda_runtime.h>
#include "../../include/subpixal_b200.h"

extern "C" int spx_abi_version(void) { return SPX_ABI_VERSION; }

extern "C" const char *spx_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case SPX_E_BADARG: return "bad argument (null pointer, negative size or bad enum)";
        case SPX_E_UNSUPPORTED: return "unsupported size (cutout must be 3..256 px, fit box <= 15)";
        case SPX_E_WORKSPACE: return "workspace missing or too small";
        case SPX_E_DTYPE: return "unsupported element type";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}
