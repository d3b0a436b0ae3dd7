// dct_abi.cu -- library-level entry points of include/dct_b200.h
#include "dct_common.cuh"

#include <cstdlib>
#include <cstring>

namespace dct {
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("DCT_B200_PDL");
        return !(e != nullptr && std::strcmp(e, "0") == 0);
    }();
    return on;
}
}  // namespace dct

using namespace dct;

extern "C" int dct_abi_version(void) { return DCT_ABI_VERSION; }

extern "C" const char* dct_error_string(int code) {
    switch (code) {
        case DCT_OK: return "ok";
        case DCT_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size, or inconsistent options)";
        case DCT_ERR_UNSUPPORTED: return "unsupported shape (K > 8, C > 64 or B > 65535)";
        case DCT_ERR_MISALIGNED: return "misaligned pointer";
        case DCT_ERR_CUDA: return "CUDA launch failed (see dct_last_cuda_error)";
        case DCT_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
        default: return "unknown error code";
    }
}

extern "C" const char* dct_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda_error); }

extern "C" int dct_device_check(int ordinal) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || ordinal < 0 || ordinal >= n) { g_last_cuda_error = e; return DCT_ERR_NO_DEVICE; }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ordinal);
    if (e != cudaSuccess || major != 10) { g_last_cuda_error = e; return DCT_ERR_NO_DEVICE; }
    return DCT_OK;
}

extern "C" size_t dct_workspace_bytes(void) { return sizeof(Workspace); }
