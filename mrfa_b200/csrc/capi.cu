// ABI bookkeeping entry points.
#include "common.cuh"

extern "C" int mrfa_abi_version(void) { return MRFA_B200_ABI_VERSION; }

extern "C" const char* mrfa_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case MRFA_E_BADARG: return "mrfa_b200: bad argument (null pointer, non-positive extent or unsupported enum)";
    case MRFA_E_SHAPE: return "mrfa_b200: shape outside what the kernel is specialised for";
    case MRFA_E_ALIGN: return "mrfa_b200: pointer not 32-byte aligned";
    case MRFA_E_DRIVER: return "mrfa_b200: cuTensorMapEncodeTiled unavailable or failed";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "mrfa_b200: unknown error";
  }
}
