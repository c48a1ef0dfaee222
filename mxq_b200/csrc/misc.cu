// Version / error strings of the C ABI.
#include "common.cuh"

extern "C" int mxq_version(void) { return MXQ_VERSION; }

extern "C" const char* mxq_error_string(int code) {
  switch (code) {
    case MXQ_OK: return "ok";
    case MXQ_E_NULL: return "required pointer is NULL";
    case MXQ_E_SHAPE: return "shape or divisibility requirement violated";
    case MXQ_E_DTYPE: return "unknown dtype";
    case MXQ_E_ALIGN: return "pointer is not 16-byte aligned";
    case MXQ_E_UNSUPPORTED: return "configuration not supported";
    case MXQ_E_WORKSPACE: return "workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}
