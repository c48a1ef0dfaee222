// Prefill dequant-GEMM on tcgen05/TMEM -- placeholder until the kernel lands (returns
// MXQ_E_UNSUPPORTED so callers fail loudly).
#include "common.cuh"

extern "C" size_t mxq_gemm_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }

extern "C" int mxq_gemm(const void*, mxq_packed_t, void*, int64_t, int64_t, int64_t, void*, size_t,
                        void*) {
  return MXQ_E_UNSUPPORTED;
}
