// AWQ uniform 4-bit prefill GEMM (sm_100a).
//
// Replaces mxq_quant/cuda_kernel/csrc/quantization/gemm_cuda_gen.cu:28-478 (Ampere-class
// mma.sync.m16n8k16 + ldmatrix, M tiles of 16, fp16 split-K partials summed by a second torch op;
// the reference never compiles or exports it).  Interface contract (:418-423):
//   in_feats [M, IC] fp16;  kernel [IC, OC/8] int32;  scales [IC/G, OC] fp16;  zeros [IC/G, OC/8] int32
//   W[k][n] = scales[k/G][n] * (q[k][n] - z[k/G][n]), rounded to fp16 (:131-138: sub.f16x2, fma.rn.f16x2)
// where output channel 8j + c takes nibble {0, 4, 1, 5, 2, 6, 3, 7}[c] of word j -- the order in which
// dequantize_s4_to_fp16x2 (dequantize.cuh:15-77) leaves its eight halves.
// Two kernels: (1) dequantize + transpose to a K-contiguous fp16 W^T [OC, IC] in the workspace
// (64 x 64 tiles through shared memory, 128-byte rows both ways); (2) the tcgen05 / TMA / TMEM
// pipeline of gemm_tcgen05.cu with that dense operand (y = x @ W^T, fp32 accumulation, no split-K
// rounding).  The intermediate costs 4 B per weight of L2 / HBM traffic next to a 2*M FLOP-per-weight
// GEMM: 13 % of the time at M = 2048 on 4096^2 (measured, profiles/); a fused 4-bit producer for the
// pair kernel is the next step.
#include "common.cuh"

namespace mxq {

// tile: 64 input channels (k) x 64 output channels (n).  256 threads.
__global__ void __launch_bounds__(256) awq_dequant_t_kernel(const uint32_t* __restrict__ kernel,
                                                            const __half* __restrict__ scales,
                                                            const uint32_t* __restrict__ zeros,
                                                            __half* __restrict__ Wt, int IC, int OC, int G) {
  __shared__ __half tile[64][64 + 8];              // [n][k], padded: conflict-free transposed stores
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  // 64 k x 8 words: thread -> (k = tid / 4 .. , word pair)
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int k = i >> 3, w = i & 7;
    const int kk = k0 + k, g = kk / G;
    const uint32_t q = kernel[(size_t)kk * (OC / 8) + (n0 / 8) + w];
    const uint32_t z = zeros[(size_t)g * (OC / 8) + (n0 / 8) + w];
    const __half* sp = scales + (size_t)g * OC + n0 + w * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // channel c of the word holds nibble (c & 1) * 4 + (c >> 1): {0, 4, 1, 5, 2, 6, 3, 7}
      const int nib = (j & 1) * 4 + (j >> 1);
      const float qv = (float)((q >> (4 * nib)) & 0xF), zv = (float)((z >> (4 * nib)) & 0xF);
      // (q - z) is exact in fp16; one fp16 rounding of the product, as the reference's fma.rn.f16x2
      tile[w * 8 + j][k] = __float2half_rn(__fmul_rn(__fsub_rn(qv, zv), __half2float(sp[j])));
    }
  }
  __syncthreads();
  // 64 n rows x 128 bytes
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int n = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(Wt + (size_t)(n0 + n) * IC + k0 + c * 8) = *reinterpret_cast<const uint4*>(&tile[n][c * 8]);
  }
}

}  // namespace mxq

using namespace mxq;

extern "C" size_t mxq_awq_gemm_workspace_bytes(int64_t IC, int64_t OC) {
  return IC > 0 && OC > 0 ? (size_t)IC * (size_t)OC * 2 : 0;
}

extern "C" int mxq_awq_gemm(const void* x, const int32_t* kernel, const void* scales, const int32_t* zeros, void* y,
                            int64_t M, int64_t IC, int64_t OC, int group_size, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (M < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(y);
  MXQ_CHECK_PTR(kernel);
  MXQ_CHECK_PTR(workspace);
  if (!scales || !zeros) return MXQ_E_NULL;
  // gemm_cuda_gen.cu:447-454: OC % 64, group_size % 32, OC % group_size
  if (group_size <= 0 || group_size % 32 || IC % group_size || IC % 64 || IC == 0 || OC % 64 || OC % group_size) return MXQ_E_SHAPE;
  if (workspace_bytes < mxq_awq_gemm_workspace_bytes(IC, OC)) return MXQ_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  awq_dequant_t_kernel<<<dim3((unsigned)(IC / 64), (unsigned)(OC / 64)), 256, 0, st>>>(
      (const uint32_t*)kernel, (const __half*)scales, (const uint32_t*)zeros, (__half*)workspace, (int)IC, (int)OC, group_size);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return mxq_gemm_dense(x, workspace, y, M, IC, OC, stream);
}
