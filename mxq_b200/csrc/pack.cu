// Unpacker for the mixed 2/4-bit layout consumed by the reference GEMV
// (mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-208; layout in include/mxq_b200.h):
// decode formula of :131-136,152-153,178-179,191-192, bit positions of :101-103,109-110,144-199.
// (The packer lives in ptq.cu, fused with fasterquant.)
#include "common.cuh"

namespace mxq {

// decode: one thread per (row, block); writes 64 values
template <typename OutT>
__global__ void __launch_bounds__(256) unpack_kernel(mxq_packed_t in, OutT* __restrict__ out,
                                                     int OC, int IC) {
  const int nblk = IC / 64;
  const int nchunk = (nblk + 63) / 64;
  const int64_t total = (int64_t)OC * nblk;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int oc = (int)(i / nblk), blk = (int)(i % nblk);
  const uint4 w = *reinterpret_cast<const uint4*>(in.weight + (size_t)oc * nblk * 4 + (size_t)blk * 4);
  const uint32_t wl = (uint32_t)in.weight_last[(size_t)oc * nblk + blk];
  const int chunk = blk >> 6, bp = blk & 63;
  const int word = chunk * 32 + (bp & 31), p = bp >> 5;
  const uint32_t zs = (uint32_t)in.zeros_and_scales[(size_t)oc * 32 * nchunk + word] >> (16 * p);
  const uint32_t z2w = (uint32_t)in.zeros_2nd[(size_t)(oc >> 2) * 32 * nchunk + word] >> (8 * p);
  const __half* s2 = reinterpret_cast<const __half*>(in.scales_2nd) + (size_t)(oc >> 2) * nblk * 3 + (size_t)blk * 3;
  const float s4 = __half2float(reinterpret_cast<const __half*>(in.scales_4b)[oc]);
  const float z4 = (float)(((uint32_t)in.zeros_4b[oc >> 3] >> (4 * (oc & 7))) & 0xF);
  OutT* o = out + (size_t)oc * IC + (size_t)blk * 64;
  const uint32_t ww[3] = {w.x, w.y, w.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float z1 = (float)((zs >> (2 * k)) & 3);
    const float c = (float)((zs >> (8 + 2 * k)) & 3);
    const float z2 = (float)((z2w >> (2 * k)) & 3);
    const float sc = __fmul_rn(__half2float(s2[k]), __fsub_rn(c, z2));
#pragma unroll
    for (int j = 0; j < 16; ++j)
      o[k * 16 + j] = (OutT)__fmul_rn(sc, __fsub_rn((float)((ww[k] >> (2 * j)) & 3), z1));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o[48 + j] = (OutT)__fmul_rn(s4, __fsub_rn((float)((w.w >> (4 * j)) & 0xF), z4));
    o[56 + j] = (OutT)__fmul_rn(s4, __fsub_rn((float)((wl >> (4 * j)) & 0xF), z4));
  }
}

}  // namespace mxq

using namespace mxq;

static int check_packed(const mxq_packed_t& p) {
  MXQ_CHECK_PTR(p.weight);
  MXQ_CHECK_PTR(p.weight_last);
  MXQ_CHECK_PTR(p.zeros_and_scales);
  MXQ_CHECK_PTR(p.zeros_2nd);
  if (!p.scales_2nd || !p.scales_4b || !p.zeros_4b) return MXQ_E_NULL;
  return MXQ_OK;
}

extern "C" int mxq_unpack(mxq_packed_t in, int64_t OC, int64_t IC, void* out, int out_dtype,
                          void* stream) {
  if (OC < 0 || IC < 0) return MXQ_E_SHAPE;
  if (OC == 0 || IC == 0) return MXQ_OK;
  int rc = check_packed(in);
  if (rc) return rc;
  MXQ_CHECK_PTR(out);
  if (OC % 8 || IC % 64) return MXQ_E_SHAPE;
  const int64_t total = OC * (IC / 64);
  const unsigned grid = (unsigned)ceil_div(total, 256);
  cudaStream_t st = as_stream(stream);
  if (out_dtype == MXQ_F32)
    unpack_kernel<float><<<grid, 256, 0, st>>>(in, (float*)out, (int)OC, (int)IC);
  else if (out_dtype == MXQ_F16)
    unpack_kernel<__half><<<grid, 256, 0, st>>>(in, (__half*)out, (int)OC, (int)IC);
  else
    return MXQ_E_DTYPE;
  MXQ_LAUNCH_RESULT();
}

// out[m, 16 g .. 16 g + 15] = in[m, 16 perm[g] ..]: the importance-driven column order applied to
// fp16 rows -- the weights before packing, the activations before the prefill GEMM (the decode GEMV
// applies it while it stages x).  One 32-byte group per thread pair.
namespace mxq {
__global__ void __launch_bounds__(256) gather_groups_kernel(const uint4* __restrict__ in, const int32_t* __restrict__ perm,
                                                            uint4* __restrict__ out, int64_t rows, int ngroups) {
  const int64_t total = rows * ngroups * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (ngroups * 2);
    const int c = (int)(i - r * ngroups * 2);
    const int g = c >> 1;
    out[i] = in[r * ngroups * 2 + (int64_t)__ldg(perm + g) * 2 + (c & 1)];
  }
}
}  // namespace mxq

extern "C" int mxq_gather_groups(const void* in, const int32_t* group_perm, void* out, int64_t rows, int64_t cols,
                                 void* stream) {
  if (rows < 0 || cols < 0 || cols % 16) return MXQ_E_SHAPE;
  if (rows == 0 || cols == 0) return MXQ_OK;
  MXQ_CHECK_PTR(in);
  MXQ_CHECK_PTR(out);
  if (!group_perm) return MXQ_E_NULL;
  if (in == out) return MXQ_E_UNSUPPORTED;
  const int64_t total = rows * (cols / 16) * 2;
  int64_t grid = mxq::ceil_div(total, 256);
  if (grid > mxq::kNumSMs * 16) grid = mxq::kNumSMs * 16;
  mxq::gather_groups_kernel<<<(unsigned)grid, 256, 0, mxq::as_stream(stream)>>>((const uint4*)in, group_perm, (uint4*)out, rows,
                                                                               (int)(cols / 16));
  MXQ_LAUNCH_RESULT();
}
