// Packer / unpacker for the mixed 2/4-bit layout consumed by the reference GEMV
// (mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-208; layout in include/mxq_b200.h).
// The reference has no producer for this layout: the ENCODE policy is ours (oracle
// pack_mxq); the DECODE formula is the reference's and is what mxq_unpack / mxq_gemv / mxq_gemm
// implement.  Codes are bit-exact against oracle/mxq_oracle.py:pack_mxq.
#include "common.cuh"

namespace mxq {

__device__ __forceinline__ float clampf(float v, float lo, float hi) {
  return fminf(fmaxf(v, lo), hi);
}

__device__ __forceinline__ void load_w8(const __half* W, const uint8_t* dead, size_t off, int col,
                                        float* f) {
  const uint4 ch = *reinterpret_cast<const uint4*>(W + off + col);
  DT<__half>::unpack(ch, f);
  const uint2 dm = *reinterpret_cast<const uint2*>(dead + col);
  if (dm.x | dm.y) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t w = e < 4 ? dm.x : dm.y;
      if ((w >> (8 * (e & 3))) & 0xFF) f[e] = 0.f;
    }
  }
}

__global__ void pack_dead_mask_kernel(const float* __restrict__ colstat,
                                      uint8_t* __restrict__ dead, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) dead[c] = colstat ? (colstat[c] == 0.f) : 0;
}

// per-row 4-bit pool: min/max (zero included) -> s4 (fp16), z4; zeros_4b nibble-packed 8 rows/word
__global__ void __launch_bounds__(256) pack_pool_kernel(const __half* __restrict__ W,
                                                        const uint8_t* __restrict__ dead,
                                                        float2* __restrict__ pool, // (s4f, z4)
                                                        __half* __restrict__ scales_4b,
                                                        uint32_t* __restrict__ zeros_4b, int OC,
                                                        int IC) {
  __shared__ uint32_t z4s[8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + w;  // OC % 8 == 0
  const int nblk = IC / 64;
  float mn = 0.f, mx = 0.f;
  const size_t roff = (size_t)row * IC;
  for (int m = lane; m < nblk * 2; m += 32) {
    float f[8];
    load_w8(W, dead, roff, (m >> 1) * 64 + 48 + (m & 1) * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  __half s4h = __float2half_rn(__fdiv_rn(__fsub_rn(mx, mn), 15.f));
  if (__half2float(s4h) == 0.f) s4h = __float2half_rn(1.f);
  const float s4f = __half2float(s4h);
  const float z4 = clampf(rintf(__fdiv_rn(-mn, s4f)), 0.f, 15.f);
  if (lane == 0) {
    pool[row] = make_float2(s4f, z4);
    scales_4b[row] = s4h;
    z4s[w] = (uint32_t)z4;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) word |= z4s[i] << (4 * i);
    zeros_4b[blockIdx.x] = word;
  }
}

// warp unit = 16 rows x one 64-column block; lane = (row r = lane>>1, half h = lane&1)
__global__ void __launch_bounds__(256) pack_tile_kernel(const __half* __restrict__ W,
                                                        const uint8_t* __restrict__ dead,
                                                        const float2* __restrict__ pool,
                                                        mxq_packed_t out, int OC, int IC) {
  const int lane = threadIdx.x & 31;
  const int r = lane >> 1, h = lane & 1;
  const int nblk = IC / 64;
  const int nchunk = (nblk + 63) / 64;
  const int64_t units = (int64_t)(OC / 16) * nblk;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t wstride = (int64_t)gridDim.x * 8;
  uint16_t* zs16 = reinterpret_cast<uint16_t*>(out.zeros_and_scales);
  uint8_t* z2b = reinterpret_cast<uint8_t*>(out.zeros_2nd);
  __half* s2o = reinterpret_cast<__half*>(out.scales_2nd);
  for (int64_t u = warp0; u < units; u += wstride) {
    const int tile = (int)(u / nblk), blk = (int)(u % nblk);
    const int row = tile * 16 + r;
    const size_t roff = (size_t)row * IC;
    float f[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) load_w8(W, dead, roff, blk * 64 + k * 16 + h * 8, f[k]);
    uint32_t words[4];
    uint32_t zbyte = 0, cbyte = 0;
    __half s2h[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float lo = 0.f, hi = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { lo = fminf(lo, f[k][e]); hi = fmaxf(hi, f[k][e]); }
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 1));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
      float s = __fdiv_rn(__fsub_rn(hi, lo), 3.f);
      if (s == 0.f) s = 1.f;
      float smax = s;
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, 2));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, 4));
      s2h[k] = __float2half_rn(__fdiv_rn(smax, 3.f));
      const float s2f = __half2float(s2h[k]);
      const float c = clampf(rintf(__fdiv_rn(s, s2f)), 1.f, 3.f);
      const float S = __fmul_rn(s2f, c);
      const float z1 = clampf(rintf(__fdiv_rn(-lo, S)), 0.f, 3.f);
      const float rS = __frcp_rn(S);
      uint32_t bits = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float q = clampf(__fadd_rn(rintf(div_rn_by(f[k][e], S, rS)), z1), 0.f, 3.f);
        bits |= (uint32_t)q << (2 * e);
      }
      const uint32_t other = __shfl_xor_sync(0xffffffffu, bits, 1);
      words[k] = h ? (other | (bits << 16)) : (bits | (other << 16));
      zbyte |= (uint32_t)z1 << (2 * k);
      cbyte |= (uint32_t)c << (2 * k);
    }
    {
      const float2 pz = pool[row];
      const float rS = __frcp_rn(pz.x);
      uint32_t bits = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float q = clampf(__fadd_rn(rintf(div_rn_by(f[3][e], pz.x, rS)), pz.y), 0.f, 15.f);
        bits |= (uint32_t)q << (4 * e);
      }
      words[3] = bits;  // h == 0: columns 48..55 -> weight word 3 ; h == 1: 56..63 -> weight_last
    }
    if (h == 0) {
      *reinterpret_cast<uint4*>(out.weight + (size_t)row * nblk * 4 + (size_t)blk * 4) =
          make_uint4(words[0], words[1], words[2], words[3]);
      const int chunk = blk >> 6, bp = blk & 63;
      const int word = chunk * 32 + (bp & 31), p = bp >> 5;
      zs16[((size_t)row * 32 * nchunk + word) * 2 + p] = (uint16_t)(zbyte | (cbyte << 8));
      if ((r & 3) == 0) {
        z2b[((size_t)(row >> 2) * 32 * nchunk + word) * 4 + p] = 0;  // z2 == 0 policy
        __half* d = s2o + (size_t)(row >> 2) * nblk * 3 + (size_t)blk * 3;
        d[0] = s2h[0]; d[1] = s2h[1]; d[2] = s2h[2];
      }
    } else {
      out.weight_last[(size_t)row * nblk + blk] = (int32_t)words[3];
    }
  }
}

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

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

static int check_packed(const mxq_packed_t& p) {
  MXQ_CHECK_PTR(p.weight);
  MXQ_CHECK_PTR(p.weight_last);
  MXQ_CHECK_PTR(p.zeros_and_scales);
  MXQ_CHECK_PTR(p.zeros_2nd);
  if (!p.scales_2nd || !p.scales_4b || !p.zeros_4b) return MXQ_E_NULL;
  return MXQ_OK;
}

extern "C" size_t mxq_pack_workspace_bytes(int64_t OC, int64_t IC) {
  if (OC < 0 || IC < 0) return 0;
  return align16((size_t)IC) + align16((size_t)OC * sizeof(float2)) + 16;
}

extern "C" int mxq_pack(const void* W, const float* colstat, int64_t OC, int64_t IC,
                        mxq_packed_t out, void* workspace, size_t workspace_bytes, void* stream) {
  if (OC < 0 || IC < 0) return MXQ_E_SHAPE;
  if (OC == 0 || IC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(W);
  MXQ_CHECK_PTR(workspace);
  int rc = check_packed(out);
  if (rc) return rc;
  if (OC % 16 || IC % 64 || OC > INT32_MAX || IC > (1 << 24)) return MXQ_E_SHAPE;
  if (workspace_bytes < mxq_pack_workspace_bytes(OC, IC)) return MXQ_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  uint8_t* dead = (uint8_t*)workspace;
  float2* pool = (float2*)((uint8_t*)workspace + align16((size_t)IC));
  const int nblk = (int)(IC / 64);
  const int nchunk = (nblk + 63) / 64;
  if (nblk % 64) {  // padding half-words of the last metadata chunk
    cudaMemsetAsync(out.zeros_and_scales, 0, (size_t)OC * 32 * nchunk * 4, st);
    cudaMemsetAsync(out.zeros_2nd, 0, (size_t)(OC / 4) * 32 * nchunk * 4, st);
  }
  pack_dead_mask_kernel<<<(unsigned)ceil_div(IC, 256), 256, 0, st>>>(colstat, dead, (int)IC);
  pack_pool_kernel<<<(unsigned)(OC / 8), 256, 0, st>>>((const __half*)W, dead, pool,
                                                       (__half*)out.scales_4b,
                                                       (uint32_t*)out.zeros_4b, (int)OC, (int)IC);
  const int64_t units = (OC / 16) * (int64_t)nblk;
  int64_t grid = ceil_div(units, 8);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  pack_tile_kernel<<<(unsigned)grid, 256, 0, st>>>((const __half*)W, dead, pool, out, (int)OC, (int)IC);
  MXQ_LAUNCH_RESULT();
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
