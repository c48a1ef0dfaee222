// Decode GEMV for the packed mixed 2/4-bit layout (sm_100a), plus the AWQ uniform-4-bit GEMV.
//
// Replaces mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-273 (IC hard-wired to
// 4096, one scalar cvt+FMA chain per weight, legacy stream) and gemv_cuda.cu:45-242,346-399.
// This kernel is weight-stream (HBM) bound: 0.3756 B per weight.
//   * warp = one second-order group (4 output rows) x a K-slice; lane = one 64-column block, so
//     a lane's 128-bit weight load, its 4-bit tail word and its metadata are issued up front for
//     all 4 rows (>= 100 B in flight per lane) and the activations of the block are loaded once
//     for the 4 rows;
//   * dequant two codes per SHF+LOP3 into fp16 {bias+q} pairs (code in the top mantissa bits) that
//     FHFMA (fma.rn.f32.f16, sm_100) multiplies straight into an fp32 accumulator -- no
//     int->float conversions, no per-weight subtraction; bias and zero-point are removed once per
//     (row, group) with the group's activation sum;
//   * the scale s2*(c - z2) is applied once per (row, group) on the fp32 group sum;
//   * K-slices of a row group are reduced through shared memory in a fixed order.
// Any IC % 64 == 0 (metadata tiled in 4096-column chunks; identical to the reference at 4096);
// the reference's activation-offset and batch-stride bugs (gemv_mxq_cuda.cu:50,119) are not
// reproduced.
#include <cstdlib>

#include "common.cuh"

namespace mxq {

// a2 = {w_lo, w_hi} pairs with x halves selected independently
__device__ __forceinline__ float fhfma_sel(uint32_t a2, int ah, uint32_t b2, int bh, float acc) {
  unsigned short a0, a1, b0, b1;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(a0), "=h"(a1) : "r"(a2));
  asm("mov.b32 {%0,%1}, %2;" : "=h"(b0), "=h"(b1) : "r"(b2));
  const unsigned short a = ah ? a1 : a0, b = bh ? b1 : b0;
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(b));
  return acc;
}
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(orv));
  return d;
}
__device__ __forceinline__ uint32_t hsub2_u32(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

constexpr int kGemvWarps = 8;

// Dequant without subtraction: a code is moved to the TOP mantissa bits of an fp16 whose exponent
// makes one mantissa step equal to 1, so the register already holds the number bias + q
// (bias 4 for 2-bit codes, 16 for 4-bit codes) and FHFMA consumes it directly.  The bias and the
// zero-point are removed once per (row, group) on the fp32 group sum:
//     sum_j (q_j - z) x_j = sum_j (bias + q_j) x_j - (bias + z) * sum_j x_j
// (fp32 accumulation loses only log2(bias) bits to the cancellation).
//
// x16: 16 activations of one group as 8 half2 words (cols 0..15 in order).
// 2-bit word: code j at bits [2j+1:2j]; shifting by 8-2j puts code j at bits [9:8] of the low
// half and code j+8 at bits [9:8] of the high half.
__device__ __forceinline__ float dot_group_2b(uint32_t w, const uint32_t* x16) {
  float p = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t t = (8 - 2 * j) >= 0 ? (w << (8 - 2 * j)) : (w >> (2 * j - 8));
    const uint32_t h = lop3_and_or(t, 0x03000300u, 0x44004400u);   // {4 + q_j, 4 + q_{j+8}}
    p = fhfma_sel(h, 0, x16[j >> 1], j & 1, p);
    p = fhfma_sel(h, 1, x16[(j + 8) >> 1], (j + 8) & 1, p);
  }
  return p;
}
// 4-bit word: nibble j at bits [4j+3:4j]; shifting by 6-4j puts nibble j at bits [9:6] of the low
// half and nibble j+4 at bits [9:6] of the high half: {16 + n_j, 16 + n_{j+4}}.
__device__ __forceinline__ float dot_word_4b(uint32_t w, const uint32_t* x8, float p) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t t = (6 - 4 * j) >= 0 ? (w << (6 - 4 * j)) : (w >> (4 * j - 6));
    const uint32_t h = lop3_and_or(t, 0x03C003C0u, 0x4C004C00u);
    p = fhfma_sel(h, 0, x8[j >> 1], j & 1, p);
    p = fhfma_sel(h, 1, x8[(j + 4) >> 1], (j + 4) & 1, p);
  }
  return p;
}
// sum of the 2n halves held in n words, fp32
template <int N>
__device__ __forceinline__ float sum_halves(const uint32_t* x) {
  float p = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    p = fhfma_sel(0x3C003C00u, 0, x[i], 0, p);
    p = fhfma_sel(0x3C003C00u, 0, x[i], 1, p);
  }
  return p;
}

template <int NB>
__global__ void __launch_bounds__(kGemvWarps * 32) gemv_mxq_kernel(
    const __half* __restrict__ x, mxq_packed_t w, __half* __restrict__ y, int B, int IC, int OC,
    int KS, int ks_shift) {
  __shared__ float red[kGemvWarps][4][NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RG = kGemvWarps >> ks_shift;
  const int rg = warp >> ks_shift, ks = warp & (KS - 1);
  const int nblk = IC >> 6;
  const int nchunk = (nblk + 63) >> 6;
  const int grp = blockIdx.x * RG + rg;          // second-order group (4 rows)
  const int b0 = blockIdx.y * NB;
  const bool live = grp * 4 < OC;
  float acc[4][NB];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;

  if (live) {
    const int oc0 = grp * 4;
    const __half* s4p = reinterpret_cast<const __half*>(w.scales_4b) + oc0;
    const uint32_t z4w = (uint32_t)w.zeros_4b[oc0 >> 3] >> (4 * (oc0 & 7));
    float s4[4], z4b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      s4[r] = __half2float(s4p[r]);
      z4b[r] = __uint_as_float(0x41800000u + (((z4w >> (4 * r)) & 0xF) << 19));   // 16 + z4
    }
    const uint16_t* zs16 = reinterpret_cast<const uint16_t*>(w.zeros_and_scales);
    const uint8_t* z2b = reinterpret_cast<const uint8_t*>(w.zeros_2nd);
    const __half* s2p = reinterpret_cast<const __half*>(w.scales_2nd) + (size_t)grp * nblk * 3;

    for (int blk = ks * 32 + lane; blk < nblk; blk += KS * 32) {
      uint4 wq[4];
      uint32_t wl[4], zs[4];
      const int chunk = blk >> 6, bp = blk & 63;
      const int word = chunk * 32 + (bp & 31), p = bp >> 5;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const size_t row = (size_t)(oc0 + r);
        wq[r] = ld_stream(w.weight + row * nblk * 4 + (size_t)blk * 4);
        wl[r] = (uint32_t)__ldg(w.weight_last + row * nblk + blk);
        zs[r] = __ldg(zs16 + (row * 32 * nchunk + word) * 2 + p);
      }
      const uint32_t z2 = __ldg(z2b + ((size_t)grp * 32 * nchunk + word) * 4 + p);
      float s2[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) s2[k] = __half2float(__ldg(s2p + (size_t)blk * 3 + k));

#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t xv[NB][8];
        float xs[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int bb = min(b0 + b, B - 1);
          const uint4* xp = reinterpret_cast<const uint4*>(x + (size_t)bb * IC + (size_t)blk * 64 + k * 16);
          const uint4 v0 = __ldg(xp), v1 = __ldg(xp + 1);
          xv[b][0] = v0.x; xv[b][1] = v0.y; xv[b][2] = v0.z; xv[b][3] = v0.w;
          xv[b][4] = v1.x; xv[b][5] = v1.y; xv[b][6] = v1.z; xv[b][7] = v1.w;
          xs[b] = sum_halves<8>(xv[b]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (k < 3) {
            const uint32_t wk = k == 0 ? wq[r].x : (k == 1 ? wq[r].y : wq[r].z);
            const float zb = __uint_as_float(0x40800000u + (((zs[r] >> (2 * k)) & 3) << 21));   // 4 + z1
            const int d = 3 + (int)((zs[r] >> (8 + 2 * k)) & 3) - (int)((z2 >> (2 * k)) & 3);   // 3 + c - z2
            const float scale = s2[k] * (__uint_as_float(0x41000000u + ((uint32_t)d << 20)) - 11.0f);  // :136
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const float pz = fmaf(-zb, xs[b], dot_group_2b(wk, xv[b]));
              acc[r][b] = fmaf(scale, pz, acc[r][b]);
            }
          } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              float pz = dot_word_4b(wq[r].w, xv[b], 0.f);
              pz = dot_word_4b(wl[r], xv[b] + 4, pz);
              pz = fmaf(-z4b[r], xs[b], pz);
              acc[r][b] = fmaf(s4[r], pz, acc[r][b]);               // :179,192
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float v = warp_sum(acc[r][b]);
      if (lane == 0) red[warp][r][b] = v;
    }
  __syncthreads();
  // one thread per (row group, row, batch)
  const int t = threadIdx.x;
  if (t < RG * 4 * NB) {
    const int g = t / (4 * NB), r = (t / NB) & 3, b = t % NB;
    const int oc = (blockIdx.x * RG + g) * 4 + r;
    if (oc < OC && b0 + b < B) {
      float s = 0.f;
      for (int k = 0; k < KS; ++k) s += red[(g << ks_shift) + k][r][b];
      y[(size_t)(b0 + b) * OC + oc] = __float2half_rn(s);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// AWQ uniform 4-bit GEMV (gemv_cuda.cu:45-242): kernel[OC, IC/8] (nibble j of word i = column
// 8i+j), zeros[OC, zw] 4-bit per group (8 groups per word), scales fp16 [OC, zw*8]
// w = scale * (q - zero).   One warp per output row, lanes stride over 16-byte weight chunks.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) awq_gemv_kernel(const __half* __restrict__ x,
                                                       const uint32_t* __restrict__ kernel,
                                                       const __half* __restrict__ scales,
                                                       const uint32_t* __restrict__ zeros,
                                                       __half* __restrict__ y, int IC, int OC,
                                                       int G, int zw) {
  const int lane = threadIdx.x & 31;
  const int oc = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (oc >= OC) return;
  const int ww = IC / 8;
  float acc = 0.f;
  for (int c = lane; c < ww / 4; c += 32) {  // 4 words = 32 columns per chunk
    const uint4 wv = ld_stream(kernel + (size_t)oc * ww + (size_t)c * 4);
    const int g = (c * 32) / G;
    const uint32_t z = (zeros[(size_t)oc * zw + (g >> 3)] >> (4 * (g & 7))) & 0xF;
    const float sc = __half2float(scales[(size_t)oc * zw * 8 + g]);
    const float zb = __uint_as_float(0x41800000u + (z << 19));   // 16 + zero
    const uint4* xp = reinterpret_cast<const uint4*>(x + (size_t)b * IC + (size_t)c * 32);
    const uint32_t wd[4] = {wv.x, wv.y, wv.z, wv.w};
    float p = 0.f, xs = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 xv = __ldg(xp + i);
      const uint32_t xr[4] = {xv.x, xv.y, xv.z, xv.w};
      p = dot_word_4b(wd[i], xr, p);
      xs += sum_halves<4>(xr);
    }
    acc = fmaf(sc, fmaf(-zb, xs, p), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) y[(size_t)b * OC + oc] = __float2half_rn(acc);
}

}  // namespace mxq

using namespace mxq;

extern "C" int mxq_gemv(const void* x, mxq_packed_t w, void* y, int64_t B, int64_t IC, int64_t OC,
                        void* stream) {
  if (B < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (B == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(y);
  MXQ_CHECK_PTR(w.weight);
  if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b ||
      !w.zeros_4b)
    return MXQ_E_NULL;
  if (IC % 64 || OC % 8 || IC == 0 || IC > (1 << 24) || OC > INT32_MAX || B > 65535 * 4)
    return MXQ_E_SHAPE;
  const int nblk = (int)(IC / 64);
  int KS = 1, ks_shift = 0;
  while (KS < kGemvWarps && nblk > 32 * KS) { KS <<= 1; ++ks_shift; }
  if (const char* e = getenv("MXQ_GEMV_KS")) {   // tuning knob (profiles/sweep_gemv.py)
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) { KS = v; ks_shift = v == 1 ? 0 : v == 2 ? 1 : v == 4 ? 2 : 3; }
  }
  const int RG = kGemvWarps / KS;
  const unsigned gx = (unsigned)ceil_div(OC / 4, RG);
  cudaStream_t st = as_stream(stream);
  const __half* xh = (const __half*)x;
  __half* yh = (__half*)y;
  if (B == 1) {
    gemv_mxq_kernel<1><<<dim3(gx, 1), kGemvWarps * 32, 0, st>>>(xh, w, yh, (int)B, (int)IC, (int)OC, KS, ks_shift);
  } else if (B == 2) {
    gemv_mxq_kernel<2><<<dim3(gx, 1), kGemvWarps * 32, 0, st>>>(xh, w, yh, (int)B, (int)IC, (int)OC, KS, ks_shift);
  } else {
    gemv_mxq_kernel<4><<<dim3(gx, (unsigned)ceil_div(B, 4)), kGemvWarps * 32, 0, st>>>(xh, w, yh, (int)B, (int)IC, (int)OC, KS, ks_shift);
  }
  MXQ_LAUNCH_RESULT();
}

extern "C" int mxq_awq_gemv(const void* x, const int32_t* kernel, const void* scales,
                            const int32_t* zeros, void* y, int64_t B, int64_t IC, int64_t OC,
                            int group_size, void* stream) {
  if (B < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (B == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(y);
  MXQ_CHECK_PTR(kernel);
  if (!scales || !zeros) return MXQ_E_NULL;
  if (group_size != 32 && group_size != 64 && group_size != 128) return MXQ_E_UNSUPPORTED;
  if (IC % 32 || IC % group_size || IC == 0 || B > 65535) return MXQ_E_SHAPE;
  // zeros row width (words) per variant: g128 ceil(ng/8) (gemv_cuda.cu:200), g64 rounded up to 2
  // (:129), g32 rounded up to 4 (:56); scales row width = zeros_w * 8.  Group g = col / G uses
  // nibble g%8 of zeros word g/8 and scale g -- exactly the reference's g64/g128 indexing (its
  // g32 variant indexes scales/zeros inconsistently, :70-71, and is not reproduced).
  const int rnd = group_size == 128 ? 1 : (group_size == 64 ? 2 : 4);
  const int zw = (int)(ceil_div(ceil_div(IC / group_size, 8), rnd) * rnd);
  awq_gemv_kernel<<<dim3((unsigned)ceil_div(OC, 8), (unsigned)B), 256, 0, as_stream(stream)>>>(
      (const __half*)x, (const uint32_t*)kernel, (const __half*)scales, (const uint32_t*)zeros,
      (__half*)y, (int)IC, (int)OC, group_size, zw);
  MXQ_LAUNCH_RESULT();
}
