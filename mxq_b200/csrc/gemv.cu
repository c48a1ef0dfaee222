// Decode GEMV for the packed mixed 2/4-bit layout (sm_100a), plus the AWQ uniform-4-bit GEMV.
//
// Replaces mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-273 (IC hard-wired to
// 4096, one scalar cvt+FMA chain per weight, legacy stream) and gemv_cuda.cu:45-242,346-399.
// This kernel is weight-stream (HBM) bound: 0.3756 B per weight.  Design (numbers: profiles/):
//   * persistent grid, one CTA per SM owning a contiguous range of output rows, so each packed
//     tensor's share is ONE contiguous byte range: a stage of the shared-memory ring is fetched
//     with four 1-D TMA bulk copies (cp.async.bulk + mbarrier), issued before the activations
//     are needed;
//   * programmatic dependent launch: back-to-back GEMVs overlap the next kernel's launch and
//     weight fetch with the current kernel's arithmetic (griddepcontrol.wait before x is read);
//   * warp = one second-order group (4 output rows) x K slices of 2048 columns; lane = one
//     64-column block (conflict-free 128-bit shared loads);
//   * the activations are converted ONCE per CTA to block floating point (int16 mantissas per
//     16-column group, split into signed high / unsigned low bytes) and the 2/4-bit codes meet
//     them in dp4a -- IDP4A co-issues with the LOP3/SHF unpacking, fp16 FMAs do not
//     (profiles/probes/fhfma_probe.cu); zero-point and scale s2*(c - z2) are applied once per
//     (row, group) on the exact integer sum;
//   * K slices of a row group are reduced through shared memory in a fixed order.
// Any IC % 64 == 0 (metadata tiled in 4096-column chunks; identical to the reference at 4096);
// the reference's activation-offset and batch-stride bugs (gemv_mxq_cuda.cu:50,119) are not
// reproduced.
#include <cstdio>
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

constexpr int kXBlkBytes = 144;   // 64 fp16 activations of a block + 16 B pad (conflict-free LDS.128)

// ---------------------------------------------------------------------------------------------
// Dequant without shifts or subtractions.  A 2-bit code sitting at mantissa bits [p+1:p] of an
// fp16 whose exponent field makes bit p worth 1.0 IS the number bias_p + q (bias 1024 at p=0,
// 256 at p=2, 64 at p=4, 16 at p=6, 4 at p=8), so one LOP3 (mask | exponent) on the packed word
// yields the fp16 pair {bias + q_j, bias + q_{j+8}} for j = 0..4 with no shift at all, and one
// SHF (>> 6) exposes j = 5..7 the same way: 9 instructions per 16 codes.  FHFMA
// (fma.rn.f32.f16, sm_100) multiplies the pair straight into an fp32 accumulator (products are
// exact, 11 x 11 bits).  The bias and the zero-point leave on the fp32 side, once per (row, group):
//     sum_j (q_j - z) x_j = sum_j (bias_j + q_j) x_j - [ sum_j bias_j x_j + z * sum_j x_j ]
// The bracket's two sums are row independent: the CTA computes them once per group while it stages
// the activations in shared memory (tables tabA/tabB below) and they seed the accumulator.
// ---------------------------------------------------------------------------------------------
// bias of code j (and j+8) of a 2-bit word, of nibble j (and j+4) of a 4-bit word
__host__ __device__ constexpr float bias2(int j) {
  j &= 7;
  return j == 0 ? 1024.f : j == 1 ? 256.f : j == 2 ? 64.f : j == 3 ? 16.f : j == 4 ? 4.f
       : j == 5 ? 64.f : j == 6 ? 16.f : 4.f;
}
__host__ __device__ constexpr float bias4(int j) {
  j &= 3;
  return j == 0 ? 1024.f : j == 1 ? 64.f : j == 2 ? 256.f : 16.f;
}

__host__ __device__ constexpr uint32_t bias_half(float b) {   // fp16 bits of a power of two
  return b == 1024.f ? 0x6400u : b == 256.f ? 0x5C00u : b == 64.f ? 0x5400u : b == 16.f ? 0x4C00u
       : 0x4400u;
}

// h[i] = {bias2(i) + q_i, bias2(i) + q_{i+8}}, i = 0..7
__device__ __forceinline__ void dequant_word_2b(uint32_t w, uint32_t* h) {
  h[0] = lop3_and_or(w, 0x00030003u, 0x64006400u);
  h[1] = lop3_and_or(w, 0x000C000Cu, 0x5C005C00u);
  h[2] = lop3_and_or(w, 0x00300030u, 0x54005400u);
  h[3] = lop3_and_or(w, 0x00C000C0u, 0x4C004C00u);
  h[4] = lop3_and_or(w, 0x03000300u, 0x44004400u);
  const uint32_t t = w >> 6;
  h[5] = lop3_and_or(t, 0x00300030u, 0x54005400u);
  h[6] = lop3_and_or(t, 0x00C000C0u, 0x4C004C00u);
  h[7] = lop3_and_or(t, 0x03000300u, 0x44004400u);
}
// h[i] = {bias4(i) + n_i, bias4(i) + n_{i+4}}, i = 0..3
__device__ __forceinline__ void dequant_word_4b(uint32_t w, uint32_t* h) {
  h[0] = lop3_and_or(w, 0x000F000Fu, 0x64006400u);
  h[1] = lop3_and_or(w, 0x00F000F0u, 0x54005400u);
  const uint32_t t = w >> 6;
  h[2] = lop3_and_or(t, 0x003C003Cu, 0x5C005C00u);
  h[3] = lop3_and_or(t, 0x03C003C0u, 0x4C004C00u);
}
// p + sum over the 16 codes of a 2-bit word;  x16 = 16 activations as 8 half2 words.  Four
// independent accumulation chains: FHFMA has a long dependent-issue latency and a CTA runs only
// 3-4 warps per scheduler.
__device__ __forceinline__ float dot16_2b(const uint32_t* h, const uint32_t* x16, float p) {
  float q0 = p, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    q0 = fhfma_sel(h[j], 0, x16[j >> 1], j & 1, q0);
    q1 = fhfma_sel(h[j], 1, x16[(j + 8) >> 1], j & 1, q1);
    q2 = fhfma_sel(h[j + 1], 0, x16[(j + 1) >> 1], (j + 1) & 1, q2);
    q3 = fhfma_sel(h[j + 1], 1, x16[(j + 9) >> 1], (j + 1) & 1, q3);
  }
  return (q0 + q1) + (q2 + q3);
}
// p + sum over the 8 nibbles of two 4-bit words;  x16 = the pool's 16 activations
__device__ __forceinline__ float dot16_4b(const uint32_t* ha, const uint32_t* hb, const uint32_t* x16,
                                          float p) {
  float q0 = p, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q0 = fhfma_sel(ha[j], 0, x16[j >> 1], j & 1, q0);
    q1 = fhfma_sel(ha[j], 1, x16[(j + 4) >> 1], j & 1, q1);
    q2 = fhfma_sel(hb[j], 0, x16[4 + (j >> 1)], j & 1, q2);
    q3 = fhfma_sel(hb[j], 1, x16[4 + ((j + 4) >> 1)], j & 1, q3);
  }
  return (q0 + q1) + (q2 + q3);
}
// 4.0f + the 2-bit field of v at bit position pos (exact, no int->float conversion)
__device__ __forceinline__ float four_plus_field(uint32_t v, int pos) {
  return __uint_as_float(((v & (3u << pos)) << (21 - pos)) + 0x40800000u);
}
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t b, float c) {   // lo(a) * lo(b) + c
  return fhfma_sel(a, 0, b, 0, c);
}

// ---------------------------------------------------------------------------------------------
// Weight stream.  The rows of a CTA are one contiguous range of every packed tensor, so a "stage"
// (rpr consecutive 4-row groups, all of K) is fetched with four 1-D TMA bulk copies
// (cp.async.bulk, UBLKCP) into a shared-memory ring, completion on an mbarrier:
//   [0)       weight            group g, row r, block b at (g*4 + r)*nblk*16 + b*16
//   [off_wl)  weight_last       (g*4 + r)*nblk*4 + b*4
//   [off_zs)  zeros_and_scales  (g*4 + r)*128*nchunk + word*4 (+2 for the upper half)
//   [off_z2)  zeros_2nd         g*128*nchunk + word*4 (+ byte)
//   [off_s2)  scales_2nd        the 16-byte-aligned range enclosing the stage's groups (they start
//                               on 2-byte boundaries): (g*nblk + b)*6 + 2k + slop
// ---------------------------------------------------------------------------------------------
struct GemvStage {
  int off_wl, off_zs, off_z2, off_s2, bytes;
};
__host__ __device__ inline GemvStage gemv_stage_layout(int rpr, int nblk, int nchunk) {
  GemvStage L;
  L.off_wl = rpr * nblk * 64;
  L.off_zs = L.off_wl + rpr * nblk * 16;
  L.off_z2 = L.off_zs + rpr * 512 * nchunk;
  L.off_s2 = L.off_z2 + rpr * 128 * nchunk;
  // scales_2nd: rpr * nblk * 6 bytes starting on a 2-byte boundary -> fetched as the enclosing
  // 16-byte-aligned range (up to 14 bytes of slop in front, 14 behind)
  L.bytes = L.off_s2 + ((rpr * nblk * 6 + 15) & ~15) + 16;
  return L;
}
// thread 0: fetch `rc` row groups starting at `grp0` into stage `st`.  s2_total = byte size of
// the scales_2nd tensor: the aligned range is clipped to it and a clipped tail (< 16 bytes, only
// when the tensor size is not a multiple of 16) is copied with plain loads -- ordered before the
// consumers by the release of arrive.expect_tx / acquire of the barrier wait.
__device__ __forceinline__ void gemv_fill(unsigned char* st, const GemvStage& L, uint64_t* bar,
                                          const mxq_packed_t& w, int grp0, int rc, int nblk,
                                          int nchunk, size_t s2_total) {
  const uint32_t b_wq = (uint32_t)rc * nblk * 64, b_wl = (uint32_t)rc * nblk * 16,
                 b_zs = (uint32_t)rc * 512 * nchunk, b_z2 = (uint32_t)rc * 128 * nchunk;
  const size_t s2_b0 = (size_t)grp0 * nblk * 6, s2_b1 = s2_b0 + (size_t)rc * nblk * 6;
  const size_t s2_a0 = s2_b0 & ~(size_t)15;
  size_t s2_a1 = (s2_b1 + 15) & ~(size_t)15;
  if (s2_a1 > s2_total) s2_a1 = s2_total & ~(size_t)15;
  const unsigned char* s2g = reinterpret_cast<const unsigned char*>(w.scales_2nd);
  for (size_t b = s2_a1 > s2_a0 ? s2_a1 : s2_a0; b < s2_b1; b += 2)
    *reinterpret_cast<unsigned short*>(st + L.off_s2 + (b - s2_a0)) =
        *reinterpret_cast<const unsigned short*>(s2g + b);
  const uint32_t b_s2 = s2_a1 > s2_a0 ? (uint32_t)(s2_a1 - s2_a0) : 0u;
  mbar_arrive_expect_tx(bar, b_wq + b_wl + b_zs + b_z2 + b_s2);
  bulk_g2s(st, reinterpret_cast<const unsigned char*>(w.weight) + (size_t)grp0 * nblk * 64, b_wq, bar);
  bulk_g2s(st + L.off_wl, reinterpret_cast<const unsigned char*>(w.weight_last) + (size_t)grp0 * nblk * 16, b_wl, bar);
  bulk_g2s(st + L.off_zs, reinterpret_cast<const unsigned char*>(w.zeros_and_scales) + (size_t)grp0 * 512 * nchunk, b_zs, bar);
  bulk_g2s(st + L.off_z2, reinterpret_cast<const unsigned char*>(w.zeros_2nd) + (size_t)grp0 * 128 * nchunk, b_z2, bar);
  if (b_s2) bulk_g2s(st + L.off_s2, s2g + s2_a0, b_s2, bar);
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t ldg_u16(const void* p) {
  unsigned short v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}

// 4-bit pool scale / zero words of a row group (per output row: 8 + 2 bytes per group, read with
// ordinary loads one round ahead)
struct GemvPre {
  uint2 s4;
  uint32_t z4w;
};
__device__ __forceinline__ void gemv_prefetch_meta(GemvPre& pre, const mxq_packed_t& w, int grp, bool skip) {
  if (skip) return;
  const int oc0 = grp * 4;
  pre.s4 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(w.scales_4b) + oc0));
  pre.z4w = (uint32_t)__ldg(w.zeros_4b + (oc0 >> 3)) >> (4 * (oc0 & 7));
}

struct GemvRegs {     // one lane's share of a (4-row group, 64-column block): 27 registers
  uint4 wq[4];
  uint32_t wl[4], zs[4], z2, s2[3];
};
// gi = row group index inside the stage
// s2_slop = (first group of the stage * nblk * 6) & 15
__device__ __forceinline__ void gemv_read(const unsigned char* st, const GemvStage& L, int gi,
                                          int blk, int nblk, int nchunk, int s2_slop, GemvRegs& g) {
  const int chunk = blk >> 6, bp = blk & 63;
  const int word = chunk * 32 + (bp & 31), p = bp >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = gi * 4 + r;
    g.wq[r] = *reinterpret_cast<const uint4*>(st + (row * nblk + blk) * 16);
    g.wl[r] = *reinterpret_cast<const uint32_t*>(st + L.off_wl + (row * nblk + blk) * 4);
    g.zs[r] = *reinterpret_cast<const uint16_t*>(st + L.off_zs + (row * 32 * nchunk + word) * 4 + p * 2);
  }
  g.z2 = st[L.off_z2 + (gi * 32 * nchunk + word) * 4 + p];
  const unsigned short* s2p =
      reinterpret_cast<const unsigned short*>(st + L.off_s2 + s2_slop + (gi * nblk + blk) * 6);
#pragma unroll
  for (int k = 0; k < 3; ++k) g.s2[k] = s2p[k];
}

// Shared-memory image of the activations, per batch row b (stride `xb_stride` bytes):
//   [nblk][144 B]   the block's 64 fp16 activations (+ pad)
//   tabA float4[nblk] = { 4*xs_k - xsb_k (k = 0,1,2),  -xsb_3 }   xs = sum x, xsb = sum bias*x
//   tabB float4[nblk] = { -xs_k (k = 0..3) }
// ---------------------------------------------------------------------------------------------
// Integer inner product.  Measured on B200 (profiles/probes/fhfma_probe.cu): LOP3/SHF issue at one
// warp-instruction per 2 cycles per scheduler, fp16 FMAs (HFMA2, FHFMA) compete with them for the
// same issue slots, but IDP4A (dp4a) and FFMA/IMAD co-issue with them for free.  So the dot
// products run on dp4a: the CTA converts each 16-column group of activations once into block
// floating point -- int16 mantissas X_j = rint(x_j * 2^(14-E)), E = exponent of the group's
// largest |x| (exact for every element within 2^-4 of it, absolute error <= 2^-16 of it otherwise)
// -- stored as a signed high byte and an unsigned low byte; codes are unpacked four per register
// ((w >> 2c) & 0x03030303, 7 ALU instructions per 16 codes); two dp4a chains give
// sum q_j*hi_j and sum q_j*lo_j exactly, and
//     sum_j (q_j - z) x_j = 2^(E-14) * [ 256*HI + LO - z * sum_j X_j ]        (integers < 2^24)
// is converted once per (row, group) and scaled by s2*(c - z2) in fp32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {   // a: u8 x4, b: s8 x4
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {   // a, b: u8 x4
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int imad(int a, int b, int c) {
  int d;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// x8: the group's activations as bytes, [0..3] high (signed), [4..7] low (unsigned); register c
// holds elements {c, c+4, c+8, c+12} (2-bit groups) in bytes 0..3.   Returns 256*HI + LO.
__device__ __forceinline__ int idot16_2b(uint32_t w, const uint32_t* x8) {
  const uint32_t b0 = w & 0x03030303u, b1 = (w >> 2) & 0x03030303u, b2 = (w >> 4) & 0x03030303u,
                 b3 = (w >> 6) & 0x03030303u;
  int hi = dp4a_us(b0, x8[0], 0), lo = dp4a_uu(b0, x8[4], 0);
  hi = dp4a_us(b1, x8[1], hi); lo = dp4a_uu(b1, x8[5], lo);
  hi = dp4a_us(b2, x8[2], hi); lo = dp4a_uu(b2, x8[6], lo);
  hi = dp4a_us(b3, x8[3], hi); lo = dp4a_uu(b3, x8[7], lo);
  return imad(hi, 256, lo);
}
// pool: word a = columns 48..55, word b = 56..63 (nibble j at bits [4j+3:4j]); registers 0/1 hold
// the even/odd columns of the first word, 2/3 of the second.
__device__ __forceinline__ int idot16_4b(uint32_t wa, uint32_t wb, const uint32_t* x8) {
  const uint32_t a0 = wa & 0x0F0F0F0Fu, a1 = (wa >> 4) & 0x0F0F0F0Fu, c0 = wb & 0x0F0F0F0Fu,
                 c1 = (wb >> 4) & 0x0F0F0F0Fu;
  int hi = dp4a_us(a0, x8[0], 0), lo = dp4a_uu(a0, x8[4], 0);
  hi = dp4a_us(a1, x8[1], hi); lo = dp4a_uu(a1, x8[5], lo);
  hi = dp4a_us(c0, x8[2], hi); lo = dp4a_uu(c0, x8[6], lo);
  hi = dp4a_us(c1, x8[3], hi); lo = dp4a_uu(c1, x8[7], lo);
  return imad(hi, 256, lo);
}

// Shared-memory image of the activations, per batch row b (stride `xb_stride` bytes):
//   [nblk][144 B]     per block 4 groups x {4 words of high bytes, 4 words of low bytes} (+ pad)
//   tabI int4[nblk]   -sum_j X_j of the block's 4 groups
//   tabF float4[nblk] 2^(E-14) of the block's 4 groups
template <int NB>
__device__ __forceinline__ void gemv_block(const GemvRegs& g, const unsigned char* xsm,
                                           int xb_stride, int blk, int nblk, const float* s4,
                                           const int* z4, float (&acc)[4][NB]) {
  int4 tI[NB];
  float4 tF[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const unsigned char* base = xsm + (size_t)b * xb_stride + (size_t)nblk * kXBlkBytes;
    tI[b] = reinterpret_cast<const int4*>(base)[blk];
    tF[b] = reinterpret_cast<const float4*>(base + (size_t)nblk * 16)[blk];
  }
  uint32_t zsh[4];     // zeros_and_scales >> 4: c_1 at [7:6], c_2 at [9:8], z1_2 at [1:0]
#pragma unroll
  for (int r = 0; r < 4; ++r) zsh[r] = g.zs[r] >> 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t xv[NB][8];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint4* xp = reinterpret_cast<const uint4*>(xsm + (size_t)b * xb_stride +
                                                       (size_t)blk * kXBlkBytes + k * 32);
      const uint4 v0 = xp[0], v1 = xp[1];
      xv[b][0] = v0.x; xv[b][1] = v0.y; xv[b][2] = v0.z; xv[b][3] = v0.w;
      xv[b][4] = v1.x; xv[b][5] = v1.y; xv[b][6] = v1.z; xv[b][7] = v1.w;
    }
    if (k < 3) {
      // S = s2 * (c - z2) = (cb + c) * s2 - (cb + z2) * s2,  cb = 4 (k = 0, 2) or 16 (k = 1)
      const float s2f = __half2float(__ushort_as_half((unsigned short)g.s2[k]));
      const float cb_m4 = k == 1 ? 12.f : 0.f;
      const float S0 = -(four_plus_field(g.z2, 2 * k) + cb_m4) * s2f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t wk = k == 0 ? g.wq[r].x : (k == 1 ? g.wq[r].y : g.wq[r].z);
        const uint32_t hc = k == 0 ? lop3_and_or(g.zs[r], 0x0300u, 0x4400u)
                          : k == 1 ? lop3_and_or(zsh[r], 0x00C0u, 0x4C00u)
                                   : lop3_and_or(zsh[r], 0x0300u, 0x4400u);
        const float S = fhfma_lo(hc, g.s2[k], S0);                       // :136
        const int z1 = (int)(k == 0 ? (g.zs[r] & 3u) : k == 1 ? ((g.zs[r] >> 2) & 3u) : (zsh[r] & 3u));
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int nxs = k == 0 ? tI[b].x : (k == 1 ? tI[b].y : tI[b].z);
          const float xsc = k == 0 ? tF[b].x : (k == 1 ? tF[b].y : tF[b].z);
          const int d = imad(z1, nxs, idot16_2b(wk, xv[b]));
          acc[r][b] = fmaf(S * xsc, (float)d, acc[r][b]);               // :153
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int d = imad(z4[r], tI[b].w, idot16_4b(g.wq[r].w, g.wl[r], xv[b]));
          acc[r][b] = fmaf(s4[r] * tF[b].w, (float)d, acc[r][b]);       // :179,192
        }
      }
    }
  }
}

// Programmatic dependent launch (PDL): when the launch carries the programmatic-serialization
// attribute this grid may become resident while its predecessor in the stream is still running
// (the host sizes the CTA to at most half an SM when it can, so two generations fit).  Before
// griddep_wait() it only touches the packed weights -- every warp fills both stages of its ring --
// never x or y.  Contract (include/mxq_b200.h): the packed tensors must not be written by a
// kernel that itself signals early completion (griddepcontrol.launch_dependents) immediately
// before this call; MXQ_GEMV_NO_PDL otherwise.  Without the attribute both are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct GemvPlan {
  int q;        // consecutive 4-row groups per CTA
  int wpr;      // warps sharing one row group (each takes K slices sl0, sl0 + wpr, ...)
  int rpr;      // row groups per round = warps / wpr
  int rounds;   // ceil(q / rpr)
  int spw;      // K slices per warp and round = ceil(ksl / wpr)
  int ksl;      // K slices of 32 blocks (2048 columns) per row
  int nstages;  // ring depth
  int dbg;      // profiling only (MXQ_GEMV_DBG): 1 = skip the dot products, 2 = skip staging
};

constexpr int kGemvMaxWarps = 16, kGemvMaxStages = 4;

// Up to kGemvMaxGroup linears of one shape that share the activation vector (q/k/v, gate/up) run
// as ONE launch: the grid is split evenly between them, every CTA stages x once and walks a longer
// row range, so the fixed latencies of a launch (prologue, activation staging, first fill) are paid
// once per group instead of once per linear.  n == 1 is the plain GEMV.
constexpr int kGemvMaxGroup = 4;
struct GemvGroup {
  mxq_packed_t w[kGemvMaxGroup];
  __half* y[kGemvMaxGroup];
  int n;       // linears in the launch
  int gxl;     // CTAs per linear (gridDim.x = n * gxl)
  const int32_t* gperm;   // optional: 16-column group g of the packed weight reads activation group gperm[g]
};

// Profiling only (MXQ_GEMV_DBG & 8): per-CTA %globaltimer stamps {start, waited, staged, done}.
__device__ unsigned long long g_gemv_trace[4 * 160];
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Persistent: one CTA per SM walks its q row groups in rounds of rpr groups (= one ring stage).
// warp = (row group of the round, K-slice phase); lane = 64-column block of the slice.
// Dynamic shared memory: [nstages][stage] weight ring, then the activation image (gemv_block).
template <int NB>
__global__ void __launch_bounds__(NB == 1 ? 512 : 256, 1) gemv_mxq_kernel(
    const __half* __restrict__ x, const __grid_constant__ GemvGroup G, int B, int IC, int OC,
    GemvPlan plan) {
  const int li = blockIdx.x / G.gxl;                     // which linear of the group (CTA-uniform)
  const int cta = blockIdx.x - li * G.gxl;
  const mxq_packed_t w = G.w[li];
  __half* __restrict__ y = G.y[li];
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ float red[2][kGemvMaxWarps][4][NB];
  __shared__ uint64_t full[kGemvMaxStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = IC >> 6;
  const int nchunk = (nblk + 63) >> 6;
  const int ngrp_all = OC >> 2;
  const int grp_base = cta * plan.q;
  const int qc = min(plan.q, ngrp_all - grp_base);      // row groups of this CTA
  const int rounds = (qc + plan.rpr - 1) / plan.rpr;     // <= plan.rounds (the last CTA may be short)
  const int rgl = warp / plan.wpr, sl0 = warp - rgl * plan.wpr;
  const bool warp_on = rgl < plan.rpr;
  const int b0 = blockIdx.y * NB;
  const int xb_stride = nblk * (kXBlkBytes + 32);
  const GemvStage L = gemv_stage_layout(plan.rpr, nblk, nchunk);
  const int stage_stride = (L.bytes + 127) & ~127;
  const size_t s2_total = (size_t)ngrp_all * nblk * 6;
  unsigned char* xsm = smem + (size_t)plan.nstages * stage_stride;

  const bool trace = (plan.dbg & 8) && threadIdx.x == 0 && blockIdx.x < 160;
  if (trace) g_gemv_trace[blockIdx.x * 4 + 0] = gtimer_ns();
  if (threadIdx.x == 0) {
    for (int i = 0; i < plan.nstages; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    for (int i = 0; i < plan.nstages && i < rounds; ++i) {
      const int g0 = i * plan.rpr;
      if (g0 < qc)
        gemv_fill(smem + (size_t)i * stage_stride, L, &full[i], w, grp_base + g0,
                  min(plan.rpr, qc - g0), nblk, nchunk, s2_total);
    }
  }

  // 4-bit pool parameters of the warp's first row group (weights only: allowed before the wait)
  GemvPre pre;
  gemv_prefetch_meta(pre, w, grp_base + rgl, !(warp_on && rgl < qc));

  griddep_wait();
  // Only now may the next kernel of the stream become resident (one generation of look-ahead).
  // Triggering before the wait was measured slower (profiles/probes/pdl_probe.cu): the scheduler
  // then stacks several CTAs of one grid on the SMs that happen to be free.
  griddep_launch_dependents();
  if (trace) g_gemv_trace[blockIdx.x * 4 + 1] = gtimer_ns();

  // stage activations + per-group sums (one thread per (batch row, 16-column group))
  if (!(plan.dbg & 2)) {
    const int ngrp = IC >> 4;
    for (int i = threadIdx.x; i < NB * ngrp; i += blockDim.x) {
      const int b = NB == 1 ? 0 : i / ngrp, g = i - b * ngrp;
      const int bb = min(b0 + b, B - 1);
      const int gsrc = G.gperm ? __ldg(G.gperm + g) : g;      // importance-driven column order (per 16-column group)
      const uint4* xp = reinterpret_cast<const uint4*>(x + (size_t)bb * IC + (size_t)gsrc * 16);
      const uint4 v0 = __ldg(xp), v1 = __ldg(xp + 1);
      const int xblk = g >> 2, k = g & 3;
      unsigned char* xb = xsm + (size_t)b * xb_stride;
      uint4* dst = reinterpret_cast<uint4*>(xb + (size_t)xblk * kXBlkBytes + k * 32);
      const uint32_t xw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      float f[16];
      float gmax = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f[j] = __half2float(__ushort_as_half((unsigned short)(xw[j >> 1] >> (16 * (j & 1)))));
        gmax = fmaxf(gmax, fabsf(f[j]));
      }
      // block floating point: X_j = rint(x_j * 2^(14-E)), E = unbiased exponent of gmax
      gmax = fminf(gmax, 65504.f);                      // inf/nan activations: keep the bit tricks finite
      const uint32_t eb = __float_as_uint(gmax) >> 23;   // biased exponent (fp16 -> fp32 is never subnormal)
      const float up = gmax > 0.f ? __uint_as_float((268u - eb) << 23) : 0.f;
      const float xsc = gmax > 0.f ? __uint_as_float((eb - 14u) << 23) : 0.f;
      uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
      int xs = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int X = __float2int_rn(fminf(fmaxf(f[j] * up, -32767.f), 32767.f));
        xs += X;
        // 2-bit groups: register j & 3, byte j >> 2;  pool: register (j>>3)*2 + (j&1), byte (j&7)>>1
        const int reg = k < 3 ? (j & 3) : ((j >> 3) * 2 + (j & 1));
        const int byte = k < 3 ? (j >> 2) : ((j & 7) >> 1);
        hi[reg] |= ((uint32_t)(X >> 8) & 0xFFu) << (8 * byte);
        lo[reg] |= ((uint32_t)X & 0xFFu) << (8 * byte);
      }
      dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      dst[1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      int* tI = reinterpret_cast<int*>(xb + (size_t)nblk * kXBlkBytes) + xblk * 4 + k;
      float* tF = reinterpret_cast<float*>(xb + (size_t)nblk * (kXBlkBytes + 16)) + xblk * 4 + k;
      *tI = -xs;
      *tF = xsc;
    }
  }
  __syncthreads();
  if (trace) g_gemv_trace[blockIdx.x * 4 + 2] = gtimer_ns();

  for (int round = 0; round < rounds; ++round) {
    const int slot = round % plan.nstages;
    const uint32_t parity = (uint32_t)(round / plan.nstages) & 1u;
    const unsigned char* st = smem + (size_t)slot * stage_stride;
    const int gl = round * plan.rpr + rgl;
    const bool on = warp_on && gl < qc;
    const int grp = grp_base + gl;
    float acc[4][NB];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
    float s4[4];
    int z4[4];
    {
      const __half2 lo = *reinterpret_cast<const __half2*>(&pre.s4.x);
      const __half2 hi = *reinterpret_cast<const __half2*>(&pre.s4.y);
      s4[0] = __low2float(lo); s4[1] = __high2float(lo);
      s4[2] = __low2float(hi); s4[3] = __high2float(hi);
#pragma unroll
      for (int r = 0; r < 4; ++r) z4[r] = (int)((pre.z4w >> (4 * r)) & 0xF);
    }
    GemvRegs cur;
    // pool parameters of the warp's row group of the next round
    gemv_prefetch_meta(pre, w, grp + plan.rpr, !(on && round + 1 < rounds && gl + plan.rpr < qc));
    const int s2_slop = (int)(((size_t)(grp_base + round * plan.rpr) * nblk * 6) & 15);
    mbar_wait(&full[slot], parity);
    if (on) {
      for (int sl = sl0; sl < plan.ksl; sl += plan.wpr) {
        const int blk = sl * 32 + lane;
        if (blk < nblk) {
          gemv_read(st, L, rgl, blk, nblk, nchunk, s2_slop, cur);
          if (!(plan.dbg & 1)) gemv_block<NB>(cur, xsm, xb_stride, blk, nblk, s4, z4, acc);
          else acc[0][0] += __uint_as_float(cur.wq[0].x ^ cur.wq[1].y ^ cur.wq[2].z ^ cur.wq[3].w ^ cur.wl[0] ^ cur.wl[1] ^ cur.wl[2] ^ cur.wl[3] ^ cur.zs[0] ^ cur.zs[1] ^ cur.zs[2] ^ cur.zs[3] ^ cur.z2 ^ cur.s2[0] ^ cur.s2[1] ^ cur.s2[2]);
        }
      }
    }
    float (*rd)[4][NB] = red[round & 1];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float v = warp_sum(acc[r][b]);
        if (lane == 0) rd[warp][r][b] = v;
      }
    __syncthreads();      // every warp is done with the stage (and red[] is complete)
    const int t = threadIdx.x;
    if (t == 0) {
      const int g0 = (round + plan.nstages) * plan.rpr;
      if (round + plan.nstages < rounds) {
        fence_proxy_async();
        gemv_fill(smem + (size_t)slot * stage_stride, L, &full[slot], w, grp_base + g0,
                  min(plan.rpr, qc - g0), nblk, nchunk, s2_total);
      }
    }
    // one thread per (row group of the round, row, batch)
    if (t < plan.rpr * 4 * NB) {
      const int g = t / (4 * NB), r = (t / NB) & 3, b = t % NB;
      const int glw = round * plan.rpr + g;
      if (glw < qc && b0 + b < B) {
        float sum = 0.f;
        for (int k = 0; k < plan.wpr; ++k) sum += rd[g * plan.wpr + k][r][b];
        y[(size_t)(b0 + b) * OC + (size_t)(grp_base + glw) * 4 + r] = __float2half_rn(sum);
      }
    }
  }
  if (trace) g_gemv_trace[blockIdx.x * 4 + 3] = gtimer_ns();
}

// 4-bit word with one bias for all nibbles (shift per pair): nibble j at bits [4j+3:4j]; shifting
// by 6-4j puts nibble j at bits [9:6] of the low half and nibble j+4 at bits [9:6] of the high
// half: {16 + n_j, 16 + n_{j+4}}.
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

// ---------------------------------------------------------------------------------------------
// AWQ uniform 4-bit GEMV (gemv_cuda.cu:45-242): kernel[OC, IC/8] (nibble j of word i = column
// 8i+j), zeros[OC, zw] 4-bit per group (8 groups per word), scales fp16 [OC, zw*8]
// w = scale * (q - zero).
// CTA = 8 warps x 2 rows.  The activation row is staged once per CTA in shared memory, chunk-
// transposed (the j-th 16-byte piece of every 32-column chunk is contiguous, so a warp's LDS.128
// is conflict-free), together with the fp32 sum of every chunk -- sum_j (16 + q_j - 16 - z) x_j =
// sum_j (16 + q_j) x_j - (16 + z) * sum_j x_j needs the chunk sum once, not once per row.  A lane
// owns 32-column chunks (one 128-bit weight load per row, 512 contiguous bytes per warp and row)
// and keeps two chunks x two rows of loads in flight.
// ---------------------------------------------------------------------------------------------
constexpr int kAwqRows = 2;
__global__ void __launch_bounds__(256) awq_gemv_kernel(const __half* __restrict__ x,
                                                       const uint32_t* __restrict__ kernel,
                                                       const __half* __restrict__ scales,
                                                       const uint32_t* __restrict__ zeros,
                                                       __half* __restrict__ y, int IC, int OC,
                                                       int G, int zw) {
  extern __shared__ __align__(16) unsigned char awq_smem[];
  const int nch = IC / 32;                                    // 32-column chunks
  uint4* xs4 = reinterpret_cast<uint4*>(awq_smem);            // [4][nch]
  float* xsum = reinterpret_cast<float*>(awq_smem + (size_t)nch * 64);   // [nch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < nch; c += 256) {
    const uint4* xp = reinterpret_cast<const uint4*>(x + (size_t)b * IC + (size_t)c * 32);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v = __ldg(xp + j);
      xs4[j * nch + c] = v;
      const uint32_t xr[4] = {v.x, v.y, v.z, v.w};
      sum += sum_halves<4>(xr);
    }
    xsum[c] = sum;
  }
  __syncthreads();
  const int ww = IC / 8;
  const int row0 = (blockIdx.x * 8 + warp) * kAwqRows;
  float acc[kAwqRows];
#pragma unroll
  for (int i = 0; i < kAwqRows; ++i) acc[i] = 0.f;
  auto chunk = [&](int c, const uint4* wv) {
    uint32_t xr[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v = xs4[j * nch + c];
      xr[j][0] = v.x; xr[j][1] = v.y; xr[j][2] = v.z; xr[j][3] = v.w;
    }
    const float xsc = xsum[c];
    const int g = (c * 32) / G;
#pragma unroll
    for (int i = 0; i < kAwqRows; ++i) {
      const int oc = min(row0 + i, OC - 1);
      const uint32_t z = (__ldg(zeros + (size_t)oc * zw + (g >> 3)) >> (4 * (g & 7))) & 0xF;
      const float sc = __half2float(__ldg(scales + (size_t)oc * zw * 8 + g));
      const float zb = __uint_as_float(0x41800000u + (z << 19));   // 16 + zero
      const uint32_t wd[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) p = dot_word_4b(wd[j], xr[j], p);
      acc[i] = fmaf(sc, fmaf(-zb, xsc, p), acc[i]);
    }
  };
  int c = lane;
  for (; c + 32 < nch; c += 64) {               // two chunks x two rows of weight loads in flight
    uint4 w0[kAwqRows], w1[kAwqRows];
#pragma unroll
    for (int i = 0; i < kAwqRows; ++i) {
      const size_t ro = (size_t)min(row0 + i, OC - 1) * ww;
      w0[i] = ld_stream(kernel + ro + (size_t)c * 4);
      w1[i] = ld_stream(kernel + ro + (size_t)(c + 32) * 4);
    }
    chunk(c, w0);
    chunk(c + 32, w1);
  }
  for (; c < nch; c += 32) {
    uint4 w0[kAwqRows];
#pragma unroll
    for (int i = 0; i < kAwqRows; ++i)
      w0[i] = ld_stream(kernel + (size_t)min(row0 + i, OC - 1) * ww + (size_t)c * 4);
    chunk(c, w0);
  }
#pragma unroll
  for (int i = 0; i < kAwqRows; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0 && row0 + i < OC) y[(size_t)b * OC + row0 + i] = __float2half_rn(v);
  }
}

}  // namespace mxq

using namespace mxq;

namespace {

constexpr size_t kSmemPerSM = 227 * 1024, kSmemCtaReserve = 4096;   // static smem + per-CTA reserve

template <int NB>
int launch_gemv(const __half* x, const mxq_packed_t* ws, void* const* ys, int n, int B, int IC, int OC,
                const int32_t* gperm, bool pdl, cudaStream_t st) {
  constexpr int kMaxWarps = NB == 1 ? kGemvMaxWarps : 8;
  const int nblk = IC / 64, ngrp = OC / 4, nchunk = (nblk + 63) / 64;
  const size_t ximg = (size_t)NB * nblk * (kXBlkBytes + 32);
  GemvPlan plan;
  const int cpl = kNumSMs / n;                            // CTAs per linear
  plan.q = (int)ceil_div(ngrp, cpl);
  plan.ksl = (int)ceil_div(nblk, 32);
  plan.dbg = 0;
  if (const char* e = getenv("MXQ_GEMV_DBG")) plan.dbg = atoi(e);
  // Choose (warps, warps per row group, ring depth).  Cost = sequential units per warp
  // (rounds x slices), +30 % if two CTAs cannot share an SM (no overlap with the next GEMV under
  // PDL), +15 % unless the ring holds the whole CTA share or >= 2 stages and 64 KB; ties -> fewer warps.
  // (A latency model fitted to same-shape chains -- 8 warps so that two generations fit in the
  // register file -- won 13 % on 4096^2 alone but lost 17 % on the mixed-shape decode chain of
  // bench.py: profiles/sweep_gemv_grouped.py, profiles/README.md.)
  int W = 0, force_w = 0, force_wpr = 0, force_s = 0;
  if (const char* e = getenv("MXQ_GEMV_WARPS")) force_w = atoi(e);    // tuning knobs
  if (const char* e = getenv("MXQ_GEMV_WPR")) force_wpr = atoi(e);
  if (const char* e = getenv("MXQ_GEMV_STAGES")) force_s = atoi(e);
  double best = 1e30;
  size_t smem_best = 0;
  for (int warps = 4; warps <= kMaxWarps; ++warps) {
    if (force_w && warps != force_w) continue;
    for (int wpr = 1; wpr <= warps && wpr <= plan.ksl; ++wpr) {
      if (force_wpr && wpr != force_wpr) continue;
      const int rpr = warps / wpr;
      const int rounds = (int)ceil_div(plan.q, rpr);
      const size_t stage = ((size_t)gemv_stage_layout(rpr, nblk, nchunk).bytes + 127) & ~(size_t)127;
      for (int ns = 1; ns <= kGemvMaxStages; ++ns) {
        if (force_s && ns != force_s) continue;
        if (ns > rounds && ns > 1) continue;
        const size_t smem = ns * stage + ximg;
        if (smem + kSmemCtaReserve > kSmemPerSM) continue;
        const bool twice = NB == 1 && 2 * (smem + kSmemCtaReserve) <= kSmemPerSM && warps <= 16;
        const bool deep = ns >= rounds || (ns >= 2 && ns * stage >= 64 * 1024);
        const double cost = (double)rounds * (double)ceil_div(plan.ksl, wpr) * (twice ? 1.0 : 1.3) *
                                (deep ? 1.0 : 1.15) + 1e-3 * warps + 1e-4 * wpr + 1e-5 * ns;
        if (cost < best) {
          best = cost; W = warps; plan.wpr = wpr; plan.nstages = ns; smem_best = smem;
        }
      }
    }
  }
  if (W == 0) return MXQ_E_SHAPE;
  plan.rpr = W / plan.wpr;
  plan.rounds = (int)ceil_div(plan.q, plan.rpr);
  plan.spw = (int)ceil_div(plan.ksl, plan.wpr);
  GemvGroup G{};
  G.n = n;
  G.gperm = gperm;
  G.gxl = (int)ceil_div(ngrp, plan.q);
  for (int i = 0; i < n; ++i) { G.w[i] = ws[i]; G.y[i] = (__half*)ys[i]; }
  const unsigned gx = (unsigned)(n * G.gxl);
  const unsigned gy = (unsigned)ceil_div(B, NB);
  const size_t smem = smem_best;
  if (getenv("MXQ_GEMV_VERBOSE"))
    fprintf(stderr, "mxq_gemv %dx%d B=%d: warps %d wpr %d rpr %d rounds %d stages %d smem %zu grid %u\n",
            OC, IC, B, W, plan.wpr, plan.rpr, plan.rounds, plan.nstages, smem, gx);
  // Same shared-memory configuration as the IMMA kernel (gemv_mma.cu): opt-in limit raised once, carve-out
  // pinned to "max shared", so a decode chain that alternates between the two kernels (and between
  // shapes) never makes an SM re-partition L1 / shared memory between consecutive launches.
  {
    static bool configured = false;                  // benign race: idempotent attribute writes
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(gemv_mxq_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(kSmemPerSM - kSmemCtaReserve));
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(gemv_mxq_kernel<NB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               (int)cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy);
  cfg.blockDim = dim3(W * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemv_mxq_kernel<NB>, x, G, B, IC, OC, plan);
  return e == cudaSuccess ? MXQ_OK : (int)e;
}

}  // namespace

static int gemv_check_packed(const mxq_packed_t& w) {
  MXQ_CHECK_PTR(w.weight);
  if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b ||
      !w.zeros_4b)
    return MXQ_E_NULL;
  return MXQ_OK;
}

// Generic-shape path (any IC % 64 == 0) behind mxq_gemv_grouped (gemv_mma.cu dispatches).
namespace mxq {
int gemv_ring_grouped(const void* x, const mxq_packed_t* w, void* const* y, int n, int64_t B, int64_t IC,
                      int64_t OC, const int32_t* gperm, unsigned flags, void* stream) {
  if (B < 0 || IC < 0 || OC < 0 || n < 0 || n > kGemvMaxGroup) return MXQ_E_SHAPE;
  if (B == 0 || OC == 0 || n == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  if (!w || !y) return MXQ_E_NULL;
  for (int i = 0; i < n; ++i) {
    MXQ_CHECK_PTR(y[i]);
    const int rc = gemv_check_packed(w[i]);
    if (rc) return rc;
  }
  if (IC % 64 || OC % 8 || IC == 0 || IC > (1 << 24) || OC > INT32_MAX || B > 65535 * 4)
    return MXQ_E_SHAPE;
  cudaStream_t st = as_stream(stream);
  const __half* xh = (const __half*)x;
  const bool pdl = !(flags & MXQ_GEMV_NO_PDL);
  if (B == 1) return launch_gemv<1>(xh, w, y, n, (int)B, (int)IC, (int)OC, gperm, pdl, st);
  if (B == 2) return launch_gemv<2>(xh, w, y, n, (int)B, (int)IC, (int)OC, gperm, pdl, st);
  return launch_gemv<4>(xh, w, y, n, (int)B, (int)IC, (int)OC, gperm, pdl, st);
}
}  // namespace mxq

// profiling aid, not part of the documented surface: copies the stamps of the last traced launch
extern "C" __attribute__((visibility("default"))) int mxq_debug_gemv_trace(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_gemv_trace, sizeof(unsigned long long) * 4 * 160);
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
  const size_t smem = (size_t)(IC / 32) * (64 + 4);
  if (smem > 200 * 1024) return MXQ_E_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(awq_gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  awq_gemv_kernel<<<dim3((unsigned)ceil_div(OC, 8 * kAwqRows), (unsigned)B), 256, smem, as_stream(stream)>>>(
      (const __half*)x, (const uint32_t*)kernel, (const __half*)scales, (const uint32_t*)zeros,
      (__half*)y, (int)IC, (int)OC, group_size, zw);
  MXQ_LAUNCH_RESULT();
}
