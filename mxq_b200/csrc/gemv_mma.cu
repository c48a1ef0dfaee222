// Decode GEMV for the packed mixed 2/4-bit layout on the integer tensor cores (sm_100a).
//
// Replaces mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-273 (one scalar cvt + FMA
// chain per weight, IC hard-wired to 4096).  Round-1's kernel (csrc/gemv.cu, kept as the
// generic-shape path) unpacked the codes into dp4a operands at 3.8 instructions per weight and
// was issue/latency bound at 0.28 of the HBM rate.  Here the inner products run on
// mma.sync.m16n8k32 (IMMA.16832.U8.S8, measured 0.5 per cycle per SM:
// profiles/r2_probe_imma_ldgsts.txt), which leaves the unpack ((w >> 2c) & 0x03030303: 7 ALU
// instructions per 16 codes) and one scale/zero epilogue per (row, 16-column group):
//
//   * A (16 x 32, u8): mma row m = one output row, the quad's thread t supplies the 8 codes
//     {c, c+4, c+8, c+12 | c = 2i, 2i+1} of ITS OWN 64-column block b0+t -- so one mma mixes four
//     blocks along k;
//   * B (32 x 8, s8): the activations in block floating point (int16 mantissas per 16-column
//     group, split in a signed high and a signed low byte).  Column n = 2t' + h carries the
//     high (h=0) / low (h=1) bytes of block b0+t' on the k slots of thread t' and ZERO on all
//     other k slots, which separates the four blocks again: D[m][2t+h] is the exact integer
//     sum over the 16 codes of (row m, block b0+t, group r) -- and lands in the registers of the
//     thread that holds that (row, block)'s scale / zero words;
//   * two mmas (codes c = 0,1 and c = 2,3) per group round, three 2-bit rounds + one 4-bit
//     round per unit of 16 rows x 256 columns;
//   * mma rows are labelled so that a thread's two rows (m = g, g + 8) lie in the same
//     second-order 4-row group: s2 * (c - z2) needs one s2 / z2 decode per thread and round.
//
// Weight stream: every warp owns a contiguous range of units and copies exactly its own units
// with cp.async (LDGSTS, 16 B per lane-instruction) into a private shared-memory ring, one
// commit group per unit -- no block barrier and no mbarrier in the main loop; per-warp rings
// stream at the full HBM rate (6.9 TB/s in the probe).  The copies of the first `ring` units
// are issued BEFORE griddepcontrol.wait, i.e. while the previous kernel of the stream is still
// running (programmatic dependent launch).  The activation image is built once per CTA after
// the wait.  K-partials of the 8 warps are added in a fixed order (deterministic).
//
// Shapes: IC % 256 == 0 (16-byte alignment of the quad's weight_last / scales_2nd pieces); other
// IC % 64 == 0 shapes take the generic kernel in gemv.cu.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace mxq {

// generic-shape kernel (gemv.cu)
int gemv_ring_grouped(const void* x, const mxq_packed_t* w, void* const* y, int n, int64_t B, int64_t IC,
                      int64_t OC, const int32_t* gperm, unsigned flags, void* stream);

namespace g2 {

constexpr int kRing = 4;         // ring slots per warp (57 KB per CTA; the probe streams at full rate with 4)
constexpr int kMaxGroup = 4;
// one unit = 16 rows (4 second-order groups) x 4 blocks (256 columns); slot layout in bytes
constexpr int kOffW = 0;        // [16 mma rows][4 blocks][16 B]   weight words 4*blk .. 4*blk+3
constexpr int kOffWL = 1024;    // [16 mma rows][4 blocks][4 B]    weight_last
constexpr int kOffZS = 1280;    // [16 mma rows][4 words]          zeros_and_scales words w0 .. w0+3
constexpr int kOffZ2 = 1536;    // [4 row groups][4 words]         zeros_2nd
constexpr int kOffS2 = 1600;    // [4 row groups][4 blocks][3 fp16] scales_2nd
constexpr int kUnitBytes = 1792;  // 1696 used

struct Plan {
  int q;        // 4-row groups per CTA
  int T;        // 16-row tiles per CTA = ceil(q / 4)
  int nqb;      // 256-column quad-blocks per row
  int gxl;      // CTAs per linear
  int ximg;     // bytes of one batch row's activation image
  int dbg;      // profiling only (MXQ_GEMV_DBG & 8): per-CTA %globaltimer stamps
  int early;    // griddepcontrol.launch_dependents at the top of the kernel instead of after the wait
};

// Profiling only: {start, copies issued, waited, staged, loop done, done} per CTA.
__device__ unsigned long long g_trace[6 * 160];
__device__ long long g_itrace[16 * 16 * 6];     // [warp][iteration][stamp] clock64 of CTA 1 (MXQ_GEMV_DBG & 64)
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Group {
  mxq_packed_t w[kMaxGroup];
  __half* y[kMaxGroup];
  const int32_t* gperm;   // optional: 16-column group g of the packed weight reads activation group gperm[g]
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {   // warp-uniform
  switch (pending) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    default: cp_async_wait<5>(); break;
  }
}

__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                     uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int imad(int a, int b, int c) {
  int d;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(orv));
  return d;
}
// lo(a) * lo(b) + c with fp16 operands and an fp32 accumulator (exact product)
__device__ __forceinline__ float fhfma_lo(uint32_t a2, uint32_t b2, float acc) {
  unsigned short a, b;
  asm("{.reg .b16 t; mov.b32 {%0,t}, %1;}" : "=h"(a) : "r"(a2));
  asm("{.reg .b16 t; mov.b32 {%0,t}, %1;}" : "=h"(b) : "r"(b2));
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(b));
  return acc;
}
// 4.0f + the 2-bit field of v at bit position pos (exact, no int->float conversion)
__device__ __forceinline__ float four_plus_field(uint32_t v, int pos) {
  return __uint_as_float(((v & (3u << pos)) << (21 - pos)) + 0x40800000u);
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// mma row label m (0..15) -> row offset inside the 16-row tile: rows m and m + 8 of a thread
// share one 4-row second-order group
__device__ __forceinline__ int tile_row(int m) { return ((m & 7) >> 1) * 4 + (m & 1) + 2 * (m >> 3); }

// ---------------------------------------------------------------------------------------------
// Activation image, per batch row (stride plan.ximg):
//   [16 B zeros]
//   [ngrp16][32 B]   per 16-column group: 16 B of signed high bytes, 16 B of signed low bytes;
//                    2-bit groups: register c, byte j = element 4j + c; pooled group: register
//                    (e>>3)*2 + (e&1), byte (e&7)>>1 = element e
//   tabI int4[nblk4]   -sum_j X_j of the block's 4 groups
//   tabF float4[nblk4] 2^(E-14) of the block's 4 groups
// X_j = rint(x_j * 2^(14-E)), E = exponent of 1.0078 * max|x| of the group, so |X| <= 32514 and
// X = 256 * hi + lo with both bytes signed (lo in [-128,127], hi = (X + 128) >> 8 in [-127,127]).
// Elements within 2^-4 of the group maximum are exact; the others carry an absolute error of at
// most 2^-16 of it (measured against the fp64 oracle in tests/test_gpu_packed.py).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_group(const __half* __restrict__ xg, bool live, int k,
                                            unsigned char* dst, int* tI, float* tF) {
  uint32_t xw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(xg));
    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(xg) + 1);
    xw[0] = v0.x; xw[1] = v0.y; xw[2] = v0.z; xw[3] = v0.w;
    xw[4] = v1.x; xw[5] = v1.y; xw[6] = v1.z; xw[7] = v1.w;
  }
  // group maximum of |x| on packed halves
  uint32_t m2 = xw[0] & 0x7FFF7FFFu;
#pragma unroll
  for (int i = 1; i < 8; ++i) m2 = P16<__half>::vmax(m2, xw[i] & 0x7FFF7FFFu);
  float gmax = fmaxf(P16<__half>::lo(m2), P16<__half>::hi(m2));
  gmax = fminf(gmax, 65504.f) * 1.0078125f;
  const uint32_t eb = __float_as_uint(gmax) >> 23;          // biased exponent, never subnormal in fp32
  const float up = gmax > 0.f ? __uint_as_float((268u - eb) << 23) : 0.f;
  const float xsc = gmax > 0.f ? __uint_as_float((eb - 14u) << 23) : 0.f;
  // bits(fma(x, up, 1.5*2^23 + 128)) = 0x4B400080 + X: byte 0 ^ 0x80 = low byte, byte 1 = (X+128)>>8
  uint32_t F[16];
  uint32_t xs = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float f = __half2float(__ushort_as_half((unsigned short)(xw[e >> 1] >> (16 * (e & 1)))));
    F[e] = __float_as_uint(__fmaf_rn(f, up, 12583040.0f));
    xs += F[e];
  }
  xs -= 16u * 0x4B400080u;                                   // sum_j X_j (mod 2^32, exact)
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int e0, e1, e2, e3;
    if (k < 3) { e0 = c; e1 = 4 + c; e2 = 8 + c; e3 = 12 + c; }
    else { const int b = (c >> 1) * 8 + (c & 1); e0 = b; e1 = b + 2; e2 = b + 4; e3 = b + 6; }
    const uint32_t P = prmt_b32(F[e0], F[e1], 0x5140u);      // {lo0, lo1, hi0, hi1}
    const uint32_t Q = prmt_b32(F[e2], F[e3], 0x5140u);
    lo[c] = prmt_b32(P, Q, 0x5410u) ^ 0x80808080u;
    hi[c] = prmt_b32(P, Q, 0x7632u);
  }
  reinterpret_cast<uint4*>(dst)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  reinterpret_cast<uint4*>(dst)[1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *tI = -(int)xs;
  *tF = xsc;
}

// ---------------------------------------------------------------------------------------------
// The kernel.  Shared memory: [kWarps][kRing][unit] rings | per-row tables (scales_4b, zeros_4b
// words of the CTA's rows) | activation images | red[kWarps][NB][rows].
// cp.async groups per thread, in commit order: {per-row tables}, then one group per unit (empty
// groups pad the prologue and the tail so that "at most kRing - 1 groups pending" always means
// "the unit about to be consumed has landed").
// ---------------------------------------------------------------------------------------------
template <int NB, bool kDbg, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 2) gemv_mma_kernel(const __half* __restrict__ x,
                                                               const __grid_constant__ Group G,
                                                               int B, int IC, int OC, const Plan plan) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kThreads = kWarps * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int li = blockIdx.x / plan.gxl;
  const int cta = blockIdx.x - li * plan.gxl;
  const mxq_packed_t w = G.w[li];
  __half* __restrict__ y = G.y[li];
  const int nblk = IC >> 6;
  const int nchunk = (nblk + 63) >> 6;
  const int ngrp_all = OC >> 2;
  const int grp_base = cta * plan.q;
  const int qc = min(plan.q, ngrp_all - grp_base);         // 4-row groups of this CTA (>= 1)
  const int grp_last = grp_base + qc - 1;
  const int T = (qc + 3) >> 2;
  const int nqb = plan.nqb;
  const int U = T * nqb;
  const int u_begin = (int)(((long long)warp * U) / kWarps);
  const int u_end = (int)(((long long)(warp + 1) * U) / kWarps);
  const int b0g = blockIdx.y * NB;
  const int rows_cta = plan.q * 4;

  unsigned char* ring = smem + (size_t)warp * kRing * kUnitBytes;
  unsigned char* rowtab = smem + (size_t)kWarps * kRing * kUnitBytes;    // [rows_cta] fp16 | z4 words
  const int rowtab_bytes = (rows_cta * 2 + (rows_cta / 8 + 2) * 4 + 15) & ~15;
  unsigned char* ximg = rowtab + rowtab_bytes;
  float* red = reinterpret_cast<float*>(ximg + (size_t)NB * plan.ximg);
  const uint32_t ring_s = smem_u32(ring);
  const bool trace = kDbg && (plan.dbg & 8) && threadIdx.x == 0 && blockIdx.x < 160 && blockIdx.y == 0;
  if (trace) g_trace[blockIdx.x * 6 + 0] = gtimer_ns();
  if (plan.early) griddep_launch_dependents();

  // ---- copy pattern: per-lane byte offsets relative to (first row group of the tile, ks = 0) ----
  const uint32_t w_row = (uint32_t)nblk * 16, wl_row = (uint32_t)nblk * 4, zs_row = (uint32_t)nchunk * 128,
                 s2_grp = (uint32_t)nblk * 6;
  const unsigned char* Wp = reinterpret_cast<const unsigned char*>(w.weight);
  const unsigned char* S2p = reinterpret_cast<const unsigned char*>(w.scales_2nd);
  const unsigned char* Z2p = reinterpret_cast<const unsigned char*>(w.zeros_2nd);
  // (1,2) weight: lane (g, t) copies the two 16-byte pieces it reads itself (mma rows g, g + 8)
  const uint32_t offW = ((uint32_t)(g >> 1) * 4 + (g & 1)) * w_row + t * 16;
  const uint32_t dstW = kOffW + g * 64 + t * 16;
  // (3) lanes 0-15: weight_last of mma row m3, lanes 16-31: zeros_and_scales words of mma row m3
  const int m3 = lane & 15;
  const bool is_wl = lane < 16;
  const unsigned char* P3 = reinterpret_cast<const unsigned char*>(is_wl ? (const void*)w.weight_last : (const void*)w.zeros_and_scales);
  const uint32_t row3 = is_wl ? wl_row : zs_row;
  const uint32_t off3 = ((uint32_t)((m3 & 7) >> 1) * 4 + (m3 & 1) + 2 * (m3 >> 3)) * row3;
  const uint32_t dst3 = (is_wl ? kOffWL : kOffZS) + m3 * 16;
  // (4) lanes 0-3: zeros_2nd of row group `lane`; (5) lanes 4-15: scales_2nd, 8 bytes each
  const int l5 = lane - 4;
  const uint32_t dst5 = kOffS2 + (l5 / 3) * 24 + (l5 % 3) * 8;
  const uint32_t off5 = (l5 % 3) * 8;

  // running state of the issue side
  int i_tile = 0, i_ks = 0, i_slot = 0, issued = u_begin;
  if (u_begin < u_end) { i_tile = u_begin / nqb; i_ks = u_begin - i_tile * nqb; }
  auto issue = [&]() {
    if (issued < u_end && !(kDbg && (plan.dbg & 4))) {
      const uint32_t s = ring_s + (uint32_t)i_slot * kUnitBytes;
      const int rg0 = grp_base + i_tile * 4;
      const int zoff = (i_ks >> 4) * 128 + (i_ks & 7) * 16;
      if (rg0 + (g >> 1) <= grp_last) {
        const unsigned char* src = Wp + (size_t)rg0 * 4 * w_row + offW + (uint32_t)i_ks * 64;
        cp_async16(s + dstW, src);
        cp_async16(s + dstW + 512, src + 2 * w_row);
      }
      if (rg0 + ((m3 & 7) >> 1) <= grp_last)
        cp_async16(s + dst3, P3 + (size_t)rg0 * 4 * row3 + off3 + (is_wl ? (uint32_t)i_ks * 16 : (uint32_t)zoff));
      if (lane < 4) {
        cp_async16(s + kOffZ2 + lane * 16, Z2p + (size_t)min(rg0 + lane, grp_last) * zs_row + zoff);
      } else if (lane < 16) {
        cp_async8(s + dst5, S2p + (size_t)min(rg0 + l5 / 3, grp_last) * s2_grp + (uint32_t)i_ks * 24 + off5);
      }
      ++issued;
      if (++i_ks == nqb) { i_ks = 0; ++i_tile; }
      if (++i_slot == kRing) i_slot = 0;
      cp_async_commit();
    }
  };

  // ---- prologue: weights only (allowed before the dependency wait) -----------------------------
  {
    // per-row tables of the CTA: scales_4b (8 bytes per row group), zeros_4b words
    const unsigned char* S4p = reinterpret_cast<const unsigned char*>(w.scales_4b);
    const unsigned char* Z4p = reinterpret_cast<const unsigned char*>(w.zeros_4b);
    for (int i = threadIdx.x; i < qc; i += kThreads) cp_async8(smem_u32(rowtab) + i * 8, S4p + (size_t)(grp_base + i) * 8);
    const int nz = qc / 2 + 2, z0 = grp_base >> 1;
    for (int i = threadIdx.x; i < nz; i += kThreads)
      cp_async4(smem_u32(rowtab) + rows_cta * 2 + i * 4, Z4p + (size_t)min(z0 + i, (OC >> 3) - 1) * 4);
    cp_async_commit();
  }
  // units issued before the dependency wait (the rest right after it): profiling knob, default all
  const int pre = kDbg ? ((plan.dbg >> 8) & 7 ? min((plan.dbg >> 8) & 7, kRing) : kRing) : kRing;
#pragma unroll
  for (int i = 0; i < kRing; ++i) if (i < pre) issue();
  for (int i = threadIdx.x; i < kWarps * NB * rows_cta; i += kThreads) red[i] = 0.f;

  if (trace) g_trace[blockIdx.x * 6 + 1] = gtimer_ns();
  griddep_wait();
  if (!plan.early) griddep_launch_dependents();
  if (trace) g_trace[blockIdx.x * 6 + 2] = gtimer_ns();
#pragma unroll
  for (int i = 0; i < kRing; ++i) if (i >= pre) issue();

  // ---- activation image ------------------------------------------------------------------------
  {
    const int ng = nqb * 16;                                 // groups incl. the zero padding
    for (int i = threadIdx.x; i < NB * ng; i += kThreads) {
      const int b = NB == 1 ? 0 : i / ng, gi = i - b * ng;
      const int bb = min(b0g + b, B - 1);
      unsigned char* xb = ximg + (size_t)b * plan.ximg;
      const bool live = gi * 16 < IC;
      const int gsrc = (live && G.gperm) ? __ldg(G.gperm + gi) : gi;
      stage_group(x + (size_t)bb * IC + (size_t)gsrc * 16, live, gi & 3, xb + 16 + (size_t)gi * 32,
                  reinterpret_cast<int*>(xb + 16 + (size_t)ng * 32) + gi,
                  reinterpret_cast<float*>(xb + 16 + (size_t)ng * 36) + gi);
    }
    if (threadIdx.x < NB * 4) reinterpret_cast<uint32_t*>(ximg + (size_t)(threadIdx.x >> 2) * plan.ximg)[threadIdx.x & 3] = 0u;
  }
  if (issued - u_begin >= kRing) cp_async_wait<kRing>(); else cp_async_wait<0>();   // this thread's piece of the row tables
  __syncthreads();
  if (trace) g_trace[blockIdx.x * 6 + 3] = gtimer_ns();

  // ---- main loop ---------------------------------------------------------------------------------
  const bool bact = (g >> 1) == t;                           // this lane feeds a non-zero B column
  const int ng32 = nqb * 16 * 32;
  const int i0 = (g >> 1) * 4 + (g & 1);                     // row offsets (in the tile) of m = g, g + 8: i0, i0 + 2
  float acc[NB][2];
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = 0.f;
  int tile = i_tile, ks = i_ks, slot = 0;                    // consume side (recomputed below)
  if (u_begin < u_end) { tile = u_begin / nqb; ks = u_begin - tile * nqb; }
  float s4[2] = {0.f, 0.f};
  int z4[2] = {0, 0};
  auto tile_params = [&](int tl) {                          // 4-bit pool scale / zero of this thread's two rows
    const int r = tl * 16 + i0;
    const __half* s4p = reinterpret_cast<const __half*>(rowtab);
    const uint32_t* z4w = reinterpret_cast<const uint32_t*>(rowtab + rows_cta * 2);
    if (tl * 4 + (g >> 1) < qc) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = grp_base * 4 + r + 2 * j;
        s4[j] = __half2float(s4p[r + 2 * j]);
        z4[j] = (int)((z4w[(row >> 3) - (grp_base >> 1)] >> (4 * (row & 7))) & 0xF);
      }
    }
  };
  auto flush = [&](int tl) {
    // quad reduction over the four blocks, then lane t == 0 stores this warp's tile partials
    const bool ok = tl * 4 + (g >> 1) < qc;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v = acc[b][j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t == 0 && ok) red[((size_t)warp * NB + b) * rows_cta + (size_t)(tl * 16 + i0 + 2 * j)] = v;
        acc[b][j] = 0.f;
      }
    }
  };
  if (u_begin < u_end) tile_params(tile);

  const bool itrace = kDbg && (plan.dbg & 64) && blockIdx.x == 1 && blockIdx.y == 0 && lane == 0;
  for (int u = u_begin; u < u_end; ++u) {
    long long* it = g_itrace + ((size_t)warp * 16 + min(u - u_begin, 15)) * 6;
    if (itrace) it[0] = clock64();
    // groups of this thread, oldest first: row tables, then one per issued unit.  While units are
    // still being issued exactly kRing - 1 younger groups follow the one consumed now; afterwards
    // everything outstanding is awaited at once (at most kRing - 1 units, all needed soon).
    if (issued < u_end) cp_async_wait<kRing - 1>(); else cp_async_wait<0>();
    __syncwarp();
    if (itrace) it[1] = clock64();
    const unsigned char* s = ring + (size_t)slot * kUnitBytes;
    const int p = (ks >> 3) & 1;                             // half-word / byte of the metadata words
    // A side
    const uint4 wa = *reinterpret_cast<const uint4*>(s + kOffW + g * 64 + t * 16);
    const uint4 wb = *reinterpret_cast<const uint4*>(s + kOffW + (g + 8) * 64 + t * 16);
    const uint32_t wla = *reinterpret_cast<const uint32_t*>(s + kOffWL + g * 16 + t * 4);
    const uint32_t wlb = *reinterpret_cast<const uint32_t*>(s + kOffWL + (g + 8) * 16 + t * 4);
    uint32_t zs[2];
    zs[0] = *reinterpret_cast<const uint16_t*>(s + kOffZS + g * 16 + t * 4 + p * 2);
    zs[1] = *reinterpret_cast<const uint16_t*>(s + kOffZS + (g + 8) * 16 + t * 4 + p * 2);
    const uint32_t z2 = s[kOffZ2 + (g >> 1) * 16 + t * 4 + p];
    const unsigned short* s2p = reinterpret_cast<const unsigned short*>(s + kOffS2 + (g >> 1) * 24 + t * 6);
    const uint32_t s2h[3] = {s2p[0], s2p[1], s2p[2]};
    const uint32_t zsh[2] = {zs[0] >> 4, zs[1] >> 4};
    const int blk = ks * 4 + t;
    const uint32_t xoff = bact ? (uint32_t)(16 + blk * 128 + (g & 1) * 16) : 0u;
    const uint32_t xstep = bact ? 32u : 0u;
    if (itrace) it[2] = clock64() + ((wa.x ^ wb.w ^ wla ^ wlb ^ zs[0] ^ zs[1] ^ z2 ^ s2h[2]) == 0x12345u);
    if (kDbg && (plan.dbg & 1)) {                            // profiling: data movement only
      acc[0][0] += __uint_as_float(wa.x ^ wa.y ^ wa.z ^ wa.w ^ wb.x ^ wb.y ^ wb.z ^ wb.w ^ wla ^ wlb ^ zs[0] ^ zs[1] ^ z2 ^
                                   s2h[0] ^ s2h[1] ^ s2h[2]);
    } else {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const unsigned char* xb = ximg + (size_t)b * plan.ximg;
        const int4 tI = *reinterpret_cast<const int4*>(xb + 16 + ng32 + blk * 16);
        const float4 tF = *reinterpret_cast<const float4*>(xb + 16 + ng32 + (ng32 >> 3) + blk * 16);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint4 X = *reinterpret_cast<const uint4*>(xb + xoff + k * xstep);
          const uint32_t wka = k == 0 ? wa.x : (k == 1 ? wa.y : wa.z);
          const uint32_t wkb = k == 0 ? wb.x : (k == 1 ? wb.y : wb.z);
          int d[4] = {0, 0, 0, 0};
          imma(d, wka & 0x03030303u, wkb & 0x03030303u, (wka >> 2) & 0x03030303u, (wkb >> 2) & 0x03030303u, X.x, X.y);
          imma(d, (wka >> 4) & 0x03030303u, (wkb >> 4) & 0x03030303u, (wka >> 6) & 0x03030303u,
               (wkb >> 6) & 0x03030303u, X.z, X.w);
          // S = s2 * (c - z2) = (cb + c) * s2 - (cb + z2) * s2,  cb = 4 (k = 0, 2) or 16 (k = 1)
          const float s2f = __half2float(__ushort_as_half((unsigned short)s2h[k]));
          const float S0 = -(four_plus_field(z2, 2 * k) + (k == 1 ? 12.f : 0.f)) * s2f;
          const int nxs = k == 0 ? tI.x : (k == 1 ? tI.y : tI.z);
          const float xsc = k == 0 ? tF.x : (k == 1 ? tF.y : tF.z);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t hc = k == 0 ? lop3_and_or(zs[j], 0x0300u, 0x4400u)
                              : k == 1 ? lop3_and_or(zsh[j], 0x00C0u, 0x4C00u)
                                       : lop3_and_or(zsh[j], 0x0300u, 0x4400u);
            const float S = fhfma_lo(hc, s2h[k], S0);                          // gemv_mxq_cuda.cu:136
            const int z1 = (int)(k == 0 ? (zs[j] & 3u) : k == 1 ? ((zs[j] >> 2) & 3u) : (zsh[j] & 3u));
            const int dd = imad(z1, nxs, imad(d[2 * j], 256, d[2 * j + 1]));
            acc[b][j] = fmaf(S * xsc, (float)dd, acc[b][j]);                   // :153
          }
        }
        {
          const uint4 X = *reinterpret_cast<const uint4*>(xb + xoff + 3 * xstep);
          int d[4] = {0, 0, 0, 0};
          imma(d, wa.w & 0x0F0F0F0Fu, wb.w & 0x0F0F0F0Fu, (wa.w >> 4) & 0x0F0F0F0Fu, (wb.w >> 4) & 0x0F0F0F0Fu, X.x, X.y);
          imma(d, wla & 0x0F0F0F0Fu, wlb & 0x0F0F0F0Fu, (wla >> 4) & 0x0F0F0F0Fu, (wlb >> 4) & 0x0F0F0F0Fu, X.z, X.w);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int dd = imad(z4[j], tI.w, imad(d[2 * j], 256, d[2 * j + 1]));
            acc[b][j] = fmaf(s4[j] * tF.w, (float)dd, acc[b][j]);              // :179,192
          }
        }
      }
    }
    if (itrace) it[3] = clock64() + (acc[0][0] == 1.2345f);
    __syncwarp();                                            // every lane is done with the slot
    issue();
    if (itrace) it[4] = clock64();
    if (++slot == kRing) slot = 0;
    if (++ks == nqb) {                                       // tile finished (for this warp)
      flush(tile);
      ks = 0;
      ++tile;
      if (u + 1 < u_end) tile_params(tile);
    } else if (u + 1 == u_end) {
      flush(tile);
    }
    if (itrace) it[5] = clock64();
  }
  __syncthreads();
  if (trace) g_trace[blockIdx.x * 6 + 4] = gtimer_ns();

  // ---- K partials of the 8 warps, fixed order ----------------------------------------------------
  for (int i = threadIdx.x; i < NB * qc * 4; i += kThreads) {
    const int b = i / (qc * 4), r = i - b * (qc * 4);
    if (b0g + b < B) {
      float sum = 0.f;
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) sum += red[((size_t)wv * NB + b) * rows_cta + r];
      y[(size_t)(b0g + b) * OC + (size_t)grp_base * 4 + r] = __float2half_rn(sum);
    }
  }
  if (trace) g_trace[blockIdx.x * 6 + 5] = gtimer_ns();
}

constexpr size_t kSmemPerSM = 227 * 1024, kSmemCtaReserve = 1024;

template <int NB, bool kDbg, int kWarps>
int launch_k(const __half* x, const Group& G, int n, int B, int IC, int OC, bool pdl, const Plan& plan, size_t smem,
             cudaStream_t st) {
  // One shared-memory configuration for every shape: the opt-in limit is raised once to half an SM
  // and the carve-out is pinned to "max shared", so consecutive GEMVs of different shapes never make
  // the SM re-partition L1 / shared memory (which would serialise them under PDL).
  static bool configured = false;                    // benign race: idempotent attribute writes
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemv_mma_kernel<NB, kDbg, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSmemPerSM - kSmemCtaReserve));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemv_mma_kernel<NB, kDbg, kWarps>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n * plan.gxl), (unsigned)ceil_div(B, NB));
  cfg.blockDim = dim3(kWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemv_mma_kernel<NB, kDbg, kWarps>, x, G, B, IC, OC, plan);
  return e == cudaSuccess ? MXQ_OK : (int)e;
}

template <int NB>
int launch(const __half* x, const mxq_packed_t* ws, void* const* ys, int n, int B, int IC, int OC,
           const int32_t* gperm, bool pdl, cudaStream_t st) {
  const int nblk = IC / 64, ngrp = OC / 4;
  Plan plan;
  const int cpl = kNumSMs / n;
  plan.q = (int)ceil_div(ngrp, cpl);
  plan.T = (plan.q + 3) / 4;
  plan.nqb = nblk / 4;
  plan.gxl = (int)ceil_div(ngrp, plan.q);
  const int ng = plan.nqb * 16;
  plan.ximg = ((16 + ng * 32 + ng * 8) + 127) & ~127;
  plan.dbg = 0;
  if (const char* e = getenv("MXQ_GEMV_DBG")) plan.dbg = atoi(e);
  plan.early = 0;
  if (const char* e = getenv("MXQ_GEMV_EARLY")) plan.early = atoi(e);
  const int rows_cta = plan.q * 4;
  const size_t rowtab = (size_t)((rows_cta * 2 + (rows_cta / 8 + 2) * 4 + 15) & ~15);
  // 8 warps (110 registers) or 12 warps (78 registers): two CTAs -- this launch and, under PDL, the
  // next one -- fit on an SM either way.  12 warps build the activation image of a long row in fewer
  // rounds and split the units 12 ways; with only a few units per warp 8 is better balanced.
  int warps = ((int64_t)plan.T * plan.nqb >= 64 && NB == 1) ? 12 : 8;
  if (const char* e = getenv("MXQ_GEMV_WARPS2")) { const int v = atoi(e); if (v == 8 || v == 12) warps = v; }
  const size_t smem = (size_t)warps * kRing * kUnitBytes + rowtab + (size_t)NB * plan.ximg +
                      (size_t)warps * NB * rows_cta * sizeof(float);
  if (smem + kSmemCtaReserve > kSmemPerSM) return MXQ_E_SHAPE;
  Group G{};
  G.gperm = gperm;
  for (int i = 0; i < n; ++i) { G.w[i] = ws[i]; G.y[i] = (__half*)ys[i]; }
  if (getenv("MXQ_GEMV_VERBOSE"))
    fprintf(stderr, "mxq_gemv(mma) %dx%d B=%d n=%d: q %d T %d nqb %d units/warp %d smem %zu grid %d\n", OC, IC, B, n,
            plan.q, plan.T, plan.nqb, (int)ceil_div((int64_t)plan.T * plan.nqb, warps), smem, n * plan.gxl);
  if (plan.dbg) return warps == 12 ? launch_k<NB, true, 12>(x, G, n, B, IC, OC, pdl, plan, smem, st)
                                   : launch_k<NB, true, 8>(x, G, n, B, IC, OC, pdl, plan, smem, st);
  return warps == 12 ? launch_k<NB, false, 12>(x, G, n, B, IC, OC, pdl, plan, smem, st)
                     : launch_k<NB, false, 8>(x, G, n, B, IC, OC, pdl, plan, smem, st);
}

}  // namespace g2
}  // namespace mxq

using namespace mxq;

static int gemv_check_packed(const mxq_packed_t& w) {
  MXQ_CHECK_PTR(w.weight);
  if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b || !w.zeros_4b)
    return MXQ_E_NULL;
  if ((reinterpret_cast<uintptr_t>(w.weight_last) | reinterpret_cast<uintptr_t>(w.zeros_and_scales) |
       reinterpret_cast<uintptr_t>(w.zeros_2nd) | reinterpret_cast<uintptr_t>(w.scales_2nd) |
       reinterpret_cast<uintptr_t>(w.scales_4b) | reinterpret_cast<uintptr_t>(w.zeros_4b)) & 15)
    return MXQ_E_ALIGN;
  return MXQ_OK;
}

extern "C" int mxq_gemv_grouped_perm(const void* x, const mxq_packed_t* w, void* const* y, int n, int64_t B,
                                     int64_t IC, int64_t OC, const int32_t* group_perm, unsigned flags, void* stream) {
  if (B < 0 || IC < 0 || OC < 0 || n < 0 || n > g2::kMaxGroup) return MXQ_E_SHAPE;
  if (B == 0 || OC == 0 || n == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  if (!w || !y) return MXQ_E_NULL;
  for (int i = 0; i < n; ++i) {
    MXQ_CHECK_PTR(y[i]);
    const int rc = gemv_check_packed(w[i]);
    if (rc) return rc;
  }
  if (IC % 64 || OC % 8 || IC == 0 || IC > (1 << 24) || OC > INT32_MAX || B > 65535 * 4) return MXQ_E_SHAPE;
  // Kernel choice.  Measured on B200 (profiles/r2_gemv_chain.txt, CUDA graph over > L2 of weights, PDL):
  //   same-shape chains   4096^2: IMMA 3.9 us, ring 4.2 us;  11008x4096: 8.4 vs 8.7;  4096x11008: 9.3 vs 8.6
  //   bench.py's mixed 56-linear chain: ring 336 us, IMMA 350 us, alternating between the two 383 us
  //   (two different kernels cannot be co-resident -- 59 K + 28 K registers -- so the PDL overlap is lost)
  // A decode step is a chain of mixed shapes, so ONE kernel serves a whole process: the ring kernel by
  // default, the IMMA kernel with MXQ_GEMV_IMPL=mma (it needs IC % 256 == 0; other shapes fall back).
  const char* impl = getenv("MXQ_GEMV_IMPL");
  const bool use_mma = impl && impl[0] == 'm' && IC % 256 == 0 && IC <= 32768;
  if (!use_mma) return gemv_ring_grouped(x, w, y, n, B, IC, OC, group_perm, flags, stream);
  cudaStream_t st = as_stream(stream);
  const __half* xh = (const __half*)x;
  const bool pdl = !(flags & MXQ_GEMV_NO_PDL);
  int rc;
  if (B == 1) rc = g2::launch<1>(xh, w, y, n, (int)B, (int)IC, (int)OC, group_perm, pdl, st);
  else if (B == 2) rc = g2::launch<2>(xh, w, y, n, (int)B, (int)IC, (int)OC, group_perm, pdl, st);
  else rc = g2::launch<4>(xh, w, y, n, (int)B, (int)IC, (int)OC, group_perm, pdl, st);
  if (rc == MXQ_E_SHAPE) return gemv_ring_grouped(x, w, y, n, B, IC, OC, group_perm, flags, stream);
  return rc;
}

extern "C" int mxq_gemv_grouped(const void* x, const mxq_packed_t* w, void* const* y, int n, int64_t B,
                                int64_t IC, int64_t OC, unsigned flags, void* stream) {
  return mxq_gemv_grouped_perm(x, w, y, n, B, IC, OC, nullptr, flags, stream);
}

// profiling aid, not part of the documented surface
extern "C" __attribute__((visibility("default"))) int mxq_debug_gemv2_trace(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g2::g_trace, sizeof(unsigned long long) * 6 * 160);
}

extern "C" __attribute__((visibility("default"))) int mxq_debug_gemv2_itrace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g2::g_itrace, sizeof(long long) * 16 * 16 * 6);
}

extern "C" int mxq_gemv_ex(const void* x, mxq_packed_t w, void* y, int64_t B, int64_t IC, int64_t OC,
                           unsigned flags, void* stream) {
  void* ys[1] = {y};
  return mxq_gemv_grouped(x, &w, ys, 1, B, IC, OC, flags, stream);
}

extern "C" int mxq_gemv(const void* x, mxq_packed_t w, void* y, int64_t B, int64_t IC, int64_t OC, void* stream) {
  return mxq_gemv_ex(x, w, y, B, IC, OC, 0u, stream);
}
