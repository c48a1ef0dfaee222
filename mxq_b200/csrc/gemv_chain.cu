// Persistent decode-GEMV chain for the packed mixed 2/4-bit layout (sm_100a).
//
// Replaces a SEQUENCE of gemv_mxq_forward_cuda calls
// (mxq_quant/cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-273, one launch per linear) by ONE
// launch that walks a job list (x_j, W_j, y_j).  Measured on B200: a 6.3 MB
// 4096 x 4096 GEMV is 0.96 us of HBM time, but every dependent launch costs 2.7 us of skeleton
// (griddepcontrol.wait return, activation image, fill, epilogue) even with no copies and no arithmetic,
// so per-linear launches cannot exceed ~0.45 of the HBM rate.  Here the weight stream never stops at a
// linear boundary:
//
//   * one CTA per SM, 16 warps: 12 compute warps (3 sets of 4), 2 producers, 2 activation-image builders;
//   * a job's rows are cut into whole 16-row tiles, tbase (+1 on the first trem slices) per CTA; the heavier
//     slices rotate by trem per job so that the counts even out over a chain (equal row-group shares put a
//     full and a three-quarter tile on every CTA for 4096 rows: 13 % of the mma work on padding).  A
//     prologue turns the job list into one 16-byte record per tile of this CTA;
//   * the producers stream the CTA's tiles of ALL jobs back to back into shared-memory stages.  One stage =
//     one tile x one K chunk of <= 4096 columns of every packed tensor (25 KB) and costs 7-10 TMA operations:
//     weight / weight_last / zeros_and_scales / zeros_2nd arrive as 2-D tensor-map boxes
//     (cp.async.bulk.tensor; rows beyond the tensor and columns beyond the row are zero-filled, so every box
//     has the same byte count), scales_2nd / scales_4b / zeros_4b -- whose pieces are only 8-byte aligned in
//     the reference layout -- as 1-D bulk copies of the enclosing 16-byte aligned range (the consumer adds the
//     skew).  An earlier version issued one bulk copy per row (52 per stage): a UBLKCP costs its warp 60-100
//     cycles, the stage took 2.9 us;
//   * compute set s owns the tiles tau % 3 == s and a private ring of two stages -- every use of a slot
//     belongs to one set, which waits for all of them in order, so a parity wait can never be a phase behind
//     (a ring shared by the sets was tried first: waits on foreign stages are racy, see the kernel); the three
//     warps of a scheduler are out of phase (one reduces / waits while the others multiply).  Warp q of a set
//     owns the four 256-column units 4q .. 4q+3 of every stage of its tiles.  Thread (g, t) of the warp
//     multiplies two rows of one second-order row group with the four consecutive 64-column blocks of
//     unit 4q + t: all its metadata for the visit comes from ONE 16-byte load per tensor and row.
//     Inner products on mma.sync.m16n8k32 (IMMA.16832.U8.S8) as in csrc/gemv_mma.cu (the quad's thread
//     t supplies its own block along k, the B columns separate the four blocks again, block-floating
//     int16 activations split in signed high / low bytes), but the 2-bit codes are NOT shifted down:
//     code c of a byte is masked in place at bits 7:6 after a multiply by 4^(3-c) (IMAD on the FMA pipe
//     instead of SHF on the half-rate ALU pipe), i.e. all codes enter the mma as 64 * q, and the factor
//     2^-6 (2^-4 for the 4-bit nibbles) is folded into the activation image's group scales;
//   * the 4 K-partials of a tile meet in shared memory behind a named barrier of the set; one of the four
//     warps adds them in a fixed order (deterministic), stores y and, if a later job depends on this one,
//     counts the tile in the job's global counter with a release increment;
//   * the builders convert x_j into the block-floating image of job j+1 while job j computes (two image
//     buffers, or a ring allocation with the builders two images ahead when a chain has jobs shorter than a
//     conversion: place_images).  A job may name an earlier job `dep` whose y it reads as x (a real dependent
//     chain): the builders then wait for the counter of `dep` (acquire) -- the weight stream of the waiting
//     job is already in shared memory by then.  Launched cooperatively when a chain has dependencies (all
//     CTAs co-resident);
//   * chains without dependencies can be launched with programmatic dependent launch
//     (MXQ_GEMV_CHAIN_PDL): set-up and weight prefetch overlap the tail of the previous kernel of the
//     stream, activations are read after it has completed.
//
// Shapes: IC % 256 == 0, OC % 32 == 0, batch 1.  Other calls stay on gemv.cu / gemv_mma.cu.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mxq {
namespace g3 {

constexpr int kSets = 3;                      // compute warp sets; set s owns the tiles tau % kSets == s
constexpr int kSetW = 4;
constexpr int kCW = kSets * kSetW;            // compute warps
constexpr int kProdA = kCW, kProdB = kCW + 1, kBuilder0 = kCW + 2;
constexpr int kBW = 2;                        // builder warps
constexpr int kWarps = kCW + 2 + kBW;         // 16
constexpr int kThreads = kWarps * 32;
constexpr int kBuilderThreads = kBW * 32;
constexpr int kMaxJobs = MXQ_GEMV_CHAIN_MAX_JOBS;
constexpr uint32_t kMagic = 0x4D584334u;      // "MXC4"

// one stage = one 16-row tile x one K chunk of 64 blocks (16 units), every tensor dense
constexpr int kOffW = 0;                      // [16 rows][pw B]     weight words (pw = min(IC/4, 1024))
constexpr int kOffWL = 16384;                 // [16 rows][pwl B]    weight_last (pwl = min(IC/16, 256))
constexpr int kOffZS = 20480;                 // [16 rows][128 B]    zeros_and_scales of the metadata chunk
constexpr int kOffZ2 = 22528;                 // [4 row groups][128 B] zeros_2nd
constexpr int kOffS2 = 23040;                 // [4 row groups][pitch] scales_2nd (+ skew)
constexpr int kPS2 = 416;                     // pitch of the per-row-group copies: 64 blocks x 6 B + 8 skew + pad
constexpr int kOffS4 = kOffS2 + 4 * kPS2;     // 24704: 48 B around the tile's 16 fp16 scales_4b
constexpr int kOffZ4 = kOffS4 + 48;           // 24752: 32 B around the tile's zeros_4b words
constexpr int kStageBytes = 24832;
static_assert(kOffZ4 + 32 <= kStageBytes && kStageBytes % 128 == 0, "stage layout");
constexpr uint32_t kBytesBox = 2048 + 512;    // producer B: zeros_and_scales + zeros_2nd boxes

constexpr int kRedBytes = kSets * 2 * kSetW * 16 * 4;   // per set: two parities x 4 warps x 16 rows
constexpr int kMaxStages = 8;
#ifndef MXQ_CHAIN_BACKOFF
#define MXQ_CHAIN_BACKOFF 0
#endif
constexpr bool kWaitBackoff = MXQ_CHAIN_BACKOFF != 0;
#ifndef MXQ_CHAIN_XPREFETCH
#define MXQ_CHAIN_XPREFETCH 1
#endif
constexpr int kMaxTiles = 256;                // 16-row tiles of one CTA over the whole chain
constexpr size_t kSmemMax = 227 * 1024 - 1024;

struct JobD {                                 // 128 bytes
  const unsigned char* S2;
  const unsigned char* S4;
  const unsigned char* Z4;
  const __half* x;
  __half* y;
  int nblk, nch, nqb, ngrp;
  int tbase, trem, rot, dep;                   // 16-row tiles per CTA: tbase, +1 on the first trem slices
  int share, dep_target, publish, s2pitch;    // s2pitch: 0 = one copy per row group (pitch kPS2, skewed)
  int oc, ximg_blocks, pw, pwl;               // pw / pwl: row pitch of the weight / weight_last boxes
  int imgk, imgoff, imgwait;                  // activation image: sequence number, byte offset in the image pool,
  int pad[3];                                 // image that has to be released before this one is built (-1: none)
};
static_assert(sizeof(JobD) == 128, "JobD");

struct PlanH {                                // 128 bytes
  uint32_t magic;
  int n, nstages, ximg_max;
  int ncta, coop, smem, dbg;                  // dbg: profiling only (MXQ_CHAIN_DBG): 1 no arithmetic, 2 no copies, 4 no image, 8 trace
  int pad[24];
};
static_assert(sizeof(PlanH) == 128, "PlanH");

struct ChainParams {                          // passed by value (__grid_constant__): every job field is a
  PlanH H;                                    // constant-bank operand
  JobD jobs[kMaxJobs];
};
static_assert(sizeof(ChainParams) <= 32000, "kernel parameter space");

// The plan blob: ChainParams, then 4 tensor maps per job (weight, weight_last, zeros_and_scales,
// zeros_2nd).  The caller keeps a device copy of the blob for the maps (TMA descriptors are read from
// global memory).
constexpr size_t kMapsOffset = (sizeof(ChainParams) + 127) & ~size_t(127);
constexpr size_t kPlanBytes = kMapsOffset + (size_t)kMaxJobs * 4 * sizeof(CUtensorMap);

constexpr int kImgBars = 8;                   // activation images in flight (barrier k % 8 belongs to image k)

struct Bars {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t imgfull[kImgBars], imgempty[kImgBars];
};

// Blocking wait with a suspend-time hint: the warp sleeps in hardware until the phase completes instead of
// spinning.  With 16 warps of which most are waiting at any time, hint-less try_wait loops took three of
// four issue slots from the working warp of a scheduler (every traced step ran ~4x slower than alone).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  long long t0 = 0;
  uint32_t ns = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)          // up to 1 ms per attempt
        : "memory");
    if (!ok) {
      // a barrier that does not complete for ~10 s is a missing arrive / copy, not a slow peer: trap instead
      // of wedging the GPU (time-slicing or a debugger pause cannot reach that on a healthy launch)
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000LL) __trap();
      if (kWaitBackoff) {                                        // back off: a polling warp takes issue slots
        ns = ns < 128u ? ns + 32u : 128u;
        __nanosleep(ns);
      }
    }
  } while (!ok);
}
#define mbar_wait mbar_wait_sleep

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// release at gpu scope: the y stores of the whole warp (ordered before by __syncwarp) become visible first
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg16(const void* p) {   // L2 only: x may have been written by this launch
  uint4 v;
  asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tma_box(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                     uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
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
// lo(a) * h + c with fp16 operands and an fp32 accumulator (the product is exact)
__device__ __forceinline__ float fhfma(uint32_t a2, unsigned short h, float acc) {
  unsigned short a;
  asm("{.reg .b16 t; mov.b32 {%0,t}, %1;}" : "=h"(a) : "r"(a2));
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(h));
  return acc;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Profiling only (built with -DMXQ_CHAIN_TRACE, run with MXQ_CHAIN_DBG & 8): clock64 stamps of CTA 1 --
// [role][event index][stamp]; roles: compute warp 0, producer A, (globaltimer at entry / exit of EVERY CTA,
// two per CTA: profiles/r2_gemv_chain_spread.py), builder warp 0, compute warp 0 per job.
// Compiled out by default: the predicated stamps sat in the per-tile paths of every warp.
constexpr int kTraceEvents = 96;
#ifdef MXQ_CHAIN_TRACE
__device__ long long g_ctrace[5 * kTraceEvents * 4];
#define CTRACE(role, idx, k)                                                                       \
  do {                                                                                             \
    if (tr && (idx) < kTraceEvents) g_ctrace[((role) * kTraceEvents + (idx)) * 4 + (k)] = clock64(); \
  } while (0)
#define CTRACE_CTA(k)                                                                              \
  do {                                                                                             \
    if ((dbg & 8) && threadIdx.x == 0 && cta * 2 + (k) < kTraceEvents * 4) {                       \
      long long t_;                                                                                \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                       \
      g_ctrace[2 * kTraceEvents * 4 + cta * 2 + (k)] = t_;                                         \
    }                                                                                              \
  } while (0)
#else
#define CTRACE_CTA(k) \
  do {                \
  } while (0)
#define CTRACE(role, idx, k) \
  do {                       \
    (void)tr;                \
  } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// Activation image of one job (one batch row):
//   per quad of units Q = block / 16 (the four units of one compute-warp visit), 2048 B:
//     [i = block % 4][k = group][t = unit % 4][h][16 B]   signed high (h = 0) / low (h = 1) bytes of the group
//                   (2-bit groups: register c, byte j = element 4j + c; pooled group: register
//                   (e>>3)*2 + (e&1), byte (e&7)>>1 = element e) -- the 8 lanes that feed one mma's B operand
//                   (4 thread-columns t x high / low) read 128 contiguous bytes
//   tabI int4[nb], tabF float4[nb]  indexed [Q][i][t]:
//     tabI  -(sum_j X_j) * {64, 16, 4, 16}: the zero-point term for z1 masked IN PLACE
//           (zs & (3 << 2k) = z1 * 4^k) and for the nibble zero z4 (codes enter as 16 * q)
//     tabF  2^(E-14) * {2^-6, 2^-6, 2^-6, 2^-4}
// nb = nch * 64 blocks (zeros beyond the row); block = 16 Q + 4 t + i.
// X_j = rint(x_j * 2^(14-E)), E = exponent of 1.0078 * max|x| of the group, X = 256 * hi + lo with both
// bytes signed.  Same quantities as csrc/gemv_mma.cu (error bound measured in tests/test_gpu_packed.py).
// ---------------------------------------------------------------------------------------------
template <int k>
__device__ __forceinline__ void stage_group(const uint4 v0, const uint4 v1, unsigned char* dst, int* tI, float* tF) {
  // dst: the group's high bytes; the low bytes follow 16 B later
  const uint32_t xw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t m2 = xw[0] & 0x7FFF7FFFu;
#pragma unroll
  for (int i = 1; i < 8; ++i) m2 = P16<__half>::vmax(m2, xw[i] & 0x7FFF7FFFu);
  float gmax = fmaxf(P16<__half>::lo(m2), P16<__half>::hi(m2));
  gmax = fminf(gmax, 65504.f) * 1.0078125f;
  const uint32_t eb = __float_as_uint(gmax) >> 23;          // biased exponent (>= 103 when gmax > 0)
  const float up = gmax > 0.f ? __uint_as_float((268u - eb) << 23) : 0.f;
  const float xsc = gmax > 0.f ? __uint_as_float((eb - (k < 3 ? 20u : 18u)) << 23) : 0.f;
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
  constexpr int sh = k == 0 ? 6 : (k == 2 ? 2 : 4);
  *tI = (int)((0u - xs) << sh);
  *tF = xsc;
}

// which rows of job J this CTA owns
struct Share {
  bool active;
  int grp_base, qc, T;
};
__device__ __forceinline__ Share cta_share(const JobD& J, int cta, int ncta) {
  Share s;
  int slice = cta + J.rot;
  if (slice >= ncta) slice -= ncta;
  // whole 16-row tiles only (OC % 32 == 0): slice s owns tbase (+1 if s < trem) consecutive tiles.  Cutting the
  // rows into equal ROW-GROUP shares instead (7 row groups = a full and a three-quarter tile for 4096 rows on
  // 148 CTAs) spent 13 % of the mma work on padding rows; the uneven tile counts even out over the jobs of
  // a chain because `rot` moves the heavier slices on by trem every job.
  const int cnt = J.tbase + (slice < J.trem ? 1 : 0);
  const int start = slice * J.tbase + min(slice, J.trem);
  s.active = cnt > 0;
  s.grp_base = start * 4;
  s.qc = cnt * 4;
  s.T = (s.qc + 3) >> 2;
  return s;
}

// one 64-column block of this thread's two rows: three 2-bit groups + the pooled 4-bit group
//   wa / wb   the block's 4 weight words of row a / row b,  wla / wlb weight_last
//   zsa / zsb the block's 16-bit zeros_and_scales field of row a / b (junk above bit 15 allowed),
//   z2        the block's zeros_2nd byte (junk above bit 7 allowed), s2[3] its second-order scales
__device__ __forceinline__ void block_mac(const uint4 wa, const uint4 wb, const uint32_t wla, const uint32_t wlb,
                                          const uint32_t zsa, const uint32_t zsb, const uint32_t z2,
                                          const unsigned short (&s2)[3], const unsigned char* __restrict__ xb,
                                          const bool bact, const int4 tI, const float4 tF, const float s4a,
                                          const float s4b, const int z4a, const int z4b, const int c4, const int c16,
                                          const int c64, const int c256, float& acc0, float& acc1) {
  constexpr uint32_t M2 = 0xC0C0C0C0u, M4 = 0xF0F0F0F0u;
  const uint32_t zsha = zsa >> 4, zshb = zsb >> 4;
  const uint32_t z2s8 = (uint32_t)imad((int)z2, c256, 0), z2s4 = (uint32_t)imad((int)z2, c16, 0);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    uint4 X = make_uint4(0u, 0u, 0u, 0u);                               // only 8 lanes feed a non-zero B column
    if (bact) X = *reinterpret_cast<const uint4*>(xb + k * 128);
    const uint32_t wka = k == 0 ? wa.x : (k == 1 ? wa.y : wa.z);
    const uint32_t wkb = k == 0 ? wb.x : (k == 1 ? wb.y : wb.z);
    int d[4] = {0, 0, 0, 0};
    // codes c = 3, 2 (elements 4j + c of byte j), then c = 1, 0: all as 64 * q
    imma(d, wka & M2, wkb & M2, (uint32_t)imad((int)wka, c4, 0) & M2, (uint32_t)imad((int)wkb, c4, 0) & M2, X.w, X.z);
    imma(d, (uint32_t)imad((int)wka, c16, 0) & M2, (uint32_t)imad((int)wkb, c16, 0) & M2,
         (uint32_t)imad((int)wka, c64, 0) & M2, (uint32_t)imad((int)wkb, c64, 0) & M2, X.y, X.x);
    // S = s2 * (c - z2) = (cb + c) * s2 - (cb + z2) * s2, cb = 4 (k = 0, 2) or 16 (k = 1); both products are
    // exact in fp32 (gemv_mxq_cuda.cu:136)
    const uint32_t hz = k == 0 ? lop3_and_or(z2s8, 0x0300u, 0xC400u)
                      : k == 1 ? lop3_and_or(z2s4, 0x00C0u, 0xCC00u)
                               : lop3_and_or(z2s4, 0x0300u, 0xC400u);
    const float S0 = fhfma(hz, s2[k], 0.f);                              // -(cb + z2) * s2
    const int nxs = k == 0 ? tI.x : (k == 1 ? tI.y : tI.z);
    const float xsc = k == 0 ? tF.x : (k == 1 ? tF.y : tF.z);
    {
      const uint32_t hc = k == 0 ? lop3_and_or(zsa, 0x0300u, 0x4400u)
                        : k == 1 ? lop3_and_or(zsha, 0x00C0u, 0x4C00u)
                                 : lop3_and_or(zsha, 0x0300u, 0x4400u);
      const float Ssc = fhfma(hc, s2[k], S0) * xsc;
      const int z1m = (int)(zsa & (3u << (2 * k)));                      // z1 * 4^k
      const int dd = imad(z1m, nxs, imad(d[0], 256, d[1]));
      acc0 = fmaf(Ssc, (float)dd, acc0);                                 // :153
    }
    {
      const uint32_t hc = k == 0 ? lop3_and_or(zsb, 0x0300u, 0x4400u)
                        : k == 1 ? lop3_and_or(zshb, 0x00C0u, 0x4C00u)
                                 : lop3_and_or(zshb, 0x0300u, 0x4400u);
      const float Ssc = fhfma(hc, s2[k], S0) * xsc;
      const int z1m = (int)(zsb & (3u << (2 * k)));
      const int dd = imad(z1m, nxs, imad(d[2], 256, d[3]));
      acc1 = fmaf(Ssc, (float)dd, acc1);
    }
  }
  {
    uint4 X = make_uint4(0u, 0u, 0u, 0u);
    if (bact) X = *reinterpret_cast<const uint4*>(xb + 3 * 128);
    int d[4] = {0, 0, 0, 0};
    // nibbles as 16 * q: low nibbles (elements 0,2,4,6 / 8,..) shifted up, high nibbles in place
    imma(d, (uint32_t)imad((int)wa.w, c16, 0) & M4, (uint32_t)imad((int)wb.w, c16, 0) & M4, wa.w & M4, wb.w & M4, X.x, X.y);
    imma(d, (uint32_t)imad((int)wla, c16, 0) & M4, (uint32_t)imad((int)wlb, c16, 0) & M4, wla & M4, wlb & M4, X.z, X.w);
    const int dda = imad(z4a, tI.w, imad(d[0], 256, d[1]));
    const int ddb = imad(z4b, tI.w, imad(d[2], 256, d[3]));
    acc0 = fmaf(s4a * tF.w, (float)dda, acc0);                           // :179,192
    acc1 = fmaf(s4b * tF.w, (float)ddb, acc1);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemv_chain_kernel(const __grid_constant__ ChainParams P, const CUtensorMap* __restrict__ maps, int* __restrict__ sync_ws,
                  const int c4, const int c16, const int c64, const int c256) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int n = P.H.n, S = P.H.nstages, ncta = P.H.ncta, ximg_max = P.H.ximg_max, dbg = P.H.dbg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cta = blockIdx.x;

  CTRACE_CTA(0);
  // Programmatic dependent launch (mxq_gemv_chain_run flag MXQ_GEMV_CHAIN_PDL): the NEXT kernel of the stream may
  // start as soon as every CTA of this one has got here -- its CTAs take over an SM the moment this kernel's CTA
  // leaves it (the CTAs of a chain leave up to 10 us apart), build their tile lists and fill their stage rings
  // with weights, and wait for this grid only where they first read an activation vector (the builders below).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  unsigned char* stages = smem;
  unsigned char* img = stages + (size_t)S * kStageBytes;
  float* red = reinterpret_cast<float*>(img + 2 * (size_t)ximg_max);
  Bars& bars = *reinterpret_cast<Bars*>(reinterpret_cast<unsigned char*>(red) + kRedBytes);
  // job table for the compute / reducer / builder warps: a constant-bank miss per job (a new 128-byte
  // descriptor every ~2 us) showed up as a 1300-cycle gap between two jobs of a compute warp; the producers,
  // which run ahead, keep reading the constant bank (uniform operands for the TMA instructions)
  JobD* jobs = reinterpret_cast<JobD*>(reinterpret_cast<unsigned char*>(&bars) + ((sizeof(Bars) + 127) & ~size_t(127)));
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.jobs[0]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(jobs);
    for (int i = threadIdx.x; i < n * (int)(sizeof(JobD) / 4); i += kThreads) dst[i] = src[i];
  }
  // This CTA's tiles in chain order.  Pass 1, a thread per job: tiles[tau] = {job | nrg << 16, first row
  // group, K chunks | first stage-use << 8, image sequence number}.  Pass 2, a thread per tile: the record a
  // compute warp of set tau % kSets reads for its tile -- what changed since the set's PREVIOUS tile
  // (tau - kSets) is folded in, so a warp touches one 16-byte record per tile it owns and none for the others
  // (walking every tile and every job cost a compute warp ~350 instructions per owned tile, half a visit):
  //   x = job | nrg << 16 | job changed << 24 | images to step through << 25
  //   y = first row group,  z = K chunks | stages to skip before the tile << 8,  w = image sequence number
  uint4* tiles = reinterpret_cast<uint4*>(jobs + n);
  uint4* recs = tiles + kMaxTiles;
  int* tcount = reinterpret_cast<int*>(recs + kMaxTiles);    // [kMaxJobs] tiles per job, [1] total, [1] images
  __syncthreads();                                           // the shared-memory job table is complete
  if (threadIdx.x < n) {
    const JobD& J = jobs[threadIdx.x];
    const Share sh = cta_share(J, cta, ncta);
    tcount[threadIdx.x] = sh.active ? sh.T : 0;
  }
  __syncthreads();
  if (threadIdx.x < n) {
    const int j = threadIdx.x;
    int off = 0, use0 = 0;
    for (int k = 0; k < j; ++k) {
      off += tcount[k];
      use0 += tcount[k] * jobs[k].nch;
    }
    const JobD& J = jobs[j];
    const int imgk = J.imgk;
    const Share sh = cta_share(J, cta, ncta);
    const int T = tcount[j];
    for (int tile = 0; tile < T; ++tile) {
      const int nrg = min(4, sh.qc - tile * 4);
      tiles[off + tile] = make_uint4((uint32_t)j | ((uint32_t)nrg << 16), (uint32_t)(sh.grp_base + tile * 4),
                                     (uint32_t)J.nch | ((uint32_t)(use0 + tile * J.nch) << 8), (uint32_t)imgk);
    }
    if (j == n - 1) {
      tcount[kMaxJobs] = off + T;
      tcount[kMaxJobs + 1] = imgk + 1;
    }
  }
  __syncthreads();
  for (int tau = threadIdx.x; tau < tcount[kMaxJobs]; tau += kThreads) {
    const uint4 E = tiles[tau];
    uint32_t jobchg = 1, skip = E.z >> 8, steps = E.w + 1;
    if (tau >= kSets) {
      const uint4 Pv = tiles[tau - kSets];
      jobchg = (Pv.x & 0xFFu) != (E.x & 0xFFu);
      skip = (E.z >> 8) - ((Pv.z >> 8) + (Pv.z & 0xFFu));
      steps = E.w - Pv.w;
    }
    recs[tau] = make_uint4(E.x | (jobchg << 24) | (steps << 25), E.y, (E.z & 0xFFu) | (skip << 8), E.w);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bars.full[s], 2); mbar_init(&bars.empty[s], kSetW); }
    for (int b = 0; b < kImgBars; ++b) {
      mbar_init(&bars.imgfull[b], kBW);
      mbar_init(&bars.imgempty[b], kCW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp < kCW) {
    // =========================================== compute ===========================================
    const int g = lane >> 2, t = lane & 3;
    const int set = warp >> 2, qw = warp & 3;
    int nset = 0;                                             // tiles this set has finished
    const int U = qw * 4 + t;                                 // this thread-column's unit inside a stage
    const int rgl = g >> 1;                                   // row group of this thread's two rows
    const int rowa = rgl * 4 + (g & 1), rowb = rowa + 2;      // tile rows (mma rows g and g + 8)
    const uint32_t half = (uint32_t)(U >> 3);                 // blocks 32.. of a metadata chunk: high half / byte 1
    const uint32_t oZSa = kOffZS + rowa * 128 + (U & 7) * 16, oZSb = oZSa + 2 * 128;
    const uint32_t oZ2 = kOffZ2 + rgl * 128 + (U & 7) * 16;
    const uint32_t zsh = half * 16, z2sh = half * 8;
    const bool bact = (g >> 1) == t;                          // this lane feeds a non-zero B column
    const uint32_t xlane = (uint32_t)(qw * 2048 + t * 32 + (g & 1) * 16);   // + i * 512 + k * 128 (bact lanes only)
    // Thread-columns 2, 3 walk their four blocks in the order 2, 3, 0, 1: the weight loads of one instruction
    // then hit four different 16-byte bank groups of a row instead of two (the units of t and t + 2 are
    // 128 B apart).  sw selects the swapped halves of every per-visit metadata vector.
    const bool sw = (t & 2) != 0;
    const uint32_t lo32 = sw ? 32u : 0u, hi32 = sw ? 0u : 32u;

    // Every set has its OWN ring of D stages (slots set * D ..): all uses of a slot belong to one set, which
    // waits for every one of them in order -- a parity wait can never be a phase behind.  (A ring shared by
    // the sets needs waits on stages a set does not own; the producers may refill such a stage before a busy
    // set has looked, and the parity then answers for the wrong phase: wrong results, later a hang.)
    const int D = S >= 2 * kSets ? 2 : 1;
    int kown = 0;                                             // stages this set has consumed
    const bool tr = (dbg & 8) && cta == 1 && warp == 0 && lane == 0;
    int ev = 0;
    // the loop bound is re-read from shared memory once per tile: kept in a register it was spilled (128 registers
    // are live in the loop body), and the local-memory reload -- an L2 round trip beside 227 KB of shared memory --
    // held every compute warp ~5 % of its time at the loop's compare (ncu source view, stall_long_sb)
    const volatile int* counts = tcount + kMaxJobs;          // [0] tiles of this CTA, [1] images
    // per-job state, reloaded when the job changes between two tiles of this set
    const unsigned char* ximg = img;
    const unsigned char* tabI = img;
    const unsigned char* tabF = img;
    __half* yj = nullptr;
    int nqb = 0, publish = 0, lastimg = -1, nbcur = 0;
    uint32_t s2pitch = kPS2, s2odd = 0, oWa = 0, oWb = 0, oWLa = 0, oWLb = 0;
    // every compute warp passes through EVERY image in order (release the previous one, observe the next
    // one's barrier phase) whether its set uses it or not: the builders recycle a buffer only when all 12
    // warps have released it, and a parity wait must not skip a phase
    auto step_images = [&](int upto) {
      while (lastimg < upto) {
        if (lastimg >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.imgempty[lastimg & (kImgBars - 1)]);
        }
        ++lastimg;
        mbar_wait(&bars.imgfull[lastimg & (kImgBars - 1)], (lastimg >> 3) & 1);
      }
    };
    for (int e = set; e < counts[0]; e += kSets) {
      const uint4 E = recs[e];
      const int j = (int)(E.x & 0xFFu), nrg = (int)((E.x >> 16) & 0xFFu), nch = (int)(E.z & 0xFFu);
      const int rg0 = (int)E.y;
      if (E.x >> 25) step_images((int)E.w);                   // new activation image(s) since this set's last tile
      if (E.x & (1u << 24)) {                                 // another job than this set's last tile
        const JobD& J = jobs[j];
        ximg = img + J.imgoff;
        nbcur = J.ximg_blocks;
        nqb = J.nqb;
        publish = J.publish;
        yj = J.y;
        s2pitch = J.s2pitch ? (uint32_t)J.s2pitch : (uint32_t)kPS2;
        s2odd = J.s2pitch ? 0u : (uint32_t)((J.nblk * 6) & 15);
        oWa = kOffW + rowa * J.pw + U * 64;
        oWb = oWa + 2 * J.pw;
        oWLa = kOffWL + rowa * J.pwl + U * 16;
        oWLb = oWLa + 2 * J.pwl;
      }
      if (E.x & (0xFFu << 24)) {                              // the tables follow the image: job or image changed
        tabI = ximg + nbcur * 128;
        tabF = tabI + nbcur * 16;
      }
      {
        float acc0 = 0.f, acc1 = 0.f;
        float s4a = 0.f, s4b = 0.f;
        int z4a = 0, z4b = 0;
        const uint32_t oS2 = kOffS2 + rgl * s2pitch + (((rg0 + rgl) & 1) ? s2odd : 0u) + U * 24;
        for (int ch = 0; ch < nch; ++ch) {
          const int nq = min(16, nqb - ch * 16);
          const int slot = set * D + (kown & (D - 1));
          const uint32_t sphase = (uint32_t)(kown >> (D - 1)) & 1u;
          ++kown;
          CTRACE(0, ev, 0);
          mbar_wait(&bars.full[slot], sphase);
          CTRACE(0, ev, 1);
          const unsigned char* s = stages + (size_t)slot * kStageBytes;
          if (ch == 0) {                                      // 4-bit pool scale / zero of this thread's two rows
            const int R = rg0 * 4;
            const __half* s4p = reinterpret_cast<const __half*>(s + kOffS4 + ((rg0 & 1) ? 8 : 0));
            s4a = __half2float(s4p[rowa]);
            s4b = __half2float(s4p[rowb]);
            const uint32_t* z4w = reinterpret_cast<const uint32_t*>(s + kOffZ4);
            const int w0 = (R >> 3) & ~3;                     // first word of the 16-byte aligned copy
            const int Ra = R + rowa, Rb = R + rowb;
            z4a = (int)((z4w[(Ra >> 3) - w0] >> (4 * (Ra & 7))) & 0xF);
            z4b = (int)((z4w[(Rb >> 3) - w0] >> (4 * (Rb & 7))) & 0xF);
          }
          if (qw * 4 < nq && !(dbg & 1)) {
            // metadata of the thread-column's four blocks: one 16-byte load per tensor and row
            const uint4 wl4a = *reinterpret_cast<const uint4*>(s + oWLa);
            const uint4 wl4b = *reinterpret_cast<const uint4*>(s + oWLb);
            const uint4 zs4a = *reinterpret_cast<const uint4*>(s + oZSa);
            const uint4 zs4b = *reinterpret_cast<const uint4*>(s + oZSb);
            const uint4 z24 = *reinterpret_cast<const uint4*>(s + oZ2);
            const uint2 sA = *reinterpret_cast<const uint2*>(s + oS2);
            const uint2 sB = *reinterpret_cast<const uint2*>(s + oS2 + 8);
            const uint2 sC = *reinterpret_cast<const uint2*>(s + oS2 + 16);
            // blocks in processing order: (0, 1, 2, 3) or (2, 3, 0, 1)
            const uint32_t wlA[4] = {sw ? wl4a.z : wl4a.x, sw ? wl4a.w : wl4a.y, sw ? wl4a.x : wl4a.z, sw ? wl4a.y : wl4a.w};
            const uint32_t wlB[4] = {sw ? wl4b.z : wl4b.x, sw ? wl4b.w : wl4b.y, sw ? wl4b.x : wl4b.z, sw ? wl4b.y : wl4b.w};
            const uint32_t zsA[4] = {sw ? zs4a.z : zs4a.x, sw ? zs4a.w : zs4a.y, sw ? zs4a.x : zs4a.z, sw ? zs4a.y : zs4a.w};
            const uint32_t zsB[4] = {sw ? zs4b.z : zs4b.x, sw ? zs4b.w : zs4b.y, sw ? zs4b.x : zs4b.z, sw ? zs4b.y : zs4b.w};
            const uint32_t z2v[4] = {sw ? z24.z : z24.x, sw ? z24.w : z24.y, sw ? z24.x : z24.z, sw ? z24.y : z24.w};
            const uint32_t s2w[6] = {sw ? sB.y : sA.x, sw ? sC.x : sA.y, sw ? sC.y : sB.x,
                                     sw ? sA.x : sB.y, sw ? sA.y : sC.x, sw ? sB.x : sC.y};
            const unsigned char* xb0 = ximg + xlane + (uint32_t)ch * 8192u;
            const unsigned char* tI0 = tabI + (ch * 64 + qw * 16 + t) * 16;      // + block * 64
            const unsigned char* tF0 = tabF + (ch * 64 + qw * 16 + t) * 16;
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // i-th block in processing order = block (i ^ 2) for the swapped thread-columns
              const uint32_t o32 = (i & 2) ? hi32 : lo32;                        // 32 * (block >> 1)
              const uint4 wa = *reinterpret_cast<const uint4*>(s + oWa + o32 + (i & 1) * 16);
              const uint4 wb = *reinterpret_cast<const uint4*>(s + oWb + o32 + (i & 1) * 16);
              const uint32_t wla = wlA[i], wlb = wlB[i];
              const uint32_t zsa = zsA[i] >> zsh, zsb = zsB[i] >> zsh;
              const uint32_t z2 = z2v[i] >> z2sh;
              unsigned short s2[3];
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const int hidx = 3 * i + k;
                s2[k] = (unsigned short)(s2w[hidx >> 1] >> (16 * (hidx & 1)));
              }
              const int4 tI = *reinterpret_cast<const int4*>(tI0 + o32 * 4 + (i & 1) * 64);
              const float4 tF = *reinterpret_cast<const float4*>(tF0 + o32 * 4 + (i & 1) * 64);
              block_mac(wa, wb, wla, wlb, zsa, zsb, z2, s2, xb0 + o32 * 32 + (i & 1) * 512, bact, tI, tF, s4a, s4b, z4a, z4b,
                        c4, c16, c64, c256, v0, v1);
            }
            if (U < nq) { acc0 += v0; acc1 += v1; }           // units beyond the row: zero codes, but stale scales
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.empty[slot]);
          CTRACE(0, ev, 2);
          if (ch + 1 < nch) { CTRACE(0, ev, 3); ++ev; }
        }
        // ---- K partials of the set's 4 warps -> one of them adds in a fixed order and stores y -------
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
        float* rbuf = red + ((size_t)(set * 2 + (nset & 1)) * kSetW) * 16;
        if (t == 0) {
          rbuf[qw * 16 + rowa] = acc0;
          rbuf[qw * 16 + rowb] = acc1;
        }
        // The buffer of parity p is rewritten two tiles later, after another pass through this barrier,
        // which the warp that adds tile n only reaches after it has read the buffer.
        asm volatile("bar.sync %0, %1;" ::"r"(1 + set), "n"(kSetW * 32) : "memory");
        if ((nset & 3) == qw) {
          const int r = lane & 15, h = lane >> 4;
          float sum = rbuf[(h * 2) * 16 + r] + rbuf[(h * 2 + 1) * 16 + r];
          sum += __shfl_xor_sync(0xffffffffu, sum, 16);      // (warp 0 + warp 1) + (warp 2 + warp 3) on both halves
          if (lane < 16 && (r >> 2) < nrg) yj[(size_t)rg0 * 4 + r] = __float2half_rn(sum);
          if (publish) {                                      // a later job waits for this one: count the tile
            __syncwarp();
            if (lane == 0) red_release_add(sync_ws + j, 1);
          }
        }
        ++nset;
        CTRACE(0, ev, 3);
        ++ev;
      }
    }
    step_images(counts[1] - 1);                               // images that only other sets used
  } else if (warp == kProdA || warp == kProdB) {
    // =========================================== producers =========================================
    // A: the four tensor-map boxes (weight, weight_last, zeros_and_scales, zeros_2nd).  B: the 1-D bulk copies with
    // their address arithmetic (scales_2nd, scales_4b, zeros_4b).
    // Issue order: the tiles of a wave (kSets consecutive tiles, one per set) are filled chunk by chunk in
    // turn, so a tile of several K chunks -- which only fits its set's ring two chunks at a time -- does not
    // keep the in-order producer from prefetching the other sets' tiles.
    const bool isA = warp == kProdA;
    const int D = S >= 2 * kSets ? 2 : 1;
    const int ntile = tcount[kMaxJobs];
    int kset[kSets];                                          // stages issued per set (unrolled: registers)
#pragma unroll
    for (int q = 0; q < kSets; ++q) kset[q] = 0;
    const bool tr = (dbg & 8) && cta == 1 && lane == 0 && isA;
    int ev = 0;
    for (int w0 = 0; w0 < ntile; w0 += kSets) {
      // the wave's tiles and their jobs, read once (registers: the loops over q are unrolled)
      int nchq[kSets], nrgq[kSets], rg0q[kSets], nblkq[kSets], nqbq[kSets], s2pq[kSets], ocq[kSets];
      uint32_t bytesAq[kSets];
      const unsigned char* S2q[kSets];
      const unsigned char* S4q[kSets];
      const unsigned char* Z4q[kSets];
      const CUtensorMap* mpq[kSets];
      int maxch = 0;
#pragma unroll
      for (int q = 0; q < kSets; ++q) {
        nchq[q] = 0;
        if (w0 + q < ntile) {
          const uint4 E = tiles[w0 + q];
          const int j = (int)(E.x & 0xFFu);
          const JobD& J = jobs[j];
          nchq[q] = (int)(E.z & 0xFFu);
          nrgq[q] = (int)((E.x >> 16) & 0xFFu);
          rg0q[q] = (int)E.y;
          nblkq[q] = J.nblk;
          nqbq[q] = J.nqb;
          s2pq[q] = J.s2pitch;
          ocq[q] = J.oc;
          bytesAq[q] = 16u * (uint32_t)(J.pw + J.pwl);
          S2q[q] = J.S2;
          S4q[q] = J.S4;
          Z4q[q] = J.Z4;
          mpq[q] = maps + (size_t)j * 4;
          maxch = max(maxch, nchq[q]);
        }
      }
      for (int ch = 0; ch < maxch; ++ch) {
#pragma unroll
        for (int q = 0; q < kSets; ++q) {
          if (ch >= nchq[q]) continue;
          const int nrg = nrgq[q], rg0 = rg0q[q], row0 = rg0 * 4;
          const int nblk = nblkq[q], nq = min(16, nqbq[q] - ch * 16);
          const CUtensorMap* mp = mpq[q];
          const int slot = q * D + (kset[q] & (D - 1));
          const uint32_t sphase = (uint32_t)(kset[q] >> (D - 1)) & 1u;
          ++kset[q];
          CTRACE(1, ev, 0);
          mbar_wait(&bars.empty[slot], sphase ^ 1);
          CTRACE(1, ev, 1);
          unsigned char* s = stages + (size_t)slot * kStageBytes;
          uint64_t* fb = &bars.full[slot];
          if (elect_one()) {
            if (dbg & 2) {
              mbar_arrive(fb);
            } else if (isA) {
              mbar_arrive_expect_tx(fb, bytesAq[q] + kBytesBox);
              tma_box(s + kOffW, mp + 0, ch * 256, row0, fb);
              tma_box(s + kOffWL, mp + 1, ch * 64, row0, fb);
              tma_box(s + kOffZS, mp + 2, ch * 32, row0, fb);
              tma_box(s + kOffZ2, mp + 3, ch * 32, rg0, fb);
            } else {
              // scales_2nd: row group r's piece starts at (rg0 + r) * nblk * 6 + ch * 384 (8-byte aligned)
              uint32_t s2bytes;
              const uint32_t piece = (uint32_t)nq * 24;
              const uint32_t odd = (uint32_t)((nblk * 6) & 15);      // 0 or 8
              if (s2pq[q]) {
                s2bytes = (uint32_t)nrg * (uint32_t)s2pq[q];
              } else {
                s2bytes = 0;
                for (int r = 0; r < nrg; ++r) s2bytes += (piece + (((rg0 + r) & 1) ? odd : 0u) + 15u) & ~15u;
              }
              // scales_4b / zeros_4b: the 16-byte aligned range around the tile's entries, clipped to the tensor
              const uint32_t s4off = ((uint32_t)rg0 * 8u) & ~15u;
              const uint32_t s4bytes = min(48u, (uint32_t)ocq[q] * 2u - s4off);
              const uint32_t z4off = (uint32_t)((row0 >> 3) & ~3) * 4u;
              const uint32_t z4bytes = min(32u, (uint32_t)(ocq[q] >> 3) * 4u - z4off);
              mbar_arrive_expect_tx(fb, s2bytes + s4bytes + z4bytes);
              if (s2pq[q]) {
                bulk_g2s(s + kOffS2, S2q[q] + (size_t)rg0 * (size_t)s2pq[q], s2bytes, fb);
              } else {
                for (int r = 0; r < nrg; ++r) {
                  const size_t off = (size_t)(rg0 + r) * ((size_t)nblk * 6) + (size_t)ch * 384;
                  const uint32_t sk = ((rg0 + r) & 1) ? odd : 0u;
                  bulk_g2s(s + kOffS2 + r * kPS2, S2q[q] + (off - sk), (piece + sk + 15u) & ~15u, fb);
                }
              }
              bulk_g2s(s + kOffS4, S4q[q] + s4off, s4bytes, fb);
              bulk_g2s(s + kOffZ4, Z4q[q] + z4off, z4bytes, fb);
            }
          }
          __syncwarp();
          CTRACE(1, ev, 3);
          ++ev;
        }
      }
    }
  } else {
    // =========================================== image builders ====================================
    // a thread converts one 64-column block (4 groups) per round
    // Images live in a pool of 2 * ximg_max bytes; the plan gives every image its offset (ring allocation, an
    // image is never split) and the last earlier image whose bytes or barrier it reuses.  With the 30 KB
    // pool half a 11008-column vector needs, three 4096-column images fit beside it: the builders run up to
    // three images ahead.  (Two fixed buffers let them start the gate/up image only when the last set had
    // left the q/k/v image, i.e. during the 1-2 tiles of o_proj: 1 us per layer on the 56-linear chain.)
    // Every CTA builds every image, also those of jobs it has no tile of (numbering is global).
    const int bt = (warp - kBuilder0) * 32 + lane;
    const bool tr = (dbg & 8) && cta == 1 && bt == 0;
    // Everything before this point read kernel parameters, the plan and packed weights only.  The activations
    // (and the counters, the outputs) may belong to the previous kernel of the stream: wait until it has
    // completed.  No compute warp gets past its first image, no counter is touched, before the builders have.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int j = 0; j < n; ++j) {
      const JobD& J = jobs[j];
      if (J.share) continue;
      const int imgk = J.imgk, bsel = imgk & (kImgBars - 1);
      // x of a job without an in-chain producer is complete: load this thread's first block BEFORE waiting for
      // the image bytes to be released (an L2 hit takes 1-2 us next to the weight stream, most of a conversion)
      const bool pre = MXQ_CHAIN_XPREFETCH && J.dep < 0 && bt < J.nblk;
      uint4 a0 = make_uint4(0u, 0u, 0u, 0u), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
      if (pre) {
        const __half* xa = J.x + (size_t)bt * 64;
        a0 = ld_cg16(xa), a1 = ld_cg16(xa + 8), a2 = ld_cg16(xa + 16), a3 = ld_cg16(xa + 24);
        a4 = ld_cg16(xa + 32), a5 = ld_cg16(xa + 40), a6 = ld_cg16(xa + 48), a7 = ld_cg16(xa + 56);
      }
      CTRACE(3, imgk, 0);
      if (J.imgwait >= 0) mbar_wait(&bars.imgempty[J.imgwait & (kImgBars - 1)], (uint32_t)(J.imgwait >> 3) & 1u);
      CTRACE(3, imgk, 1);
      if (J.dep >= 0) {
        // (letting all 12 compute warps build a dependent job's image instead -- they are idle until x exists --
        // was measured: 303 -> 313 us on the 56-linear decoder chain; the conversion is not what a dependency
        // point costs, the drain / publish / acquire / refill around it is)
        const int* flag = sync_ws + J.dep;
        const long long t0 = clock64();
        while (ld_acquire(flag) < J.dep_target) {
          __nanosleep(64);
          if (clock64() - t0 > 40000000000LL) __trap();
        }
      }
      unsigned char* xi = img + J.imgoff;
      const int nb = J.ximg_blocks, nblk = J.nblk;
      int* tI = reinterpret_cast<int*>(xi + (size_t)nb * 128);
      float* tF = reinterpret_cast<float*>(xi + (size_t)nb * 144);
      if (!(dbg & 4)) {
        for (int b = bt; b < nb; b += kBuilderThreads) {
          // block b = 16 Q + 4 t + i -> image slot [Q][i][k][t][h], table slot [Q][i][t]
          const int Q = b >> 4, tq = (b >> 2) & 3, iq = b & 3;
          unsigned char* dst = xi + (size_t)Q * 2048 + iq * 512 + tq * 32;       // + k * 128 + h * 16
          const int ts = (Q * 16 + iq * 4 + tq) * 4;
          if (b < nblk) {
            if (!(pre && b == bt)) {
              const __half* xa = J.x + (size_t)b * 64;
              a0 = ld_cg16(xa), a1 = ld_cg16(xa + 8), a2 = ld_cg16(xa + 16), a3 = ld_cg16(xa + 24);
              a4 = ld_cg16(xa + 32), a5 = ld_cg16(xa + 40), a6 = ld_cg16(xa + 48), a7 = ld_cg16(xa + 56);
            }
            stage_group<0>(a0, a1, dst, tI + ts + 0, tF + ts + 0);
            stage_group<1>(a2, a3, dst + 128, tI + ts + 1, tF + ts + 1);
            stage_group<2>(a4, a5, dst + 256, tI + ts + 2, tF + ts + 2);
            stage_group<3>(a6, a7, dst + 384, tI + ts + 3, tF + ts + 3);
          } else {                                            // blocks beyond the row (last chunk): zeros
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              reinterpret_cast<uint4*>(dst + k * 128)[0] = z;
              reinterpret_cast<uint4*>(dst + k * 128)[1] = z;
            }
            *reinterpret_cast<uint4*>(tI + ts) = z;
            *reinterpret_cast<uint4*>(tF + ts) = z;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.imgfull[bsel]);
      CTRACE(3, imgk, 2);
    }
  }

  // ---- the last CTA out re-arms the counters (graph replays pass the same arguments) --------------
  __syncthreads();
  CTRACE_CTA(1);
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(sync_ws + kMaxJobs, 1);
    if (ticket == (int)gridDim.x - 1) {
      for (int j = 0; j < n; ++j) sync_ws[j] = 0;
      sync_ws[kMaxJobs] = 0;
      __threadfence();
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<EncodeTiledFn>(fn);
}

// int32 [rows, words] row-major, dense box [box_rows x box_words], zero fill out of bounds
static int make_map(EncodeTiledFn enc, CUtensorMap* map, const void* ptr, int64_t rows, int64_t words, int box_rows,
                    int box_words) {
  cuuint64_t dims[2] = {(cuuint64_t)words, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)words * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_words, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MXQ_OK : MXQ_E_UNSUPPORTED;
}

// Where the activation images live.  Two fixed buffers (image k in buffer k % 2) let the builders start image
// k + 1 when the last compute warp has left image k - 1, i.e. while the CTA works on the jobs of image k: fine
// as long as those take longer than the conversion.  o_proj between q/k/v and gate/up does not (1-2 tiles per
// CTA against 1.2-2.3 us for a 4096-column image): the sets waited ~1 us per layer for the gate/up image
// (161.0 us for bench.py's 56-linear chain).  Such chains get a ring allocation in the same 2 * ximg_max bytes
// (an image is never split; three 4096-column images fit beside a 11008-column one) with the builders up to
// two images ahead: 155.2 us.  Chains that do not need it keep the two buffers -- running ahead there cost
// 1.3 % (151.9 -> 153.8 us with one activation vector per q/k/v/o group), as does running further ahead
// (three images: 157.0 us).  Images are released in order, so image k waits for the LAST earlier image that
// overlaps it, and never runs more than `ahead` images in front.  MXQ_CHAIN_IMGPOOL (profiling): 0 = two
// buffers, n = ring with n images ahead.
static void place_images(JobD* D, int n, int ncta, int ximg_max) {
  const int pool = 2 * ximg_max;
  int first[kMaxJobs], nimg = 0;                      // first job of every image
  double visits[kMaxJobs];                            // chunk visits per CTA on the jobs of an image
  for (int j = 0; j < n; ++j) {
    if (!D[j].share) {
      first[nimg] = j;
      visits[nimg++] = 0.0;
    }
    visits[nimg - 1] += (double)(D[j].ngrp / 4) * D[j].nch / ncta;
  }
  int ahead = 1;
  for (int k = 0; k + 1 < nimg; ++k)                  // conversion of image k + 1: ~2 visits per 4096 columns
    if (visits[k] < 2.0 + 1.5 * D[first[k + 1]].nch) ahead = 2;
  if (const char* e = getenv("MXQ_CHAIN_IMGPOOL")) ahead = atoi(e) > 0 ? atoi(e) : 1;
  if (ahead > kImgBars - 1) ahead = kImgBars - 1;
  int off[kMaxJobs], len[kMaxJobs], wait[kMaxJobs];
  bool ring = ahead > 1;
  if (ring) {
    int cursor = 0;
    for (int k = 0; k < nimg && ring; ++k) {
      len[k] = D[first[k]].ximg_blocks * 160;
      if (cursor + len[k] > pool) cursor = 0;
      off[k] = cursor;
      cursor += len[k];
      wait[k] = k - ahead - 1;
      for (int m = k - 1; m > wait[k] && m >= 0; --m)
        if (off[m] < off[k] + len[k] && off[k] < off[m] + len[m]) {
          wait[k] = m;
          break;
        }
      if (k > 0 && wait[k] == k - 1) ring = false;    // fragmented: image k could not be built beside k - 1
    }
  }
  for (int k = 0; k < nimg && !ring; ++k) {
    off[k] = (k & 1) * ximg_max;
    wait[k] = k - 2;
  }
  for (int j = 0; j < n; ++j) {
    D[j].imgoff = off[D[j].imgk];
    D[j].imgwait = wait[D[j].imgk] < 0 ? -1 : wait[D[j].imgk];
  }
}

}  // namespace g3
}  // namespace mxq

using namespace mxq;

extern "C" size_t mxq_gemv_chain_plan_bytes(void) { return g3::kPlanBytes; }

extern "C" int mxq_gemv_chain_plan(const mxq_gemv_job_t* jobs, int n, void* plan_host) {
  if (!jobs || !plan_host) return MXQ_E_NULL;
  if (reinterpret_cast<uintptr_t>(plan_host) & 63) return MXQ_E_ALIGN;
  if (n <= 0 || n > g3::kMaxJobs) return MXQ_E_SHAPE;
  g3::EncodeTiledFn enc = nullptr;                 // looked up after the first job has passed validation
  memset(plan_host, 0, g3::kPlanBytes);
  g3::ChainParams* P = reinterpret_cast<g3::ChainParams*>(plan_host);
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(reinterpret_cast<unsigned char*>(plan_host) + g3::kMapsOffset);
  g3::PlanH& H = P->H;
  g3::JobD* D = P->jobs;
  const int ncta = kNumSMs;
  int ximg_max = 0, coop = 0, next = 0, nimg = 0;
  for (int j = 0; j < n; ++j) {
    const mxq_gemv_job_t& a = jobs[j];
    const mxq_packed_t& w = a.w;
    MXQ_CHECK_PTR(a.x);
    MXQ_CHECK_PTR(a.y);
    MXQ_CHECK_PTR(w.weight);
    if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b || !w.zeros_4b)
      return MXQ_E_NULL;
    if ((reinterpret_cast<uintptr_t>(w.weight_last) | reinterpret_cast<uintptr_t>(w.zeros_and_scales) |
         reinterpret_cast<uintptr_t>(w.zeros_2nd) | reinterpret_cast<uintptr_t>(w.scales_2nd) |
         reinterpret_cast<uintptr_t>(w.scales_4b) | reinterpret_cast<uintptr_t>(w.zeros_4b)) & 15)
      return MXQ_E_ALIGN;
    if (a.IC <= 0 || a.OC <= 0 || a.IC % 256 || a.OC % 32 || a.IC > 32768 || a.OC > (1 << 24)) return MXQ_E_UNSUPPORTED;
    if (a.dep >= j || a.dep < -1) return MXQ_E_SHAPE;
    g3::JobD& d = D[j];
    d.S2 = reinterpret_cast<const unsigned char*>(w.scales_2nd);
    d.S4 = reinterpret_cast<const unsigned char*>(w.scales_4b);
    d.Z4 = reinterpret_cast<const unsigned char*>(w.zeros_4b);
    d.x = reinterpret_cast<const __half*>(a.x);
    d.y = reinterpret_cast<__half*>(a.y);
    d.nblk = (int)(a.IC / 64);
    d.nch = (d.nblk + 63) / 64;
    d.nqb = d.nblk / 4;
    d.ngrp = (int)(a.OC / 4);
    d.oc = (int)a.OC;
    const int t_total = d.ngrp / 4;                  // 16-row tiles (OC % 32 == 0: every tile is full)
    d.tbase = t_total / ncta;
    d.trem = t_total % ncta;
    d.dep = a.dep;
    d.dep_target = a.dep >= 0 ? D[a.dep].ngrp / 4 : 0;     // number of 16-row tiles of that job over all CTAs
    if (a.dep >= 0) {
      coop = 1;
      D[a.dep].publish = 1;
    }
    // whole rows in one chunk and 16-byte aligned row groups: one scales_2nd copy per stage
    d.s2pitch = (d.nch == 1 && (d.nblk * 6) % 16 == 0) ? d.nblk * 6 : 0;
    d.ximg_blocks = d.nch * 64;
    const int bw = d.nblk * 4 < 256 ? d.nblk * 4 : 256, bwl = d.nblk < 64 ? d.nblk : 64;   // box widths in words
    d.pw = bw * 4;
    d.pwl = bwl * 4;
    // jobs that read the same x with the same shape share one activation image and one CTA assignment
    d.share = j > 0 && jobs[j - 1].x == a.x && jobs[j - 1].IC == a.IC && jobs[j - 1].OC == a.OC && jobs[j - 1].dep == a.dep;
    if (d.share && d.tbase == 0) {
      d.rot = D[j - 1].rot;                         // fewer tiles than CTAs: the same CTAs as the image's owner
    } else {                                        // (with tbase >= 1 every CTA is active under any rotation)
      d.rot = (ncta - next) % ncta;                 // slice 0 of this job = CTA `next`
      next = (next + d.trem) % ncta;                // the slices with an extra tile move on
    }
    const int ximg = d.ximg_blocks * 160;
    if (ximg > ximg_max) ximg_max = ximg;
    d.imgk = d.share ? D[j - 1].imgk : nimg++;         // images are numbered over the chain, not per CTA
    if (!enc) {
      enc = g3::get_encode();
      if (!enc) return MXQ_E_UNSUPPORTED;             // no driver (CPU-only host): cuTensorMapEncodeTiled is unavailable
    }
    int rc = g3::make_map(enc, maps + j * 4 + 0, w.weight, a.OC, (int64_t)d.nblk * 4, 16, bw);
    if (!rc) rc = g3::make_map(enc, maps + j * 4 + 1, w.weight_last, a.OC, d.nblk, 16, bwl);
    if (!rc) rc = g3::make_map(enc, maps + j * 4 + 2, w.zeros_and_scales, a.OC, (int64_t)d.nch * 32, 16, 32);
    if (!rc) rc = g3::make_map(enc, maps + j * 4 + 3, w.zeros_2nd, a.OC / 4, (int64_t)d.nch * 32, 4, 32);
    if (rc) return rc;
  }
  g3::place_images(D, n, ncta, ximg_max);
  const size_t fixed = 2 * (size_t)ximg_max + g3::kRedBytes + ((sizeof(g3::Bars) + 127) & ~size_t(127)) +
                       (size_t)n * sizeof(g3::JobD) + (size_t)g3::kMaxTiles * 32 + (2 * g3::kMaxJobs + 2) * 4 + 128;
  int tiles_max = 0;                                  // 16-row tiles of a CTA with a full share of every job
  for (int j = 0; j < n; ++j) tiles_max += D[j].tbase + (D[j].trem ? 1 : 0);
  if (tiles_max > g3::kMaxTiles) return MXQ_E_UNSUPPORTED;
  if (fixed + g3::kSets * (size_t)g3::kStageBytes > g3::kSmemMax) return MXQ_E_UNSUPPORTED;   // one stage per compute set
  int S = (int)((g3::kSmemMax - fixed) / g3::kStageBytes);
  if (S > g3::kMaxStages) S = g3::kMaxStages;
  H.magic = g3::kMagic;
  H.n = n;
  H.nstages = S;
  H.ximg_max = ximg_max;
  H.ncta = ncta;
  H.coop = coop;
  if (const char* e = getenv("MXQ_CHAIN_DBG")) H.dbg = atoi(e);
  H.smem = (int)((size_t)S * g3::kStageBytes + fixed);
  return MXQ_OK;
}

extern "C" int mxq_gemv_chain_run(const void* plan_host, const void* plan_dev, int32_t* sync_ws, unsigned flags,
                                  void* stream) {
  if (!plan_host || !plan_dev || !sync_ws) return MXQ_E_NULL;
  if (reinterpret_cast<uintptr_t>(plan_dev) & 63) return MXQ_E_ALIGN;
  const g3::ChainParams& P = *reinterpret_cast<const g3::ChainParams*>(plan_host);
  const g3::PlanH& H = P.H;
  if (H.magic != g3::kMagic || H.n <= 0 || H.n > g3::kMaxJobs) return MXQ_E_SHAPE;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  if (sms < H.ncta) return MXQ_E_UNSUPPORTED;       // the plan assumes one resident CTA per B200 SM
  // one opt-in limit for every chain (the largest a plan can ask for): concurrent callers with different
  // plans never lower each other's limit between this call and their launch
  e = cudaFuncSetAttribute(g3::gemv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g3::kSmemMax);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)H.ncta);
  cfg.blockDim = dim3(g3::kThreads);
  cfg.dynamicSmemBytes = (size_t)H.smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  if (H.coop) {                                     // all CTAs co-resident; no overlap with the previous kernel
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  } else if (flags & MXQ_GEMV_CHAIN_PDL) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  }
  const CUtensorMap* maps =
      reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const unsigned char*>(plan_dev) + g3::kMapsOffset);
  e = cudaLaunchKernelEx(&cfg, g3::gemv_chain_kernel, P, maps, reinterpret_cast<int*>(sync_ws), 4, 16, 64, 256);
  return e == cudaSuccess ? MXQ_OK : (int)e;
}

// profiling aid, not part of the documented surface
extern "C" __attribute__((visibility("default"))) int mxq_debug_chain_trace(long long* host_out) {
#ifdef MXQ_CHAIN_TRACE
  return (int)cudaMemcpyFromSymbol(host_out, g3::g_ctrace, sizeof(long long) * 5 * g3::kTraceEvents * 4);
#else
  (void)host_out;
  return MXQ_E_UNSUPPORTED;                        // build with MXQ_CHAIN_TRACE=1 (mxq_b200/build.py)
#endif
}
