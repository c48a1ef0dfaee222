// Fused MXQGPT.fasterquant(blocksize=16) + Quantizer + packer (sm_100a).
//
// Replaces mxq_quant/lib/mxqgpt.py:387-448 and mxq_quant/lib/quantizer.py:5-20,61-121,149-155
// (192 Quantizer() constructions and ~30 ATen calls per 64-column block) and produces the packed
// mixed 2/4-bit layout consumed by gemv_mxq_cuda.cu:39-208 (the reference has no producer for it;
// encode policy = oracle pack_mxq) -- in two kernels:
//   (A) pool_prepass : one warp per row: min/max over the pooled (4-bit) columns with dead
//                      columns zeroed; also the packer's per-row 4-bit scale / zero-point;
//   (B) ptq_pack_tile: one warp per (16-row tile, 64-column block).  lane = (row pair, slot):
//                      a thread owns the 16 columns of one group for rows r and r+8, so group
//                      min/max, scale and zero need no shuffles and are computed exactly once;
//                      the 16-row second-level scale quantisation (quantizer.py:114-121) is three
//                      xor-shuffles, the packer's 4-row second level two.  W is read once and
//                      both outputs (fp16 fake-quant weights, packed tensors) are written.
// All arithmetic is op-by-op IEEE fp32 like the reference; x/scale uses a correctly-rounded
// reciprocal + one Markstein step (bit-identical to division, oracle/div_check.c); rounding is
// clamp-then-magic-add (clamp and round-half-even commute because the bounds are integers).
#include <cstdlib>

#include "common.cuh"

namespace mxq {

__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float rint_any(float v) {   // |v| < 2^22
  return __fadd_rn(__fadd_rn(v, 12582912.0f), -12582912.0f);
}

// quantizer.py:81-99: scale, zero from (xmin, xmax) with the degenerate fix
__device__ __forceinline__ void find_params(float xmin, float xmax, float maxq, float& scale,
                                            float& zero) {
  if (xmin == xmax) { xmin = -1.f; xmax = 1.f; }
  scale = fdiv(__fsub_rn(xmax, xmin), maxq);
  zero = fdiv(-xmin, scale);
}

__global__ void dead_mask_kernel(const float* __restrict__ colstat, uint8_t* __restrict__ dead,
                                 int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) dead[c] = colstat ? (colstat[c] == 0.f) : 0;
}

__device__ __forceinline__ void apply_dead8(uint32_t lo, uint32_t hi, float* f) {
  if (lo | hi) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t w = e < 4 ? lo : hi;
      if ((w >> (8 * (e & 3))) & 0xFF) f[e] = 0.f;
    }
  }
}

// (A) one warp per row, 8 rows per CTA.  pool_mm[row] = (min, max) over pooled columns;
//     pack_pool[row] = (s4f, z4) and scales_4b / zeros_4b of the packed layout when kPack.
template <bool kRef, bool kPack>
__global__ void __launch_bounds__(256) pool_prepass_kernel(const __half* __restrict__ W,
                                                           const uint8_t* __restrict__ dead,
                                                           const uint8_t* __restrict__ gbits,
                                                           float2* __restrict__ pool_mm,
                                                           float2* __restrict__ pack_pool,
                                                           __half* __restrict__ scales_4b,
                                                           uint32_t* __restrict__ zeros_4b,
                                                           int rows, int cols) {
  __shared__ uint32_t z4s[8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + w;
  const int nchunks = cols / 8;   // 16-byte chunks; group 16 = 2 chunks
  float mn = INFINITY, mx = -INFINITY;
  if (row < rows) {
    const size_t roff = (size_t)row * cols;
    const int npc = kRef ? nchunks / 4 : nchunks;
    auto take = [&](const uint4& wv, const uint2& dm) {
      float f[8];
      DT<__half>::unpack(wv, f);
      apply_dead8(dm.x, dm.y, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
    };
    int m = lane;
    if (kRef) {
      // four chunks per lane in flight: a row's pooled quarter is 1 KB (4096 columns) to 2.7 KB, and a
      // warp walking it one load at a time is a chain of DRAM latencies
      for (; m + 96 < npc; m += 128) {
        uint4 wv[4];
        uint2 dm[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int mm = m + 32 * u;
          const int c = ((mm >> 1) * 4 + 3) * 2 + (mm & 1);
          wv[u] = ld_stream(W + roff + c * 8);
          dm[u] = *reinterpret_cast<const uint2*>(dead + c * 8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) take(wv[u], dm[u]);
      }
    }
    for (; m < npc; m += 32) {
      const int c = kRef ? ((m >> 1) * 4 + 3) * 2 + (m & 1) : m;
      if (!kRef && !(gbits[c >> 1] & MXQ_POOL_FLAG)) continue;
      take(*reinterpret_cast<const uint4*>(W + roff + c * 8), *reinterpret_cast<const uint2*>(dead + c * 8));
    }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if (row < rows && lane == 0) pool_mm[row] = make_float2(mn, mx);
  if (kPack) {
    const float lo = fminf(mn, 0.f), hi = fmaxf(mx, 0.f);
    __half s4h = __float2half_rn(fdiv(__fsub_rn(hi, lo), 15.f));
    if (__half2float(s4h) == 0.f) s4h = __float2half_rn(1.f);
    const float s4f = __half2float(s4h);
    const float z4 = clampf(rintf(fdiv(-lo, s4f)), 0.f, 15.f);
    if (lane == 0) {
      z4s[w] = row < rows ? (uint32_t)z4 : 0u;
      if (row < rows) {
        pack_pool[row] = make_float2(s4f, z4);
        scales_4b[row] = s4h;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t word = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) word |= z4s[i] << (4 * i);
      zeros_4b[blockIdx.x] = word;
    }
  }
}

struct TileParams {
  const __half* W;
  __half* Wq;
  uint8_t* codes;
  const uint8_t* dead;
  const uint8_t* gbits;
  const float2* pool_mm;
  const float2* pack_pool;
  mxq_packed_t out;
  int rows, cols, low_bits, pool_bits;
};

// (B)
template <bool kRef, bool kQuant, bool kPack, bool kCodes>
__global__ void __launch_bounds__(256) ptq_pack_tile_kernel(const TileParams p) {
  const int lane = threadIdx.x & 31;
  const int rr = lane >> 2, k = lane & 3;
  const int nblk = p.cols / 64;
  const int nchunk = (nblk + 63) / 64;
  const int64_t units = (int64_t)(p.rows / 16) * nblk;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t wstride = (int64_t)gridDim.x * 8;
  const float maxq_low = (float)((1 << p.low_bits) - 1);
  const float maxq_pool = (float)((1 << p.pool_bits) - 1);
  uint16_t* zs16 = reinterpret_cast<uint16_t*>(p.out.zeros_and_scales);
  __half* s2o = reinterpret_cast<__half*>(p.out.scales_2nd);

  for (int64_t u = warp0; u < units; u += wstride) {
    const int tile = (int)(u / nblk), blk = (int)(u % nblk);
    const int col0 = blk * 64 + k * 16;
    const uint4 dm = *reinterpret_cast<const uint4*>(p.dead + col0);
    float x[2][16];
    float mn[2], mx[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const size_t off = (size_t)(tile * 16 + rr + 8 * i) * p.cols + col0;
      const uint4 c0 = ld_stream(p.W + off), c1 = ld_stream(p.W + off + 8);
      DT<__half>::unpack(c0, x[i]);
      DT<__half>::unpack(c1, x[i] + 8);
      apply_dead8(dm.x, dm.y, x[i]);
      apply_dead8(dm.z, dm.w, x[i] + 8);
      mn[i] = x[i][0]; mx[i] = x[i][0];
#pragma unroll
      for (int e = 1; e < 16; ++e) { mn[i] = fminf(mn[i], x[i][e]); mx[i] = fmaxf(mx[i], x[i][e]); }
    }

    if (kQuant) {
      // ---- MXQGPT.fasterquant semantics ----------------------------------------------------
      const bool pooled = kRef ? (k == 3) : ((p.gbits[blk * 4 + k] & MXQ_POOL_FLAG) != 0);
      const float maxq = pooled ? maxq_pool : maxq_low;
      float scale[2], zero[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float a = mn[i], b = mx[i];
        if (pooled) {
          const float2 mm = p.pool_mm[tile * 16 + rr + 8 * i];
          a = mm.x; b = mm.y;
        }
        find_params(a, b, maxq, scale[i], zero[i]);
      }
      // second level over the 16 rows of the tile (quantizer.py:114-121)
      float smin = fminf(scale[0], scale[1]), smax = fmaxf(scale[0], scale[1]);
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
      }
      float s2, z2;
      find_params(smin, smax, 15.f, s2, z2);
      const float s2c = fmaxf(s2, 1e-9f);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float qs = clampf(rintf(__fadd_rn(fdiv(scale[i], s2c), z2)), 0.f, 15.f);
        const float sq = __fmul_rn(s2, __fsub_rn(qs, z2));       // dequantized scale
        const float sc = fmaxf(sq, 1e-9f);
        const float rc = __frcp_rn(sc);
        float o[16];
        uint32_t cw[4] = {0, 0, 0, 0};
        // two elements per instruction (FMUL2 / FFMA2 / FADD2 are IEEE-RN per lane: bit-identical to
        // the scalar chain).  clamp(v, 0, maxq) + M == clamp(v + M, M, M + maxq) because v -> RN(v + M)
        // is monotone and exact at the integer bounds; NaN / inf end at the same bound either way.
        const f32x2 nsc2 = pk2(-sc, -sc), rc2 = pk2(rc, rc), zp2 = pk2(zero[i], zero[i]);
        const f32x2 M2 = pk2(12582912.0f, 12582912.0f), nM2 = pk2(-12582912.0f, -12582912.0f), sq2 = pk2(sq, sq);
        const float m_hi = __fadd_rn(12582912.0f, maxq);
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const f32x2 m2 = add2(add2(div2_rn_by(pk2(x[i][e], x[i][e + 1]), nsc2, rc2), zp2), M2);
          float m0, m1;
          upk2(m2, m0, m1);
          m0 = fminf(fmaxf(m0, 12582912.0f), m_hi);
          m1 = fminf(fmaxf(m1, 12582912.0f), m_hi);
          if (kCodes) {
            cw[e >> 2] |= (uint32_t)(__float_as_int(m0) & 0xFF) << (8 * (e & 3));
            cw[e >> 2] |= (uint32_t)(__float_as_int(m1) & 0xFF) << (8 * ((e + 1) & 3));
          }
          upk2(mul2(sq2, sub2(add2(pk2(m0, m1), nM2), zp2)), o[e], o[e + 1]);
        }
        const size_t off = (size_t)(tile * 16 + rr + 8 * i) * p.cols + col0;
        *reinterpret_cast<uint4*>(p.Wq + off) = DT<__half>::pack(o);
        *reinterpret_cast<uint4*>(p.Wq + off + 8) = DT<__half>::pack(o + 8);
        if (kCodes) *reinterpret_cast<uint4*>(p.codes + off) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      }
    }

    if (kPack) {
      // ---- packed layout (encode policy: oracle pack_mxq) ----------------------------------
      // One code path for the 2-bit slots (k < 3) and the 4-bit slot (k == 3): only the scale,
      // zero-point, code range and code width differ per lane, so the warp never diverges.
      const bool four = (k == 3);
      const float maxq = four ? 15.f : 3.f;
      // codes are assembled by Horner's rule on the raw float bits: m = 1.5*2^23 + code exactly, so
      // bits(m) = 0x4B400000 + code and  word = sum_e code_e * radix^e = H - 0x4B400000 * sum_e radix^e
      const uint32_t radix = four ? 16u : 4u;
      const uint32_t horner_bias = 0x4B400000u * (four ? 0x11111111u : 0x00005555u);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = tile * 16 + rr + 8 * i;
        // second-order (4-row) scale of the 2-bit slots; slot 3 runs the shuffles but ignores them
        const float lo = fminf(mn[i], 0.f), hi = fmaxf(mx[i], 0.f);
        float s = fdiv(__fsub_rn(hi, lo), 3.f);
        if (s == 0.f) s = 1.f;
        float smax = s;
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, 4));
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, 8));
        const __half s2h = __float2half_rn(fdiv(smax, 3.f));
        const float s2f = __half2float(s2h);
        const float c = clampf(rintf(fdiv(s, s2f)), 1.f, 3.f);
        float S = __fmul_rn(s2f, c);
        float z = clampf(rintf(fdiv(-lo, S)), 0.f, 3.f);
        uint32_t meta = ((uint32_t)z | ((uint32_t)c << 8)) << (2 * k);
        if (four) {
          const float2 pz = p.pack_pool[row];
          S = pz.x; z = pz.y; meta = 0;
        }
        const float rS = __frcp_rn(S);
        const float m_lo = 12582912.0f, m_hi = __fadd_rn(12582912.0f, maxq);
        uint32_t acc_lo = 0, acc_hi = 0;                 // codes 0..7 and 8..15 of this slot
        const f32x2 nS2 = pk2(-S, -S), rS2 = pk2(rS, rS), Mp2 = pk2(12582912.0f, 12582912.0f), zz2 = pk2(z, z);
#pragma unroll
        for (int e = 6; e >= 0; e -= 2) {
          // code = clamp(rint(w / S) + z, 0, maxq), evaluated in the 1.5*2^23 "integer" domain:
          // (v + M) holds rint(v) exactly, adding the integer z and clamping stay exact.  Elements
          // e and e + 1 share a packed instruction (the same pairing as the quantizer above, so the
          // register pairs are formed once).
          float a0, a1, b0, b1;
          upk2(add2(add2(div2_rn_by(pk2(x[i][e], x[i][e + 1]), nS2, rS2), Mp2), zz2), a0, a1);
          upk2(add2(add2(div2_rn_by(pk2(x[i][e + 8], x[i][e + 9]), nS2, rS2), Mp2), zz2), b0, b1);
          acc_lo = acc_lo * radix + __float_as_uint(fminf(fmaxf(a1, m_lo), m_hi));
          acc_lo = acc_lo * radix + __float_as_uint(fminf(fmaxf(a0, m_lo), m_hi));
          acc_hi = acc_hi * radix + __float_as_uint(fminf(fmaxf(b1, m_lo), m_hi));
          acc_hi = acc_hi * radix + __float_as_uint(fminf(fmaxf(b0, m_lo), m_hi));
        }
        acc_lo -= horner_bias;
        acc_hi -= horner_bias;
        const uint32_t word = four ? acc_lo : (acc_lo | (acc_hi << 16));
        const uint32_t word_last = acc_hi;
        // gather the block's four weight words and the metadata half-word into slot 0
        const uint32_t w1 = __shfl_down_sync(0xffffffffu, word, 1);
        const uint32_t w2 = __shfl_down_sync(0xffffffffu, word, 2);
        const uint32_t w3 = __shfl_down_sync(0xffffffffu, word, 3);
        const uint32_t m1 = __shfl_down_sync(0xffffffffu, meta, 1);
        const uint32_t m2 = __shfl_down_sync(0xffffffffu, meta, 2);
        const int chunk = blk >> 6, bp = blk & 63;
        const int mword = chunk * 32 + (bp & 31), ph = bp >> 5;
        if (k == 0) {
          *reinterpret_cast<uint4*>(p.out.weight + (size_t)row * nblk * 4 + (size_t)blk * 4) =
              make_uint4(word, w1, w2, w3);
          zs16[((size_t)row * 32 * nchunk + mword) * 2 + ph] = (uint16_t)(meta | m1 | m2);
        } else if (k == 3) {
          p.out.weight_last[(size_t)row * nblk + blk] = (int32_t)word_last;
        }
        if (k < 3 && (row & 3) == 0) s2o[(size_t)(row >> 2) * nblk * 3 + (size_t)blk * 3 + k] = s2h;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (C) Reference recipe, ONE kernel per linear: a CTA owns a 16-row tile (the unit of the
// second-level scale quantisation, quantizer.py:115) and walks all of K twice:
//   phase 0  dead-column flags of the whole row (colstat == 0) into shared memory; the tile's
//            share of the all-zero zeros_2nd tensor and the padding half-words of a ragged last
//            metadata chunk are written here (no memset launches);
//   phase 1  warp = 2 rows: min/max over the pooled (4-bit) columns -- only those 32 bytes of
//            every 128-byte line are requested; the fill brings the lines into L2, where phase 2
//            finds them (W crosses the HBM interface once: the round-1 pre-pass kernel re-read it);
//   phase 2  warp = (16 rows x 64 columns) units, lane = (row pair, slot) as in (B), with the
//            element loops written on packed registers end to end: group min/max on half2, the two
//            quantizers on f32x2 (division = RN reciprocal + Markstein step, bit-identical to IEEE
//            division), one 3-input min/max per clamp, codes assembled by Horner on the raw bits.
// G2 = width of the 2-bit groups inside the 48 low columns of a block for the fake-quantizer:
// 16 (nas_quant, prune.py:409), 32, or 48 (fasterquant's default blocksize=128, mxqgpt.py:388,413-415).
// The packed layout is defined for 16 only.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hmin2_u(uint32_t a, uint32_t b) { return P16<__half>::vmin(a, b); }
__device__ __forceinline__ uint32_t hmax2_u(uint32_t a, uint32_t b) { return P16<__half>::vmax(a, b); }
__device__ __forceinline__ f32x2 h2_to_f2(uint32_t h) {      // {lo, hi} halves -> two floats
  float a, b;
  asm("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(a), "=f"(b) : "r"(h));
  return pk2(a, b);
}
__device__ __forceinline__ uint32_t f2_to_h2(f32x2 v) {      // RN to fp16, {lo, hi}
  float a, b;
  upk2(v, a, b);
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ float clamp3(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
// zero the halves of w[0..7] whose dead byte is set (dm = 16 dead bytes of the lane's columns)
__device__ __forceinline__ void apply_dead16(const uint4& dm, uint32_t* w) {
  const uint32_t d[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t two = (d[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;       // dead bytes of columns 2i, 2i+1
    if (two) w[i] &= ((two & 0xFFu) ? 0u : 0x0000FFFFu) | ((two >> 8) ? 0u : 0xFFFF0000u);
  }
}

// a / b for positive normal b and normal-or-zero a: RN reciprocal + one Markstein correction.
// Bit-identical to IEEE division: oracle/div_check_general.c (exhaustive for b = 3 and 15, 1.2e10
// random pairs incl. fp16-valued operands and all-ones divisor mantissas: no mismatch); __fdiv_rn
// costs ~15 instructions with its range check and slow-path call, this costs 6.
__device__ __forceinline__ float fdiv_n(float a, float b) { return div_rn_by(a, b, rcp_rn_normal(b)); }
__device__ __forceinline__ void find_params_n(float xmin, float xmax, float maxq, float rmaxq, float& scale,
                                              float& zero) {
  if (xmin == xmax) { xmin = -1.f; xmax = 1.f; }
  scale = div_rn_by(__fsub_rn(xmax, xmin), maxq, rmaxq);
  zero = fdiv_n(-xmin, scale);
}

struct Tile16Params {
  const __half* W;
  __half* Wq;
  uint8_t* codes;
  const float* colstat;
  mxq_packed_t out;
  int rows, cols, low_bits;
  int ksplit;    // CTAs per 16-row tile (each takes every ksplit-th group of 8 blocks)
};

template <bool kQuant, bool kPack, bool kCodes, int G2>
__global__ void __launch_bounds__(256, 2) ptq_tile16_kernel(const Tile16Params p) {
  extern __shared__ __align__(16) unsigned char t16_smem[];
  __shared__ float2 s_pool_mm[16];       // (min, max) over the pooled columns of each row
  __shared__ float2 s_pack_pool[16];     // packer: (s4 as float, z4)
  unsigned char* s_dead = t16_smem;      // [cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x / p.ksplit, kpart = blockIdx.x - tile * p.ksplit;
  const int cols = p.cols;
  const int nblk = cols >> 6;
  const int nchunk = (nblk + 63) >> 6;
  const int row_t0 = tile * 16;

  // ---- phase 0 ---------------------------------------------------------------------------------
  for (int c = threadIdx.x * 4; c < cols; c += 256 * 4) {
    uint32_t d = 0;
    if (p.colstat) {
      const float4 v = *reinterpret_cast<const float4*>(p.colstat + c);
      d = (v.x == 0.f ? 1u : 0u) | (v.y == 0.f ? 0x100u : 0u) | (v.z == 0.f ? 0x10000u : 0u) | (v.w == 0.f ? 0x1000000u : 0u);
    }
    *reinterpret_cast<uint32_t*>(s_dead + c) = d;
  }
  if (kPack && kpart == 0) {
    // z2 == 0 policy: the tile's 4 row groups of zeros_2nd are all zero; in a ragged last metadata
    // chunk the unused half-words of zeros_and_scales must be zero too (the kernel writes half-words)
    uint32_t* z2 = reinterpret_cast<uint32_t*>(p.out.zeros_2nd) + (size_t)(row_t0 >> 2) * 32 * nchunk;
    for (int i = threadIdx.x; i < 4 * 32 * nchunk; i += 256) z2[i] = 0u;
    if (nblk & 63) {
      uint32_t* zs = reinterpret_cast<uint32_t*>(p.out.zeros_and_scales) + (size_t)row_t0 * 32 * nchunk;
      for (int i = threadIdx.x; i < 16 * 32; i += 256) zs[(size_t)(i >> 5) * 32 * nchunk + (nchunk - 1) * 32 + (i & 31)] = 0u;
    }
  }
  __syncthreads();

  // ---- phase 1: pooled min / max, warp = rows 2w, 2w + 1 ------------------------------------------
  {
    uint32_t mn2[2] = {0x7C007C00u, 0x7C007C00u}, mx2[2] = {0xFC00FC00u, 0xFC00FC00u};
    const int npc = nblk * 2;                         // pooled 16-byte chunks per row
    for (int m0 = lane; m0 < npc; m0 += 128) {
      uint4 v[2][4];
      uint2 dm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = m0 + 32 * u;
        const int c8 = ((m >> 1) * 8 + 6 + (m & 1)) * 8;     // first column of the chunk
        if (m < npc) {
#pragma unroll
          for (int r = 0; r < 2; ++r) v[r][u] = ld_stream(p.W + (size_t)(row_t0 + 2 * warp + r) * cols + c8);
          dm[u] = *reinterpret_cast<const uint2*>(s_dead + c8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (m0 + 32 * u < npc) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            uint32_t w[4] = {v[r][u].x, v[r][u].y, v[r][u].z, v[r][u].w};
            if (dm[u].x | dm[u].y) {
              const uint32_t d[2] = {dm[u].x, dm[u].y};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t two = (d[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
                if (two) w[i] &= ((two & 0xFFu) ? 0u : 0x0000FFFFu) | ((two >> 8) ? 0u : 0xFFFF0000u);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { mn2[r] = hmin2_u(mn2[r], w[i]); mx2[r] = hmax2_u(mx2[r], w[i]); }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mn = fminf(P16<__half>::lo(mn2[r]), P16<__half>::hi(mn2[r]));
      float mx = fmaxf(P16<__half>::lo(mx2[r]), P16<__half>::hi(mx2[r]));
      mn = warp_min(mn);
      mx = warp_max(mx);
      if (lane == 0) {
        const int rl = 2 * warp + r;
        s_pool_mm[rl] = make_float2(mn, mx);
        if (kPack) {
          const float lo = fminf(mn, 0.f), hi = fmaxf(mx, 0.f);
          __half s4h = __float2half_rn(fdiv(__fsub_rn(hi, lo), 15.f));
          if (__half2float(s4h) == 0.f) s4h = __float2half_rn(1.f);
          const float s4f = __half2float(s4h);
          const float z4 = clampf(rintf(fdiv(-lo, s4f)), 0.f, 15.f);
          s_pack_pool[rl] = make_float2(s4f, z4);
          if (kpart == 0) reinterpret_cast<__half*>(p.out.scales_4b)[row_t0 + rl] = s4h;
        }
      }
    }
  }
  __syncthreads();
  if (kPack && kpart == 0 && threadIdx.x < 2) {
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) word |= (uint32_t)s_pack_pool[threadIdx.x * 8 + i].y << (4 * i);
    reinterpret_cast<uint32_t*>(p.out.zeros_4b)[(row_t0 >> 3) + threadIdx.x] = word;
  }

  // ---- phase 2 -------------------------------------------------------------------------------------
  const int rr = lane >> 2, k = lane & 3;
  const bool four = (k == 3);
  const float maxq_low = (float)((1 << p.low_bits) - 1);
  const f32x2 M2 = pk2(12582912.0f, 12582912.0f);
  uint16_t* zs16 = reinterpret_cast<uint16_t*>(p.out.zeros_and_scales);
  __half* s2o = reinterpret_cast<__half*>(p.out.scales_2nd);

  const float rq_low = rcp_rn_normal(maxq_low);
  for (int blk = kpart * 8 + warp; blk < nblk; blk += 8 * p.ksplit) {
    const int col0 = blk * 64 + k * 16;
    uint32_t xh[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const __half* src = p.W + (size_t)(row_t0 + rr + 8 * i) * cols + col0;
      const uint4 c0 = ld_stream(src), c1 = ld_stream(src + 8);
      xh[i][0] = c0.x; xh[i][1] = c0.y; xh[i][2] = c0.z; xh[i][3] = c0.w;
      xh[i][4] = c1.x; xh[i][5] = c1.y; xh[i][6] = c1.z; xh[i][7] = c1.w;
    }
    const uint4 dm = *reinterpret_cast<const uint4*>(s_dead + col0);
    if (dm.x | dm.y | dm.z | dm.w) { apply_dead16(dm, xh[0]); apply_dead16(dm, xh[1]); }
    float mn[2], mx[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t a = xh[i][0], b = xh[i][0];
#pragma unroll
      for (int e = 1; e < 8; ++e) { a = hmin2_u(a, xh[i][e]); b = hmax2_u(b, xh[i][e]); }
      mn[i] = fminf(P16<__half>::lo(a), P16<__half>::hi(a));
      mx[i] = fmaxf(P16<__half>::lo(b), P16<__half>::hi(b));
    }
    // fp32 pairs, formed once for both quantizers
    f32x2 x2[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) x2[i][e] = h2_to_f2(xh[i][e]);

    if (kQuant) {
      // ---- MXQGPT.fasterquant ---------------------------------------------------------------------
      const float maxq = four ? 15.f : maxq_low;
      float scale[2], zero[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float a = mn[i], b = mx[i];
        if (G2 > 16) {
          // 2-bit groups wider than a slot: {0,1},{2} (G2 = 32) or {0,1,2} (G2 = 48); slot 3 never joins
          const float a1 = __shfl_xor_sync(0xffffffffu, a, 1), b1 = __shfl_xor_sync(0xffffffffu, b, 1);
          const float a2 = __shfl_xor_sync(0xffffffffu, a, 2), b2 = __shfl_xor_sync(0xffffffffu, b, 2);
          const float a3 = __shfl_xor_sync(0xffffffffu, a, 3), b3 = __shfl_xor_sync(0xffffffffu, b, 3);
          if (G2 == 32) {
            if (k < 2) { a = fminf(a, a1); b = fmaxf(b, b1); }
          } else {
            if (k == 0) { a = fminf(a, fminf(a1, a2)); b = fmaxf(b, fmaxf(b1, b2)); }
            else if (k == 1) { a = fminf(a, fminf(a1, a3)); b = fmaxf(b, fmaxf(b1, b3)); }
            else if (k == 2) { a = fminf(a, fminf(a2, a3)); b = fmaxf(b, fmaxf(b2, b3)); }
          }
        }
        if (four) { const float2 mm = s_pool_mm[rr + 8 * i]; a = mm.x; b = mm.y; }
        find_params_n(a, b, maxq, four ? 0.0666666701436043f : rq_low, scale[i], zero[i]);
      }
      // second level over the 16 rows of the tile (quantizer.py:114-121)
      float smin = fminf(scale[0], scale[1]), smax = fmaxf(scale[0], scale[1]);
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
      }
      float s2, z2;
      find_params_n(smin, smax, 15.f, 0.0666666701436043f, s2, z2);
      const float s2c = fmaxf(s2, 1e-9f);
      const float rs2c = rcp_rn_normal(s2c);
      const float m_hi = __fadd_rn(12582912.0f, maxq);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float qs = clampf(rint_any(__fadd_rn(div_rn_by(scale[i], s2c, rs2c), z2)), 0.f, 15.f);
        const float sq = __fmul_rn(s2, __fsub_rn(qs, z2));       // dequantized scale
        const float sc = fmaxf(sq, 1e-9f);
        const float rc = rcp_rn_normal(sc);
        const f32x2 nsc2 = pk2(-sc, -sc), rc2 = pk2(rc, rc), zp2 = pk2(zero[i], zero[i]), sq2 = pk2(sq, sq);
        uint32_t oh[8];
        uint32_t cw[4] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // clamp(v, 0, maxq) + M == clamp(v + M, M, M + maxq): v -> RN(v + M) is monotone and exact at
          // the integer bounds; NaN / inf end at the same bound either way
          float m0, m1;
          upk2(add2(add2(div2_rn_by(x2[i][e], nsc2, rc2), zp2), M2), m0, m1);
          m0 = clamp3(m0, 12582912.0f, m_hi);
          m1 = clamp3(m1, 12582912.0f, m_hi);
          if (kCodes) {
            cw[e >> 1] |= (uint32_t)(__float_as_int(m0) & 0xFF) << (16 * (e & 1));
            cw[e >> 1] |= (uint32_t)(__float_as_int(m1) & 0xFF) << (16 * (e & 1) + 8);
          }
          oh[e] = f2_to_h2(mul2(sq2, sub2(sub2(pk2(m0, m1), M2), zp2)));
        }
        const size_t off = (size_t)(row_t0 + rr + 8 * i) * cols + col0;
        *reinterpret_cast<uint4*>(p.Wq + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<uint4*>(p.Wq + off + 8) = make_uint4(oh[4], oh[5], oh[6], oh[7]);
        if (kCodes) *reinterpret_cast<uint4*>(p.codes + off) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      }
    }

    if (kPack) {
      // ---- packed layout (encode policy: oracle pack_mxq), one code path for 2-bit and 4-bit slots ----
      const float maxq = four ? 15.f : 3.f;
      const uint32_t radix = four ? 16u : 4u;
      const uint32_t horner_bias = 0x4B400000u * (four ? 0x11111111u : 0x00005555u);
      const float m_hi = __fadd_rn(12582912.0f, maxq);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = row_t0 + rr + 8 * i;
        const float lo = fminf(mn[i], 0.f), hi = fmaxf(mx[i], 0.f);
        const float r3 = 0.3333333432674408f;                     // RN(1/3)
        float s = div_rn_by(__fsub_rn(hi, lo), 3.f, r3);
        if (s == 0.f) s = 1.f;
        float smax = s;
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, 4));
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, 8));
        const __half s2h = __float2half_rn(div_rn_by(smax, 3.f, r3));
        const float s2f = __half2float(s2h);
        const float c = clampf(rint_any(fdiv_n(s, s2f)), 1.f, 3.f);
        float S = __fmul_rn(s2f, c);
        float z = clampf(rint_any(fdiv_n(-lo, S)), 0.f, 3.f);
        uint32_t meta = ((uint32_t)z | ((uint32_t)c << 8)) << (2 * k);
        if (four) {
          const float2 pz = s_pack_pool[rr + 8 * i];
          S = pz.x; z = pz.y; meta = 0;
        }
        const float rS = rcp_rn_normal(S);
        const f32x2 nS2 = pk2(-S, -S), rS2 = pk2(rS, rS), zz2 = pk2(z, z);
        // Horner on the raw bits of m = 1.5 * 2^23 + code: word = H - 0x4B400000 * sum radix^e.
        // x2[i][e] holds columns 2e, 2e+1: codes 0..7 come from e = 0..3, codes 8..15 from e = 4..7.
        uint32_t acc_lo = 0, acc_hi = 0;
#pragma unroll
        for (int e = 3; e >= 0; --e) {
          float a0, a1, b0, b1;
          upk2(add2(add2(div2_rn_by(x2[i][e], nS2, rS2), M2), zz2), a0, a1);
          upk2(add2(add2(div2_rn_by(x2[i][e + 4], nS2, rS2), M2), zz2), b0, b1);
          acc_lo = acc_lo * radix + __float_as_uint(clamp3(a1, 12582912.0f, m_hi));
          acc_lo = acc_lo * radix + __float_as_uint(clamp3(a0, 12582912.0f, m_hi));
          acc_hi = acc_hi * radix + __float_as_uint(clamp3(b1, 12582912.0f, m_hi));
          acc_hi = acc_hi * radix + __float_as_uint(clamp3(b0, 12582912.0f, m_hi));
        }
        acc_lo -= horner_bias;
        acc_hi -= horner_bias;
        const uint32_t word = four ? acc_lo : (acc_lo | (acc_hi << 16));
        // every lane stores its own weight word; slot 3 also owns the block's weight_last word
        p.out.weight[(size_t)row * nblk * 4 + (size_t)blk * 4 + k] = (int32_t)word;
        if (four) p.out.weight_last[(size_t)row * nblk + blk] = (int32_t)acc_hi;
        const uint32_t m1 = __shfl_down_sync(0xffffffffu, meta, 1);
        const uint32_t m2 = __shfl_down_sync(0xffffffffu, meta, 2);
        const int chunk = blk >> 6, bp = blk & 63;
        if (k == 0) zs16[((size_t)row * 32 * nchunk + chunk * 32 + (bp & 31)) * 2 + (bp >> 5)] = (uint16_t)(meta | m1 | m2);
        if (k < 3 && (row & 3) == 0) s2o[(size_t)(row >> 2) * nblk * 3 + (size_t)blk * 3 + k] = s2h;
      }
    }
  }
}

// Generic Quantizer on an fp32 matrix, one warp per row, 16 rows per CTA of 512 threads so the
// second-level reduction stays inside the CTA.
__global__ void __launch_bounds__(512) rowquant_kernel(const float* __restrict__ x,
                                                       float* __restrict__ y,
                                                       uint8_t* __restrict__ codes,
                                                       float* __restrict__ scale_out,
                                                       float* __restrict__ zero_out, int rows,
                                                       int cols, int bits, int qq_bits) {
  __shared__ float s_scale[16];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 16 + w;
  const bool ok = row < rows;
  const float maxq = (float)((1 << bits) - 1);
  float mn = INFINITY, mx = -INFINITY;
  if (ok)
    for (int c = lane; c < cols; c += 32) {
      const float v = x[(size_t)row * cols + c];
      mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
  mn = warp_min(mn); mx = warp_max(mx);
  float scale = 1.f, zero = 0.f;
  if (ok) find_params(mn, mx, maxq, scale, zero);
  if (qq_bits > 0) {
    if (lane == 0) s_scale[w] = scale;
    __syncthreads();
    float smin = INFINITY, smax = -INFINITY;
    for (int i = 0; i < 16; ++i) { smin = fminf(smin, s_scale[i]); smax = fmaxf(smax, s_scale[i]); }
    const float qm = (float)((1 << qq_bits) - 1);
    float s2, z2;
    find_params(smin, smax, qm, s2, z2);
    float q = rintf(__fadd_rn(fdiv(scale, fmaxf(s2, 1e-9f)), z2));
    q = fminf(fmaxf(q, 0.f), qm);
    scale = __fmul_rn(s2, __fsub_rn(q, z2));
  }
  if (!ok) return;
  if (lane == 0) {
    if (scale_out) scale_out[row] = scale;
    if (zero_out) zero_out[row] = zero;
  }
  const float sc = fmaxf(scale, 1e-9f);
  for (int c = lane; c < cols; c += 32) {
    const size_t i = (size_t)row * cols + c;
    float q = rintf(__fadd_rn(fdiv(x[i], sc), zero));
    q = fminf(fmaxf(q, 0.f), maxq);
    if (y) y[i] = __fmul_rn(scale, __fsub_rn(q, zero));
    if (codes) codes[i] = (uint8_t)q;
  }
}

}  // namespace mxq

using namespace mxq;

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// workspace: [dead mask: cols bytes][pool_mm: rows float2][pack_pool: rows float2]
static size_t ws_bytes(int64_t rows, int64_t cols) {
  if (rows < 0 || cols < 0) return 0;
  return align16((size_t)cols) + 2 * align16((size_t)rows * sizeof(float2)) + 16;
}

extern "C" size_t mxq_ptq_workspace_bytes(int64_t rows, int64_t cols) { return ws_bytes(rows, cols); }
extern "C" size_t mxq_pack_workspace_bytes(int64_t OC, int64_t IC) { return ws_bytes(OC, IC); }

static int check_packed_ptrs(const mxq_packed_t& p) {
  MXQ_CHECK_PTR(p.weight);
  MXQ_CHECK_PTR(p.weight_last);
  MXQ_CHECK_PTR(p.zeros_and_scales);
  MXQ_CHECK_PTR(p.zeros_2nd);
  if (!p.scales_2nd || !p.scales_4b || !p.zeros_4b) return MXQ_E_NULL;
  return MXQ_OK;
}

template <bool kRef, bool kQuant, bool kPack>
static void launch_tile(const TileParams& tp, unsigned grid, cudaStream_t st) {
  if (tp.codes) ptq_pack_tile_kernel<kRef, kQuant, kPack, true><<<grid, 256, 0, st>>>(tp);
  else ptq_pack_tile_kernel<kRef, kQuant, kPack, false><<<grid, 256, 0, st>>>(tp);
}

template <bool kQuant, bool kPack, int G2>
static int launch_tile16(const Tile16Params& tp, cudaStream_t st) {
  const unsigned grid = (unsigned)(tp.rows / 16) * (unsigned)tp.ksplit;
  const size_t smem = (size_t)tp.cols;
  cudaError_t e = cudaSuccess;
  if (tp.codes) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(ptq_tile16_kernel<kQuant, kPack, true, G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ptq_tile16_kernel<kQuant, kPack, true, G2><<<grid, 256, smem, st>>>(tp);
  } else {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(ptq_tile16_kernel<kQuant, kPack, false, G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ptq_tile16_kernel<kQuant, kPack, false, G2><<<grid, 256, smem, st>>>(tp);
  }
  MXQ_LAUNCH_RESULT();
}

static int run_ptq_pack(const void* W, void* Wq, uint8_t* codes, const float* colstat, int64_t rows,
                        int64_t cols, int group, int low_bits, const uint8_t* group_bits,
                        const mxq_packed_t* packed, void* workspace, size_t workspace_bytes,
                        cudaStream_t st) {
  const bool quant = Wq != nullptr, pack = packed != nullptr;
  if (rows < 0 || cols < 0) return MXQ_E_SHAPE;
  if (rows == 0 || cols == 0) return MXQ_OK;
  MXQ_CHECK_PTR(W);
  if (quant) MXQ_CHECK_PTR(Wq);
  if (colstat && (reinterpret_cast<uintptr_t>(colstat) & 15)) return MXQ_E_ALIGN;
  if (pack) {
    int rc = check_packed_ptrs(*packed);
    if (rc) return rc;
    if (group_bits) return MXQ_E_UNSUPPORTED;   // the packed layout is the reference recipe only
  }
  // group = width of the 2-bit groups: 16 everywhere; 32 / 48 only for the fake-quantizer of the
  // reference recipe (fasterquant's blocksize, clipped to the 48 low columns of a block)
  if (group != 16 && !(group_bits == nullptr && !pack && (group == 32 || group == 48))) return MXQ_E_UNSUPPORTED;
  if (rows % 16 || cols % 64) return MXQ_E_SHAPE;
  if (low_bits < 1 || low_bits > 8) return MXQ_E_SHAPE;
  if (rows > INT32_MAX || cols > (1 << 24)) return MXQ_E_SHAPE;
  const bool ref = group_bits == nullptr;
  if (ref && cols <= 200 * 1024 && !getenv("MXQ_PTQ_ROUND1")) {   // MXQ_PTQ_ROUND1: profiling A/B against the 3-kernel chain
    // one kernel: dead flags, pooled min/max and both quantizers per 16-row tile
    Tile16Params tp{};
    tp.W = (const __half*)W; tp.Wq = (__half*)Wq; tp.codes = codes; tp.colstat = colstat;
    if (pack) tp.out = *packed;
    tp.rows = (int)rows; tp.cols = (int)cols; tp.low_bits = low_bits;
    // enough CTAs for two per SM (register limit): tiles x K parts, each part at least 8 blocks.
    // Measured on B200 (profiles/r2_ptq_pack.txt): 4096 rows 38.7 -> 33.4 us with 2 parts
    {
      const int64_t tiles = rows / 16, nb8 = ceil_div(cols / 64, 8);
      int64_t ks = ceil_div(2 * kNumSMs, tiles);
      if (ks > nb8) ks = nb8;
      if (ks > 8) ks = 8;
      if (const char* e = getenv("MXQ_PTQ_KSPLIT")) ks = atoi(e);
      tp.ksplit = (int)(ks < 1 ? 1 : ks);
    }
    if (quant && pack) return launch_tile16<true, true, 16>(tp, st);
    if (pack) return launch_tile16<false, true, 16>(tp, st);
    if (group == 16) return launch_tile16<true, false, 16>(tp, st);
    if (group == 32) return launch_tile16<true, false, 32>(tp, st);
    return launch_tile16<true, false, 48>(tp, st);
  }
  // explicit per-group masks: dead mask + pooled pre-pass + tile kernel
  MXQ_CHECK_PTR(workspace);
  if (workspace_bytes < ws_bytes(rows, cols)) return MXQ_E_WORKSPACE;
  uint8_t* dead = (uint8_t*)workspace;
  float2* pool_mm = (float2*)((uint8_t*)workspace + align16((size_t)cols));
  float2* pack_pool = (float2*)((uint8_t*)pool_mm + align16((size_t)rows * sizeof(float2)));
  const int nblk = (int)(cols / 64);
  const int nchunk = (nblk + 63) / 64;
  if (pack) {
    cudaMemsetAsync(packed->zeros_2nd, 0, (size_t)(rows / 4) * 32 * nchunk * 4, st);
    if (nblk % 64) cudaMemsetAsync(packed->zeros_and_scales, 0, (size_t)rows * 32 * nchunk * 4, st);
  }
  dead_mask_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, st>>>(colstat, dead, (int)cols);
  const unsigned gridA = (unsigned)ceil_div(rows, 8);
  const __half* w = (const __half*)W;
  __half* s4 = pack ? (__half*)packed->scales_4b : nullptr;
  uint32_t* z4 = pack ? (uint32_t*)packed->zeros_4b : nullptr;
  if (ref && pack) pool_prepass_kernel<true, true><<<gridA, 256, 0, st>>>(w, dead, nullptr, pool_mm, pack_pool, s4, z4, (int)rows, (int)cols);
  else if (ref) pool_prepass_kernel<true, false><<<gridA, 256, 0, st>>>(w, dead, nullptr, pool_mm, pack_pool, s4, z4, (int)rows, (int)cols);
  else pool_prepass_kernel<false, false><<<gridA, 256, 0, st>>>(w, dead, group_bits, pool_mm, pack_pool, s4, z4, (int)rows, (int)cols);
  TileParams tp{};
  tp.W = w; tp.Wq = (__half*)Wq; tp.codes = codes; tp.dead = dead; tp.gbits = group_bits;
  tp.pool_mm = pool_mm; tp.pack_pool = pack_pool;
  if (pack) tp.out = *packed;
  tp.rows = (int)rows; tp.cols = (int)cols; tp.low_bits = low_bits; tp.pool_bits = 4;
  const int64_t units = (rows / 16) * (int64_t)nblk;
  int64_t grid = ceil_div(units, 8);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (!ref) launch_tile<false, true, false>(tp, (unsigned)grid, st);
  else if (quant && pack) launch_tile<true, true, true>(tp, (unsigned)grid, st);
  else if (quant) launch_tile<true, true, false>(tp, (unsigned)grid, st);
  else launch_tile<true, false, true>(tp, (unsigned)grid, st);
  MXQ_LAUNCH_RESULT();
}

extern "C" int mxq_ptq_quant(const void* W, void* Wq, uint8_t* codes, const float* colstat,
                             int64_t rows, int64_t cols, int group, int low_bits,
                             const uint8_t* group_bits, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (!Wq && rows > 0 && cols > 0) return MXQ_E_NULL;
  return run_ptq_pack(W, Wq, codes, colstat, rows, cols, group, low_bits, group_bits, nullptr,
                      workspace, workspace_bytes, as_stream(stream));
}

extern "C" int mxq_pack(const void* W, const float* colstat, int64_t OC, int64_t IC,
                        mxq_packed_t out, void* workspace, size_t workspace_bytes, void* stream) {
  return run_ptq_pack(W, nullptr, nullptr, colstat, OC, IC, 16, 2, nullptr, &out, workspace,
                      workspace_bytes, as_stream(stream));
}

extern "C" int mxq_ptq_quant_pack(const void* W, void* Wq, uint8_t* codes, const float* colstat,
                                  int64_t rows, int64_t cols, mxq_packed_t out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!Wq && rows > 0 && cols > 0) return MXQ_E_NULL;
  return run_ptq_pack(W, Wq, codes, colstat, rows, cols, 16, 2, nullptr, &out, workspace,
                      workspace_bytes, as_stream(stream));
}

extern "C" int mxq_rowquant(const float* x, float* y, uint8_t* codes, float* scale, float* zero,
                            int64_t rows, int64_t cols, int bits, int qq_scale_bits,
                            void* stream) {
  if (rows < 0 || cols < 0) return MXQ_E_SHAPE;
  if (rows == 0 || cols == 0) return MXQ_OK;
  if (!x) return MXQ_E_NULL;
  if (bits < 1 || bits > 8 || qq_scale_bits < 0 || qq_scale_bits > 8) return MXQ_E_SHAPE;
  if (qq_scale_bits > 0 && rows % 16) return MXQ_E_SHAPE;  // quantizer.py:115 reshape(-1, 16)
  rowquant_kernel<<<(unsigned)ceil_div(rows, 16), 512, 0, as_stream(stream)>>>(
      x, y, codes, scale, zero, (int)rows, (int)cols, bits, qq_scale_bits);
  MXQ_LAUNCH_RESULT();
}
