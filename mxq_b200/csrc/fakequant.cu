// Fused MXQ fake-quant forward and clipped-STE backward (sm_100a).
//
// Replaces LLM-QAT/models/utils_quant.py:315-475 (MXAsymQuantizer.forward/backward): the
// reference runs ~2.3k tiny ATen kernels and 5 full-size temporaries per 4096-wide weight; here
// the forward is ONE pass over HBM (read x once, write out once).
//
// Forward data path:  HBM --cp.async.bulk (TMA 1-D, mbarrier tx-count)--> shared-memory ring of
// whole rows --LDS.128--> registers --STG.128 (streaming)--> HBM.  Rows must be CTA-resident
// because all 4-bit columns of a row share one min/max (utils_quant.py:347,368-377).
//   pass 1 (per row)  : min/max over the pooled (4-bit) chunks only, warp shuffle + 8 partials
//   pass 2 (per chunk): 16-byte chunk per lane, group min/max by xor-shuffles across the
//                       lanes of the group, quantize/dequantize op by op in the tensor dtype.
// Bit-exactness: IEEE fp32 ops via __f*_rn intrinsics (never contracted), division by a
// correctly-rounded reciprocal + one Markstein correction (common.cuh), round-half-even by the
// magic-constant add.  Each intermediate is rounded to the tensor dtype like the reference's
// separate ATen kernels do.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace mxq {

struct FQParams {
  const uint8_t* x;
  uint8_t* out;
  uint8_t* codes;
  const uint8_t* group_bits;
  int rows, cols;
  int cpr;          // 16-byte chunks per row
  int lpg;          // lanes (chunks) per group, power of two <= 32
  int lpg_shift;
  int low_bits, pool_bits;
  int tw;           // warps cooperating on one row (power of two)
  int tw_shift;
  int teams;        // rows per stage = 8 / tw
  int stages;
  int row_bytes;
  int stage_bytes;
  int num_sets;     // ceil(rows / teams)
};

constexpr int kFQThreads = 256;

// Template switches (all compile-time so that the hot loop carries no dead work):
//   kRef    reference recipe {low,low,low,pool} generated arithmetically (no mask table)
//   kFast16 fp16/bf16 tensors with the reference bit-widths (2-bit groups, 4-bit pool): the
//           per-element chain runs on packed f16x2/bf16x2 instructions (common.cuh P16), only
//           the division is done in fp32.  Otherwise: fp32 math + explicit rounding to T.
//   LPG     lanes (16-byte chunks) per group if known at compile time, 0 = runtime (p.lpg)
//   kCodes  also emit the integer codes
template <typename T, bool kRef, bool kFast16, int LPG, bool kCodes>
__global__ void __launch_bounds__(kFQThreads) fakequant_fwd_kernel(const FQParams p) {
  using D = DT<T>;
  using P = P16<typename std::conditional<kFast16, T, __half>::type>;
  // the reference's eager CPU kernels cast the python scalar 1e-8 to the tensor dtype first
  // (bf16: 1.0012e-8, fp16: 0), then add in fp32 and round
  const float kEps = D::rnd(1e-8f);
  constexpr int EPC = D::EPC;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* bufs = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  float2* part = reinterpret_cast<float2*>(bars + 8);
  uint8_t* gtab = reinterpret_cast<uint8_t*>(part + 8);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = warp >> p.tw_shift, tw = warp & (p.tw - 1);
  const int tthreads = p.tw * 32;
  const int lpg = LPG ? LPG : p.lpg;
  const int lpg_shift = LPG ? (LPG == 1 ? 0 : LPG == 2 ? 1 : LPG == 4 ? 2 : LPG == 8 ? 3 : LPG == 16 ? 4 : 5)
                            : p.lpg_shift;

  if (!kRef) {
    const int ng = p.cols / (lpg * EPC);
    for (int g = tid; g < ng; g += kFQThreads) gtab[g] = p.group_bits[g];
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int first = blockIdx.x, step = gridDim.x;
  const int nmine = first < p.num_sets ? (p.num_sets - first + step - 1) / step : 0;

  auto issue = [&](int it) {
    const int set = first + it * step;
    const int stage = it % p.stages;
    const int row0 = set * p.teams;
    const int nrows = min(p.teams, p.rows - row0);
    const uint32_t bytes = (uint32_t)nrows * (uint32_t)p.row_bytes;
    mbar_arrive_expect_tx(&bars[stage], bytes);
    bulk_g2s(bufs + (size_t)stage * p.stage_bytes, p.x + (size_t)row0 * p.row_bytes, bytes,
             &bars[stage]);
  };
  if (tid == 0) {
    const int n0 = min(p.stages, nmine);
    for (int s = 0; s < n0; ++s) issue(s);
  }

  const float s_low = (float)((1 << p.low_bits) - 1);
  const float s_pool = (float)((1 << p.pool_bits) - 1);
  const float rs_low = __frcp_rn(s_low), rs_pool = __frcp_rn(s_pool);

  for (int it = 0; it < nmine; ++it) {
    const int stage = it % p.stages;
    const uint32_t parity = (it / p.stages) & 1;
    const int row0 = (first + it * step) * p.teams;
    const int nrows = min(p.teams, p.rows - row0);
    const bool active = team < nrows;
    const uint8_t* rowp = bufs + (size_t)stage * p.stage_bytes + (size_t)team * p.row_bytes;
    mbar_wait(&bars[stage], parity);

    // ---- pass 1: per-row min/max over the pooled chunks -----------------------------------
    float pmin = INFINITY, pmax = -INFINITY;
    if (active) {
      if (kRef && kFast16) {
        const int npc = p.cpr >> 2;  // one group in four is pooled
        uint32_t mn2 = P::kPosInfNegInf & 0xFFFFu, mx2 = P::kPosInfNegInf >> 16;
        mn2 |= mn2 << 16; mx2 |= mx2 << 16;
        for (int m = tw * 32 + lane; m < npc; m += tthreads) {
          const int c = ((((m >> lpg_shift) << 2) + 3) << lpg_shift) + (m & (lpg - 1));
          const uint4 ch = *reinterpret_cast<const uint4*>(rowp + c * 16);
          mn2 = P::vmin(P::vmin(mn2, P::vmin(ch.x, ch.y)), P::vmin(ch.z, ch.w));
          mx2 = P::vmax(P::vmax(mx2, P::vmax(ch.x, ch.y)), P::vmax(ch.z, ch.w));
        }
        pmin = fminf(P::lo(mn2), P::hi(mn2));
        pmax = fmaxf(P::lo(mx2), P::hi(mx2));
      } else if (kRef) {
        const int npc = p.cpr >> 2;
        for (int m = tw * 32 + lane; m < npc; m += tthreads) {
          const int c = ((((m >> lpg_shift) << 2) + 3) << lpg_shift) + (m & (lpg - 1));
          const uint4 ch = *reinterpret_cast<const uint4*>(rowp + c * 16);
          float f[EPC];
          D::unpack(ch, f);
#pragma unroll
          for (int e = 0; e < EPC; ++e) { pmin = fminf(pmin, f[e]); pmax = fmaxf(pmax, f[e]); }
        }
      } else {
        for (int c = tw * 32 + lane; c < p.cpr; c += tthreads) {
          if (gtab[c >> lpg_shift] & MXQ_POOL_FLAG) {
            const uint4 ch = *reinterpret_cast<const uint4*>(rowp + c * 16);
            float f[EPC];
            D::unpack(ch, f);
#pragma unroll
            for (int e = 0; e < EPC; ++e) { pmin = fminf(pmin, f[e]); pmax = fmaxf(pmax, f[e]); }
          }
        }
      }
    }
    pmin = warp_min(pmin);
    pmax = warp_max(pmax);
    if (lane == 0) part[warp] = make_float2(pmin, pmax);
    __syncthreads();
    pmin = INFINITY; pmax = -INFINITY;
    for (int i = 0; i < p.tw; ++i) {
      const float2 v = part[(team << p.tw_shift) + i];
      pmin = fminf(pmin, v.x); pmax = fmaxf(pmax, v.y);
    }
    // utils_quant.py:369-377: fp32 subtract, cast to the tensor dtype on assignment (:383)
    const float a_pool = D::rnd(__fadd_rn(D::rnd(__fsub_rn(pmax, pmin)), kEps));
    const float b_pool = pmin;
    const float r_pool = __frcp_rn(a_pool);

    // ---- pass 2: quantize / dequantize -----------------------------------------------------
    if (active) {
      uint8_t* orow = p.out + (size_t)(row0 + team) * p.row_bytes;
      uint8_t* crow = kCodes ? p.codes + (size_t)(row0 + team) * p.cols : nullptr;
      if (kFast16) {
        const uint32_t a2_pool = P::pack(a_pool, a_pool), b2_pool = P::pack(b_pool, b_pool);
        for (int c0 = tw * 32; c0 < p.cpr; c0 += tthreads) {
          const int c = c0 + lane;
          const bool valid = c < p.cpr;
          uint4 ch = make_uint4(0, 0, 0, 0);
          uint32_t mm = P::kPosInfNegInf;            // lo half: running min, hi half: running max
          if (valid) {
            ch = *reinterpret_cast<const uint4*>(rowp + c * 16);
            uint32_t mn = P::vmin(P::vmin(ch.x, ch.y), P::vmin(ch.z, ch.w));
            uint32_t mx = P::vmax(P::vmax(ch.x, ch.y), P::vmax(ch.z, ch.w));
            mn = P::vmin(mn, prmt_b32(mn, mn, 0x1032));
            mx = P::vmax(mx, prmt_b32(mx, mx, 0x1032));
            mm = prmt_b32(mn, mx, 0x5410);
          }
#pragma unroll
          for (int o = lpg >> 1; o > 0; o >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mm, o);
            mm = prmt_b32(P::vmin(mm, other), P::vmax(mm, other), 0x7610);
          }
          if (!valid) continue;
          const int g = c >> lpg_shift;
          const bool pooled = kRef ? ((g & 3) == 3) : ((gtab[g] & MXQ_POOL_FLAG) != 0);
          uint32_t b2 = prmt_b32(mm, mm, 0x1010);
          const uint32_t alpha2 = P::sub(prmt_b32(mm, mm, 0x3232), b2);      // max - min (:358-361)
          const float ax = __fadd_rn(P::lo(alpha2), kEps);                  // alpha + 1e-8 (:456)
          uint32_t a2 = P::pack(ax, ax);                                     // rounded to the dtype
          float a = P::lo(a2);
          float r = rcp_rn_normal(a);
          uint32_t s2 = P::kS3, ch2 = P::kC3, cl2 = 0u;
          if (pooled) {
            a = a_pool; r = r_pool; a2 = a2_pool; b2 = b2_pool;
            s2 = P::kS15; ch2 = P::kC15hi; cl2 = P::kC15lo;
          }
          const uint32_t w[4] = {ch.x, ch.y, ch.z, ch.w};
          const f32x2 na_2 = pk2(-a, -a), r_2 = pk2(r, r);
          uint32_t o[4], cw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t t2 = P::sub(w[i], b2);                            // x - beta
            float n_lo, n_hi;
            upk2(div2_rn_by(pk2(P::lo(t2), P::hi(t2)), na_2, r_2), n_lo, n_hi);
            const uint32_t n2 = P::pack(n_lo, n_hi);
            const uint32_t m2 = P::add(P::mul(n2, s2), P::kMagic);           // round-half-even
            const uint32_t q2 = P::sub(m2, P::kMagic);
            const uint32_t v2 = P::fma(q2, ch2, P::mul(q2, cl2));            // == RN16(RN32(q / s))
            o[i] = P::add(P::mul(v2, a2), b2);
            cw[i] = m2 & P::kCodeMask;
          }
          st_stream(orow + c * 16, make_uint4(o[0], o[1], o[2], o[3]));
          if (kCodes) {
            *reinterpret_cast<uint2*>(crow + c * 8) =
                make_uint2(prmt_b32(cw[0], cw[1], 0x6420), prmt_b32(cw[2], cw[3], 0x6420));
          }
        }
      } else {
        for (int c0 = tw * 32; c0 < p.cpr; c0 += tthreads) {
          const int c = c0 + lane;
          const bool valid = c < p.cpr;
          float f[EPC];
          float lmin = INFINITY, lmax = -INFINITY;
          if (valid) {
            const uint4 ch = *reinterpret_cast<const uint4*>(rowp + c * 16);
            D::unpack(ch, f);
#pragma unroll
            for (int e = 0; e < EPC; ++e) { lmin = fminf(lmin, f[e]); lmax = fmaxf(lmax, f[e]); }
          }
#pragma unroll
          for (int o = lpg >> 1; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
          }
          if (!valid) continue;
          const int g = c >> lpg_shift;
          const bool pooled = kRef ? ((g & 3) == 3) : ((gtab[g] & MXQ_POOL_FLAG) != 0);
          float a = D::rnd(__fadd_rn(D::rnd(__fsub_rn(lmax, lmin)), kEps));  // alpha + 1e-8 (:456)
          float b = lmin;
          float r = rcp_rn_normal(a);
          float s = s_low, rs = rs_low;
          if (pooled) { a = a_pool; b = b_pool; r = r_pool; s = s_pool; rs = rs_pool; }
          if (sizeof(T) == 4) {
            // fp32: the chain on packed FADD2/FMUL2/FFMA2 (two elements per instruction)
            const f32x2 b_2 = pk2(b, b), a_2 = pk2(a, a), na_2 = pk2(-a, -a), r_2 = pk2(r, r);
            const f32x2 s_2 = pk2(s, s), ns_2 = pk2(-s, -s), rs_2 = pk2(rs, rs);
            const f32x2 nm_2 = pk2(-12582912.0f, -12582912.0f);
            float o4[4];
            uint32_t cw4 = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const f32x2 t = sub2(pk2(f[2 * h], f[2 * h + 1]), b_2);
              const f32x2 u = mul2(div2_rn_by(t, na_2, r_2), s_2);
              float u0, u1;
              upk2(u, u0, u1);
              const float m0 = __fadd_rn(u0, 12582912.0f), m1 = __fadd_rn(u1, 12582912.0f);
              const f32x2 q = add2(pk2(m0, m1), nm_2);
              const f32x2 w = mul2(div2_rn_by(q, ns_2, rs_2), a_2);
              float w0, w1;
              upk2(w, w0, w1);
              o4[2 * h] = __fadd_rn(w0, b);
              o4[2 * h + 1] = __fadd_rn(w1, b);
              if (kCodes)
                cw4 |= ((__float_as_uint(m0) & 0xFFu) | ((__float_as_uint(m1) & 0xFFu) << 8)) << (16 * h);
            }
            st_stream(orow + c * 16, make_uint4(__float_as_uint(o4[0]), __float_as_uint(o4[1]),
                                                __float_as_uint(o4[2]), __float_as_uint(o4[3])));
            if (kCodes) *reinterpret_cast<uint32_t*>(crow + c * 4) = cw4;
            continue;
          }
          float o[EPC];
          uint32_t cw[EPC / 4] = {};
#pragma unroll
          for (int e = 0; e < EPC; ++e) {
            float t = D::rnd(__fsub_rn(f[e], b));
            t = D::rnd(div_rn_by(t, a, r));          // input_normalized (:456)
            t = D::rnd(__fmul_rn(t, s));
            int qi;
            const float q = rint_magic(t, qi);       // torch.round (:458)
            t = D::rnd(div_rn_by(q, s, rs));         // .div(s)
            t = D::rnd(__fmul_rn(t, a));
            o[e] = D::rnd(__fadd_rn(t, b));          // (:460)
            if (kCodes) cw[e >> 2] |= (uint32_t)(qi & 0xFF) << (8 * (e & 3));
          }
          st_stream(orow + c * 16, D::pack(o));
          if (kCodes) {
            if (EPC == 4) {
              *reinterpret_cast<uint32_t*>(crow + c * 4) = cw[0];
            } else {
              *reinterpret_cast<uint2*>(crow + c * 8) = make_uint2(cw[0], cw[EPC / 4 - 1]);
            }
          }
        }
      }
    }
    __syncthreads();  // every warp is done reading this stage (and `part`)
    if (tid == 0 && it + p.stages < nmine) issue(it + p.stages);
  }
}


// -------------------------------------------------------------------------------------------
// Row-resident variant for the reference recipe ({low, low, low, pool 4} groups of G = 16 columns, or
// of 128 columns for BASELINE's "group 128"): one CTA per
// row, the row's 16-byte chunks stay in registers between the pooled min/max and the quantize
// pass, so the kernel is a single load -> reduce -> compute -> store stream with nothing staged
// in shared memory but the per-warp partial statistics.  Measured faster than the TMA ring above
// on B200 for rows of up to 3072 chunks (profiles/; 512-thread instantiations carry 70B rows of up to 6144 chunks): more independent loads in flight per SM and
// no per-stage block barrier.
// -------------------------------------------------------------------------------------------
// Up to kFQMaxTensors weights of the same row length in ONE launch (QAT quantizes q/k/v/o, then gate/up,
// of every decoder layer back to back: 672 launches of 15-25 us per step, each paying its own ramp and
// tail; profiles/): grid = total rows, a CTA finds its tensor in the prefix table.
constexpr int kFQMaxTensors = 8;
struct FQRowTable {
  const uint4* x[kFQMaxTensors];
  uint4* out[kFQMaxTensors];
  int row_end[kFQMaxTensors];     // exclusive prefix sums of the row counts
  int n;
  const uint8_t* pooled_mask;     // optional group_bits: only its POOL flag is read (which groups are pooled)
};

template <typename T, int THREADS, int NC, int G = 16, bool kMask = false>
__global__ void __launch_bounds__(THREADS) fakequant_row_kernel(const __grid_constant__ FQRowTable tab, int cpr,
                                                                int low_bits) {
  using D = DT<T>;
  constexpr bool k16 = sizeof(T) == 2;
  using P = P16<typename std::conditional<k16, T, __half>::type>;
  constexpr int EPC = D::EPC;
  constexpr int LPG = G / EPC;                        // lanes (chunks) per G-column group: 2 / 4 at the
  constexpr int LPG_SHIFT = LPG == 2 ? 1 : LPG == 4 ? 2 : LPG == 16 ? 4 : 5;   // reference G = 16, 16 / 32 at G = 128
  static_assert(LPG == 2 || LPG == 4 || LPG == 16 || LPG == 32, "group of 16 or 128 columns");
  __shared__ float2 part[THREADS / 32];
  const float kEps = D::rnd(1e-8f);
  int ti = 0, row = (int)blockIdx.x;
#pragma unroll 1
  while (ti + 1 < tab.n && row >= tab.row_end[ti]) ++ti;
  if (ti > 0) row -= tab.row_end[ti - 1];
  const uint4* xr = tab.x[ti] + (size_t)row * cpr;
  uint4* orow = tab.out[ti] + (size_t)row * cpr;
  const int lane = threadIdx.x & 31;
  // which groups share the row's pooled statistic: every fourth one (the reference's positional recipe,
  // utils_quant.py:349-353) or the ones flagged in the mask (importance-driven allocation)
  auto is_pooled = [&](int c) -> bool {
    if (kMask) return (__ldg(tab.pooled_mask + (c >> LPG_SHIFT)) & MXQ_POOL_FLAG) != 0;
    return ((c >> LPG_SHIFT) & 3) == 3;
  };

  uint4 ch[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = i * THREADS + threadIdx.x;
    ch[i] = c < cpr ? ld_stream(xr + c) : make_uint4(0, 0, 0, 0);
  }
  // ---- pooled min/max of the row (every fourth group) ----------------------------------------
  float pmin = INFINITY, pmax = -INFINITY;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = i * THREADS + threadIdx.x;
    if (c < cpr && is_pooled(c)) {
      float f[EPC];
      D::unpack(ch[i], f);
#pragma unroll
      for (int e = 0; e < EPC; ++e) { pmin = fminf(pmin, f[e]); pmax = fmaxf(pmax, f[e]); }
    }
  }
  pmin = warp_min(pmin); pmax = warp_max(pmax);
  if (lane == 0) part[threadIdx.x >> 5] = make_float2(pmin, pmax);
  __syncthreads();
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) { pmin = fminf(pmin, part[w].x); pmax = fmaxf(pmax, part[w].y); }
  // utils_quant.py:369-377: fp32 subtract, cast to the tensor dtype on assignment (:383)
  const float a_pool = D::rnd(__fadd_rn(D::rnd(__fsub_rn(pmax, pmin)), kEps));
  const float b_pool = pmin;
  const float r_pool = __frcp_rn(a_pool);
  const float s_low = (float)((1 << low_bits) - 1);
  const float rs_low = __frcp_rn(s_low), rs_pool = __frcp_rn(15.0f);

  // ---- quantize / dequantize -------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = i * THREADS + threadIdx.x;
    const bool valid = c < cpr;          // cpr % LPG == 0 and THREADS % LPG == 0: groups never straddle
    const bool pooled = valid && is_pooled(c);
    if constexpr (k16) {
      uint32_t mn = P::vmin(P::vmin(ch[i].x, ch[i].y), P::vmin(ch[i].z, ch[i].w));
      uint32_t mx = P::vmax(P::vmax(ch[i].x, ch[i].y), P::vmax(ch[i].z, ch[i].w));
      mn = P::vmin(mn, prmt_b32(mn, mn, 0x1032));
      mx = P::vmax(mx, prmt_b32(mx, mx, 0x1032));
      uint32_t mm = valid ? prmt_b32(mn, mx, 0x5410) : P::kPosInfNegInf;   // lo: min, hi: max
#pragma unroll
      for (int o = LPG >> 1; o > 0; o >>= 1) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, mm, o);
        mm = prmt_b32(P::vmin(mm, other), P::vmax(mm, other), 0x7610);
      }
      if (!valid) continue;
      uint32_t b2 = prmt_b32(mm, mm, 0x1010);
      const uint32_t alpha2 = P::sub(prmt_b32(mm, mm, 0x3232), b2);        // max - min (:358-361)
      const float ax = __fadd_rn(P::lo(alpha2), kEps);                     // alpha + 1e-8 (:456)
      uint32_t a2 = P::pack(ax, ax);
      float a = P::lo(a2);
      float r = rcp_rn_normal(a);
      uint32_t s2 = P::kS3, ch2 = P::kC3, cl2 = 0u;
      if (pooled) {
        a = a_pool; r = r_pool; a2 = P::pack(a_pool, a_pool); b2 = P::pack(b_pool, b_pool);
        s2 = P::kS15; ch2 = P::kC15hi; cl2 = P::kC15lo;
      }
      const uint32_t w[4] = {ch[i].x, ch[i].y, ch[i].z, ch[i].w};
      const f32x2 na_2 = pk2(-a, -a), r_2 = pk2(r, r);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t t2 = P::sub(w[j], b2);                              // x - beta
        float n_lo, n_hi;
        upk2(div2_rn_by(pk2(P::lo(t2), P::hi(t2)), na_2, r_2), n_lo, n_hi);
        const uint32_t n2 = P::pack(n_lo, n_hi);
        const uint32_t m2 = P::add(P::mul(n2, s2), P::kMagic);             // round-half-even
        const uint32_t q2 = P::sub(m2, P::kMagic);
        const uint32_t v2 = P::fma(q2, ch2, P::mul(q2, cl2));              // == RN16(RN32(q / s))
        o[j] = P::add(P::mul(v2, a2), b2);
      }
      st_stream(orow + c, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
      float f[EPC];
      D::unpack(ch[i], f);
      float lmin = INFINITY, lmax = -INFINITY;
      if (valid) {
#pragma unroll
        for (int e = 0; e < EPC; ++e) { lmin = fminf(lmin, f[e]); lmax = fmaxf(lmax, f[e]); }
      }
#pragma unroll
      for (int o = LPG >> 1; o > 0; o >>= 1) {
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      }
      if (!valid) continue;
      float a = D::rnd(__fadd_rn(D::rnd(__fsub_rn(lmax, lmin)), kEps));    // alpha + 1e-8 (:456)
      float b = lmin;
      float r = rcp_rn_normal(a);
      float sq = s_low, rs = rs_low;
      if (pooled) { a = a_pool; b = b_pool; r = r_pool; sq = 15.0f; rs = rs_pool; }
      const f32x2 b_2 = pk2(b, b), a_2 = pk2(a, a), na_2 = pk2(-a, -a), r_2 = pk2(r, r);
      const f32x2 s_2 = pk2(sq, sq), ns_2 = pk2(-sq, -sq), rs_2 = pk2(rs, rs);
      const f32x2 nm_2 = pk2(-12582912.0f, -12582912.0f);
      float o4[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const f32x2 t = sub2(pk2(f[2 * h], f[2 * h + 1]), b_2);
        const f32x2 u = mul2(div2_rn_by(t, na_2, r_2), s_2);
        float u0, u1;
        upk2(u, u0, u1);
        const float m0 = __fadd_rn(u0, 12582912.0f), m1 = __fadd_rn(u1, 12582912.0f);
        const f32x2 q = add2(pk2(m0, m1), nm_2);
        const f32x2 wv = mul2(div2_rn_by(q, ns_2, rs_2), a_2);
        float w0, w1;
        upk2(wv, w0, w1);
        o4[2 * h] = __fadd_rn(w0, b);
        o4[2 * h + 1] = __fadd_rn(w1, b);
      }
      st_stream(orow + c, make_uint4(__float_as_uint(o4[0]), __float_as_uint(o4[1]),
                                     __float_as_uint(o4[2]), __float_as_uint(o4[3])));
    }
  }
}

template <typename T, int G, bool kMask>
static bool launch_fq_row_tab(const FQRowTable& tab, int c, int low_bits, cudaStream_t st) {
  const unsigned g = (unsigned)tab.row_end[tab.n - 1];
  if (c <= 256) fakequant_row_kernel<T, 128, 2, G, kMask><<<g, 128, 0, st>>>(tab, c, low_bits);
  else if (c <= 512) fakequant_row_kernel<T, 128, 4, G, kMask><<<g, 128, 0, st>>>(tab, c, low_bits);
  else if (c <= 1024) fakequant_row_kernel<T, 256, 4, G, kMask><<<g, 256, 0, st>>>(tab, c, low_bits);
  else if (c <= 1536) fakequant_row_kernel<T, 256, 6, G, kMask><<<g, 256, 0, st>>>(tab, c, low_bits);
  else if (c <= 2048) fakequant_row_kernel<T, 256, 8, G, kMask><<<g, 256, 0, st>>>(tab, c, low_bits);
  else if (c <= 3072) fakequant_row_kernel<T, 256, 12, G, kMask><<<g, 256, 0, st>>>(tab, c, low_bits);
  // 70B rows (down_proj: 28672 columns = 3584 bf16 chunks) on 512 threads
  else if (c <= 4096) fakequant_row_kernel<T, 512, 8, G, kMask><<<g, 512, 0, st>>>(tab, c, low_bits);
  else if (c <= 6144) fakequant_row_kernel<T, 512, 12, G, kMask><<<g, 512, 0, st>>>(tab, c, low_bits);
  else return false;
  return true;
}

template <typename T, int G = 16>
static bool launch_fq_row(const FQParams& p, cudaStream_t st) {
  FQRowTable tab{};
  tab.x[0] = (const uint4*)p.x; tab.out[0] = (uint4*)p.out; tab.row_end[0] = p.rows; tab.n = 1;
  return launch_fq_row_tab<T, G, false>(tab, p.cpr, p.low_bits, st);
}

// dtype x {16, 128} x {positional, masked}
static bool launch_fq_row_any(const FQRowTable& tab, int dtype, int group, int cpr, int low_bits, cudaStream_t st) {
  const bool m = tab.pooled_mask != nullptr;
#define MXQ_FQ_ROW(T)                                                                                       \
  (group == 16 ? (m ? launch_fq_row_tab<T, 16, true>(tab, cpr, low_bits, st) : launch_fq_row_tab<T, 16, false>(tab, cpr, low_bits, st)) \
               : (m ? launch_fq_row_tab<T, 128, true>(tab, cpr, low_bits, st) : launch_fq_row_tab<T, 128, false>(tab, cpr, low_bits, st)))
  switch (dtype) {
    case MXQ_F32: return MXQ_FQ_ROW(float);
    case MXQ_F16: return MXQ_FQ_ROW(__half);
    default: return MXQ_FQ_ROW(__nv_bfloat16);
  }
#undef MXQ_FQ_ROW
}

// -------------------------------------------------------------------------------------------
// STE backward: pure streaming select, 3 tensors, 128-bit accesses, 4 chunks in flight / thread
// -------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ste_bwd_kernel(const uint4* __restrict__ g,
                                                      const uint4* __restrict__ x,
                                                      uint4* __restrict__ gi, int64_t nchunks,
                                                      float lo, float hi) {
  using D = DT<T>;
  constexpr int EPC = D::EPC;
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < nchunks; i += U * stride) {
    uint4 gv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { gv[u] = ld_stream(g + i + u * stride); xv[u] = ld_stream(x + i + u * stride); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float gf[EPC], xf[EPC];
      D::unpack(gv[u], gf);
      D::unpack(xv[u], xf);
#pragma unroll
      for (int e = 0; e < EPC; ++e) gf[e] = (xf[e] >= hi || xf[e] <= lo) ? 0.0f : gf[e];
      st_stream(gi + i + u * stride, D::pack(gf));
    }
  }
  for (; i < nchunks; i += stride) {
    float gf[EPC], xf[EPC];
    D::unpack(ld_stream(g + i), gf);
    D::unpack(ld_stream(x + i), xf);
#pragma unroll
    for (int e = 0; e < EPC; ++e) gf[e] = (xf[e] >= hi || xf[e] <= lo) ? 0.0f : gf[e];
    st_stream(gi + i, D::pack(gf));
  }
}

template <typename T>
__global__ void ste_bwd_tail_kernel(const T* g, const T* x, T* gi, int64_t start, int64_t n,
                                    float lo, float hi) {
  int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float xf = (float)x[i];
    gi[i] = (xf >= hi || xf <= lo) ? (T)0.0f : g[i];
  }
}

template <typename T, bool kRef, bool kFast16, int LPG>
static int launch_fq3(const FQParams& p, int smem, int grid, cudaStream_t st) {
  auto k = p.codes ? fakequant_fwd_kernel<T, kRef, kFast16, LPG, true>
                   : fakequant_fwd_kernel<T, kRef, kFast16, LPG, false>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  k<<<grid, kFQThreads, smem, st>>>(p);
  MXQ_LAUNCH_RESULT();
}

template <typename T>
static int launch_fq(const FQParams& p, bool ref, int smem, int grid, cudaStream_t st) {
  constexpr bool k16 = sizeof(T) == 2;
  constexpr int kLpg16 = 16 / DT<T>::EPC;   // lanes per group at the reference group size 16
  const bool fast = k16 && p.low_bits == 2 && p.pool_bits == 4;
  const bool g16 = p.lpg == kLpg16;
  if (ref) {
    if (fast) return g16 ? launch_fq3<T, true, k16, kLpg16>(p, smem, grid, st) : launch_fq3<T, true, k16, 0>(p, smem, grid, st);
    return g16 ? launch_fq3<T, true, false, kLpg16>(p, smem, grid, st) : launch_fq3<T, true, false, 0>(p, smem, grid, st);
  }
  if (fast) return launch_fq3<T, false, k16, 0>(p, smem, grid, st);
  return launch_fq3<T, false, false, 0>(p, smem, grid, st);
}

template <typename T>
static int launch_ste(const void* g, const void* x, void* gi, int64_t n, float lo, float hi,
                      cudaStream_t st) {
  constexpr int EPC = DT<T>::EPC;
  const int64_t nchunks = n / EPC;
  if (nchunks > 0) {
    int64_t blocks = ceil_div(nchunks, 256 * 4);
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ste_bwd_kernel<T><<<(int)blocks, 256, 0, st>>>((const uint4*)g, (const uint4*)x, (uint4*)gi,
                                                   nchunks, lo, hi);
  }
  const int64_t done = nchunks * EPC;
  if (done < n) {
    ste_bwd_tail_kernel<T><<<1, 32, 0, st>>>((const T*)g, (const T*)x, (T*)gi, done, n, lo, hi);
  }
  MXQ_LAUNCH_RESULT();
}

}  // namespace mxq

using namespace mxq;

extern "C" int mxq_fakequant_fwd(const void* x, void* out, uint8_t* codes, int64_t rows,
                                 int64_t cols, int dtype, int group, int low_bits,
                                 const uint8_t* group_bits, void* stream) {
  if (rows < 0 || cols < 0) return MXQ_E_SHAPE;
  if (rows == 0 || cols == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(out);
  if (dtype != MXQ_F32 && dtype != MXQ_F16 && dtype != MXQ_BF16) return MXQ_E_DTYPE;
  const int esize = dtype == MXQ_F32 ? 4 : 2;
  const int epc = 16 / esize;
  if (group < epc || (group & (group - 1)) || group / epc > 32) return MXQ_E_SHAPE;
  if (cols % group) return MXQ_E_SHAPE;
  if (!group_bits && cols % (4 * group)) return MXQ_E_SHAPE;
  if (low_bits < 1 || low_bits > 8) return MXQ_E_SHAPE;
  if (rows > INT32_MAX || cols > (1 << 24)) return MXQ_E_SHAPE;

  FQParams p{};
  p.x = (const uint8_t*)x; p.out = (uint8_t*)out; p.codes = codes; p.group_bits = group_bits;
  p.rows = (int)rows; p.cols = (int)cols;
  p.row_bytes = (int)cols * esize;
  p.cpr = p.row_bytes / 16;
  p.lpg = group / epc;
  p.lpg_shift = 0;
  while ((1 << p.lpg_shift) < p.lpg) ++p.lpg_shift;
  p.low_bits = low_bits; p.pool_bits = 4;
  // rows per stage: measured best on B200 (profiles/sweep_fq.py): 32 KB sets for fp32, 16 KB for
  // 16-bit tensors, 2 stages, up to 4 CTAs per SM
  int team_bytes = esize == 4 ? 32768 : 16384;
  if (const char* e = getenv("MXQ_FQ_TEAM_BYTES")) team_bytes = atoi(e);
  int teams = 8;
  while (teams > 1 && (int64_t)teams * p.row_bytes > team_bytes) teams >>= 1;
  p.teams = teams; p.tw = 8 / teams;
  p.tw_shift = 0;
  while ((1 << p.tw_shift) < p.tw) ++p.tw_shift;
  p.stage_bytes = teams * p.row_bytes;
  const int ng = (int)(cols / group);
  const int tail = 8 * 8 + 8 * 8 + ((ng + 15) & ~15);
  const int max_smem = 227 * 1024;
  if (p.stage_bytes + tail > max_smem) return MXQ_E_UNSUPPORTED;
  // two stages per CTA (one being consumed, one in flight) and as many CTAs per SM as registers
  // allow: the kernel is issue-bound, so resident warps matter more than pipeline depth
  int stages = 2;
  if (const char* e = getenv("MXQ_FQ_STAGES")) stages = atoi(e);   // tuning knob (profiles/sweep_fq.py)
  while (stages > 1 && stages * p.stage_bytes + tail > max_smem) --stages;
  p.stages = stages;
  p.num_sets = (int)ceil_div(rows, teams);
  const int smem = stages * p.stage_bytes + tail;
  int bps = (228 * 1024) / (smem + 1024);
  if (bps < 1) bps = 1;
  if (bps > 4) bps = 4;
  if (const char* e = getenv("MXQ_FQ_BPS")) { const int v = atoi(e); if (v >= 1 && v < bps) bps = v; }
  int grid = kNumSMs * bps;
  if (grid > p.num_sets) grid = p.num_sets;
  cudaStream_t st = as_stream(stream);
  const bool ref = group_bits == nullptr;
  // reference recipe at group 16 without code output: the row-resident kernel (rows of up to 6144
  // chunks; the packed 16-bit arithmetic exists for 2-bit groups only).  MXQ_FQ_RING forces the ring.
  // The same kernel instantiated for groups of 128 columns serves BASELINE's "group 128" (the
  // positional recipe over 512-column blocks): bit-identical to the mask-driven ring kernel on B200
  // (fp32 / bf16 / fp16, NaN payloads included) and 22.7 / 14.7 us instead of 32.6 / 21.3 us for a
  // 4096^2 fp32 / bf16 weight.  MXQ_FQ_ROW_G128=0 restores the ring.
  const char* e128 = getenv("MXQ_FQ_ROW_G128");
  const bool row16 = group == 16, row128 = group == 128 && !(e128 && atoi(e128) == 0);
  if (ref && (row16 || row128) && !codes && p.cpr <= 6144 && rows >= kNumSMs && (esize == 4 || low_bits == 2) &&
      !getenv("MXQ_FQ_RING")) {
    bool ok = false;
    switch (dtype) {
      case MXQ_F32: ok = row16 ? launch_fq_row<float>(p, st) : launch_fq_row<float, 128>(p, st); break;
      case MXQ_F16: ok = row16 ? launch_fq_row<__half>(p, st) : launch_fq_row<__half, 128>(p, st); break;
      default: ok = row16 ? launch_fq_row<__nv_bfloat16>(p, st) : launch_fq_row<__nv_bfloat16, 128>(p, st); break;
    }
    if (ok) MXQ_LAUNCH_RESULT();
  }
  switch (dtype) {
    case MXQ_F32: return launch_fq<float>(p, ref, smem, grid, st);
    case MXQ_F16: return launch_fq<__half>(p, ref, smem, grid, st);
    default: return launch_fq<__nv_bfloat16>(p, ref, smem, grid, st);
  }
}

// Several weights of one row length in one launch, and/or a mask that moves the pooled group
// (importance-driven allocation): both run the row-resident kernel.
extern "C" int mxq_fakequant_fwd_multi(const void* const* x, void* const* out, const int64_t* rows, int n,
                                       int64_t cols, int dtype, int group, int low_bits,
                                       const uint8_t* pooled_mask, void* stream) {
  if (n < 0 || cols < 0) return MXQ_E_SHAPE;
  if (n == 0 || cols == 0) return MXQ_OK;
  if (!x || !out || !rows) return MXQ_E_NULL;
  if (dtype != MXQ_F32 && dtype != MXQ_F16 && dtype != MXQ_BF16) return MXQ_E_DTYPE;
  const int esize = dtype == MXQ_F32 ? 4 : 2;
  if ((group != 16 && group != 128) || cols % (4 * group) || cols > (1 << 24)) return MXQ_E_SHAPE;
  if (low_bits < 1 || low_bits > 8 || (esize == 2 && low_bits != 2)) return MXQ_E_UNSUPPORTED;
  const int cpr = (int)(cols * esize / 16);
  if (cpr > 6144) return MXQ_E_UNSUPPORTED;      // longer rows: mxq_fakequant_fwd (shared-memory ring)
  cudaStream_t st = as_stream(stream);
  for (int i0 = 0; i0 < n; i0 += kFQMaxTensors) {
    FQRowTable tab{};
    tab.pooled_mask = pooled_mask;
    int64_t total = 0;
    for (int i = i0; i < n && i < i0 + kFQMaxTensors; ++i) {
      if (rows[i] < 0 || rows[i] > INT32_MAX) return MXQ_E_SHAPE;
      if (rows[i] == 0) continue;
      MXQ_CHECK_PTR(x[i]);
      MXQ_CHECK_PTR(out[i]);
      total += rows[i];
      if (total > INT32_MAX) return MXQ_E_SHAPE;
      tab.x[tab.n] = (const uint4*)x[i]; tab.out[tab.n] = (uint4*)out[i]; tab.row_end[tab.n] = (int)total;
      ++tab.n;
    }
    if (tab.n == 0) continue;
    if (!launch_fq_row_any(tab, dtype, group, cpr, low_bits, st)) return MXQ_E_UNSUPPORTED;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return MXQ_OK;
}

extern "C" int mxq_ste_bwd(const void* grad_out, const void* x, void* grad_in, int64_t n,
                           int dtype, float lo, float hi, void* stream) {
  if (n < 0) return MXQ_E_SHAPE;
  if (n == 0) return MXQ_OK;
  MXQ_CHECK_PTR(grad_out);
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(grad_in);
  cudaStream_t st = as_stream(stream);
  switch (dtype) {
    case MXQ_F32: return launch_ste<float>(grad_out, x, grad_in, n, lo, hi, st);
    case MXQ_F16: return launch_ste<__half>(grad_out, x, grad_in, n, lo, hi, st);
    case MXQ_BF16: return launch_ste<__nv_bfloat16>(grad_out, x, grad_in, n, lo, hi, st);
    default: return MXQ_E_DTYPE;
  }
}
