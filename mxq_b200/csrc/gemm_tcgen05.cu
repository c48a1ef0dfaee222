// Prefill dequant-GEMM for the packed mixed 2/4-bit layout on tcgen05 / TMEM / TMA (sm_100a).
//
//   y[m, oc] = sum_k x[m, k] * dequant(W)[oc, k]       x fp16 [M, IC], y fp16 [M, OC]
//
// The reference has no prefill kernel for the mixed layout (its AWQ mma.sync GEMM,
// gemm_cuda_gen.cu:424-478, is neither built nor exported); decode formula as in
// gemv_mxq_cuda.cu:131-136,152-153,178-179,191-192.
//
// CTA tile: 256 tokens x 256 weight rows, K block 64 (= one 64-column block of the packed
// layout).  TMEM: two 128-lane x 256-column fp32 accumulators (all 512 columns).
// Warp roles (320 threads):
//   warp 0      : TMA producer -- activations [256 x 64] fp16 per stage, 128B-swizzled, mbarrier tx
//   warp 1      : TMEM alloc + single-thread tcgen05.mma issue (2 accumulators x 4 UMMA_K per
//                 K block), tcgen05.commit frees the stage / publishes the accumulators
//   warps 2..9  : dequant producers -- thread = weight row; per K block: one 128-bit load of 48
//                 two-bit + 8 four-bit codes, one 32-bit load of 8 four-bit codes, metadata;
//                 LOP3 -> {1024+q} fp16 pairs, HSUB2 (1024+z), HMUL2 scale, eight 128-bit
//                 st.shared into the canonical K-major SWIZZLE_128B UMMA layout; next K block's
//                 packed words are prefetched into registers.  Afterwards the same warps run
//                 the epilogue: tcgen05.ld 32x32b.x32 -> fp16 -> 128-bit global stores.
// The dequantized operand never touches HBM: weights cost 0.3756 B each per M tile.
//
// M > 256 (and M <= 256 when a workspace allows a K split) runs the CTA-pair variant further down
// (cta_group::2, one M = 256 MMA per K step across two SMs).  Its tile schedule (pair::Plan) cuts the
// tiles of a partly filled last wave along K and adds the fp32 partials in a second pass
// (gemm_split_reduce_kernel); for the column-sharded multi-GPU case the epilogue -- or that second
// pass -- stores whole 512-byte row segments into every peer's buffer or once into the buffer's
// NVSwitch multicast mapping (mxq_gemm_scatter / mxq_gemm_multicast).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace mxq {
namespace gemm {

constexpr int BM = 256;       // tokens per CTA (two UMMA M=128 halves)
constexpr int BN = 256;       // weight rows per CTA (UMMA N)
constexpr int BK = 64;        // K per stage = one packed block
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 32 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int NUM_DEQ_WARPS = 8;
constexpr int THREADS = 32 * (2 + NUM_DEQ_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int STG_PITCH = 80;                 // bytes per row of the per-warp transpose buffer (4 x 16 B + pad)
constexpr int STG_BYTES = NUM_DEQ_WARPS * 32 * STG_PITCH;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/ + STG_BYTES;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle, 8-row atoms 1024 B apart (SBO); LBO unused for swizzled K-major.
// Field layout: cute/arch/mma_sm100_desc.hpp UMMA::SmemDescriptor (version = 1, layout_type 2).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= (uint64_t)0 << 16;                               // leading byte offset [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                               // version [46,48) = 1 on sm_100
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// UMMA::InstrDescriptor: c=F32 (bit 4), a=b=F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(orv));
  return d;
}
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t h2_bcast(float v) {
  const __half2 h = __float2half2_rn(v);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// 16 two-bit codes of one word -> 16 fp16 weights in column order (two 16-byte chunks).
// (w >> 2j) & 0x00030003 yields codes j (low half) and j+8 (high half); PRMT re-pairs
// consecutive columns.
__device__ __forceinline__ void dequant_2b(uint32_t w, uint32_t zmagic, uint32_t scale2, uint4& c0, uint4& c1) {
  uint32_t h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    h[j] = hmul2(hsub2(lop3_and_or(w >> (2 * j), 0x00030003u, 0x64006400u), zmagic), scale2);
  // h[j] = {col j, col j+8}.  cols (2i, 2i+1) = low halves of h[2i], h[2i+1]; cols (8+2i, 9+2i) = high halves
  c0 = make_uint4(prmt(h[0], h[1], 0x5410), prmt(h[2], h[3], 0x5410), prmt(h[4], h[5], 0x5410), prmt(h[6], h[7], 0x5410));
  c1 = make_uint4(prmt(h[0], h[1], 0x7632), prmt(h[2], h[3], 0x7632), prmt(h[4], h[5], 0x7632), prmt(h[6], h[7], 0x7632));
}
// 8 four-bit codes of one word -> 8 fp16 weights (one 16-byte chunk); (w >> 4j) & 0x000F000F = nibbles j, j+4
__device__ __forceinline__ uint4 dequant_4b(uint32_t w, uint32_t zmagic, uint32_t scale2) {
  uint32_t h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    h[j] = hmul2(hsub2(lop3_and_or(w >> (4 * j), 0x000F000Fu, 0x64006400u), zmagic), scale2);
  // h[j] = {col j, col j+4}: cols (0,1)=lo(h0,h1) (2,3)=lo(h2,h3) (4,5)=hi(h0,h1) (6,7)=hi(h2,h3)
  return make_uint4(prmt(h[0], h[1], 0x5410), prmt(h[2], h[3], 0x5410), prmt(h[0], h[1], 0x7632), prmt(h[2], h[3], 0x7632));
}

// 16 bytes to an NVSwitch multicast address: the switch replicates the store into the buffer of
// every GPU bound to the multicast object (the type only names the vector shape)
__device__ __forceinline__ void multimem_st_16(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int MAX_PEERS = 8;
struct Params {
  mxq_packed_t w;
  const __half* wdense;      // dense-B debug path only
  __half* y[MAX_PEERS];      // output base pointers: [0] local; > 1 entries = peers' buffers mapped
  int npeers;                //   over NVLink (fused column all-gather: every tile is stored to all)
  int mc;                    // y[0] is an NVLink multicast address: one multimem.st reaches every rank's buffer
  int ldy, col0;             // output row stride (elements) and first output column of this shard
  int M, IC, OC;
  int dbg;                   // profiling only (MXQ_GEMM_DBG): 1 = one M half, 2 = no TMA after the first ring fill
  unsigned long long* dbg_host;   // debugging only (MXQ_GEMM_DBG_PTR): pinned host memory for watchdog records
  // CTA-pair kernel tile schedule (pair::Plan): clusters [0, full) compute one whole 512 x 256 tile;
  // the remaining tiles -- the ones that would form a partly filled last wave -- are cut into
  // `split` K slices, one cluster each, reduced through fp32 partials in the workspace
  int mt, full, split;
  float4* partial;           // [tail tiles * split][2 ranks][8 * 8 * 256] float4, summed by gemm_split_reduce_kernel
};

template <bool kDenseB>
__global__ void __launch_bounds__(THREADS, 1) gemm_mxq_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                              const __grid_constant__ CUtensorMap tmap_w,
                                                              const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_a = bars;                 // [STAGES]
  uint64_t* full_b = bars + STAGES;        // [STAGES]
  uint64_t* empty = bars + 2 * STAGES;     // [STAGES]
  uint64_t* tmem_full = bars + 3 * STAGES; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
  uint8_t* stage_base = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int num_kb = p.IC / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    if (kDenseB) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&full_b[s], kDenseB ? 1 : NUM_DEQ_WARPS);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (activations; dense-B debug path also loads the weights) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if ((p.dbg & 2) && kb >= STAGES) {        // profiling: operands stay stale, no smem writes
          mbar_arrive(&full_a[s]);
          if (kDenseB) mbar_arrive(&full_b[s]);
          continue;
        }
        mbar_arrive_expect_tx(&full_a[s], A_STAGE_BYTES);
        tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmap_x, kb * BK, m0, &full_a[s]);
        if (kDenseB) {
          mbar_arrive_expect_tx(&full_b[s], B_STAGE_BYTES);
          tma_load_2d(smem_b + s * B_STAGE_BYTES, &tmap_w, kb * BK, n0, &full_b[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, BN);
      const uint32_t desc_hi = (uint32_t)(make_smem_desc(0) >> 32);
      const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFF) >> 4, b_lo0 = (smem_u32(smem_b) & 0x3FFFF) >> 4;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_a[s], ph);
        mbar_wait(&full_b[s], ph);
        tc_fence_after();
        // descriptors = {address >> 4 + immediate, constant high word}: the issuing thread shares
        // its scheduler with dequantizer warps, its instruction count is on the critical path
        const uint32_t a_lo = a_lo0 + s * (A_STAGE_BYTES >> 4), b_lo = b_lo0 + s * (B_STAGE_BYTES >> 4);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if ((p.dbg & 1) && h == 1) break;       // profiling: half the MMAs
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_lo + ((h * (128 * BK * 2) + k * (UMMA_K * 2)) >> 4));
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (b_lo + ((k * (UMMA_K * 2)) >> 4));
            umma_f16(tmem_base + h * BN, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);        // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);          // accumulators complete
    }
  } else {
    // ===== dequant producers (thread = weight row), then epilogue =====
    const int dw = warp - 2;                      // 0..7
    const int row_local = dw * 32 + lane;         // 0..255
    const int oc = n0 + row_local;
    if (!kDenseB) {
      const bool row_ok = oc < p.OC;
      const int ocs = row_ok ? oc : 0;
      const int nblk = num_kb;
      const int nchunk = (nblk + 63) >> 6;
      const uint4* wrow = reinterpret_cast<const uint4*>(p.w.weight + (size_t)ocs * nblk * 4);
      const int32_t* wlrow = p.w.weight_last + (size_t)ocs * nblk;
      const uint16_t* zsrow = reinterpret_cast<const uint16_t*>(p.w.zeros_and_scales) + (size_t)ocs * 64 * nchunk;
      const uint8_t* z2row = reinterpret_cast<const uint8_t*>(p.w.zeros_2nd) + (size_t)(ocs >> 2) * 128 * nchunk;
      const __half* s2row = reinterpret_cast<const __half*>(p.w.scales_2nd) + (size_t)(ocs >> 2) * nblk * 3;
      const float s4 = __half2float(reinterpret_cast<const __half*>(p.w.scales_4b)[ocs]);
      const uint32_t z4 = ((uint32_t)p.w.zeros_4b[ocs >> 3] >> (4 * (ocs & 7))) & 0xF;
      const uint32_t z4magic = (0x6400u | z4) * 0x00010001u;
      const uint32_t s4h2 = h2_bcast(row_ok ? s4 : 0.f);

      // dequantize one K block of this thread's row and publish it to the MMA warp
      auto produce = [&](int kb, const uint4& cwq, uint32_t cwl, uint32_t czs, uint32_t cz2,
                         float s20, float s21, float s22) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        uint4 ch[8];
        const uint32_t ws[3] = {cwq.x, cwq.y, cwq.z};
        const float cs2[3] = {s20, s21, s22};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint32_t z1 = (czs >> (2 * k)) & 3;
          const float c = (float)((czs >> (8 + 2 * k)) & 3);
          const float zz = (float)((cz2 >> (2 * k)) & 3);
          const float scale = row_ok ? cs2[k] * (c - zz) : 0.f;
          dequant_2b(ws[k], (0x6400u | z1) * 0x00010001u, h2_bcast(scale), ch[2 * k], ch[2 * k + 1]);
        }
        ch[6] = dequant_4b(cwq.w, z4magic, s4h2);
        ch[7] = dequant_4b(cwl, z4magic, s4h2);
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* brow = smem_b + s * B_STAGE_BYTES + row_local * 128;
        const int sw = row_local & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(brow + ((c ^ sw) << 4)) = ch[c];
        fence_proxy_async();           // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_b[s]);
      };

      if ((nblk & 3) == 0) {
        // ---- cooperative loads: 4 K blocks per group.  The 128-bit weight words are fetched with
        // 4 lanes per row (64 contiguous bytes per row, 8 rows per request instead of 32 scattered
        // rows) and transposed through a warp-private staging buffer; the per-row tail words and
        // metadata are fetched as one 16-byte vector per lane per group.  Next group's loads are
        // in flight while the current group is dequantized.
        uint8_t* stg = stage_base + dw * (32 * STG_PITCH);
        const int ngroups = nblk >> 2;
        const int q = lane & 3, r8 = lane >> 2;
        const uint4* wbase[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int orow = n0 + dw * 32 + 8 * i + r8;
          orow = orow < p.OC ? orow : 0;
          wbase[i] = reinterpret_cast<const uint4*>(p.w.weight + (size_t)orow * nblk * 4) + q;
        }
        const uint4* wl4 = reinterpret_cast<const uint4*>(wlrow);
        const uint32_t* zsw = reinterpret_cast<const uint32_t*>(p.w.zeros_and_scales) + (size_t)ocs * 32 * nchunk;
        const uint32_t* z2w = reinterpret_cast<const uint32_t*>(p.w.zeros_2nd) + (size_t)(ocs >> 2) * 32 * nchunk;
        const uint2* s2v2 = reinterpret_cast<const uint2*>(s2row);
        uint4 pw[4], pwl, pzs, pz2;
        uint2 ps2[3];
        auto fetch = [&](int g) {
#pragma unroll
          for (int i = 0; i < 4; ++i) pw[i] = __ldg(wbase[i] + 4 * g);
          pwl = __ldg(wl4 + g);
          const int kb0 = 4 * g;
          const int mword = (kb0 >> 6) * 32 + (kb0 & 31);
          pzs = __ldg(reinterpret_cast<const uint4*>(zsw + mword));
          pz2 = __ldg(reinterpret_cast<const uint4*>(z2w + mword));
#pragma unroll
          for (int i = 0; i < 3; ++i) ps2[i] = __ldg(s2v2 + 3 * g + i);
        };
        fetch(0);
        for (int g = 0; g < ngroups; ++g) {
          // transpose the weight words: lane (r8, q) holds rows 8i + r8, K block q
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(stg + (8 * i + r8) * STG_PITCH + q * 16) = pw[i];
          const uint4 cwl = pwl, czs = pzs, cz2 = pz2;
          const uint2 cs2[3] = {ps2[0], ps2[1], ps2[2]};
          __syncwarp();
          uint4 wq[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) wq[kk] = *reinterpret_cast<const uint4*>(stg + lane * STG_PITCH + kk * 16);
          __syncwarp();
          if (g + 1 < ngroups) fetch(g + 1);
          const int hsh = (((4 * g) & 63) >> 5) * 16;   // metadata half-word of these 4 K blocks
          const uint32_t wlw[4] = {cwl.x, cwl.y, cwl.z, cwl.w};
          const uint32_t zsv[4] = {czs.x, czs.y, czs.z, czs.w};
          const uint32_t z2v[4] = {cz2.x, cz2.y, cz2.z, cz2.w};
          const uint32_t s2w[6] = {cs2[0].x, cs2[0].y, cs2[1].x, cs2[1].y, cs2[2].x, cs2[2].y};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // scales_2nd halves 3*kk .. 3*kk+2 of the group's 12
            float sv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const int hidx = 3 * kk + k;
              const uint32_t wsel = s2w[hidx >> 1];
              sv[k] = __half2float(__ushort_as_half((unsigned short)((hidx & 1) ? (wsel >> 16) : (wsel & 0xFFFF))));
            }
            produce(4 * g + kk, wq[kk], wlw[kk], (zsv[kk] >> hsh) & 0xFFFF, (z2v[kk] >> (hsh >> 1)) & 0xFF,
                    sv[0], sv[1], sv[2]);
          }
        }
      } else {
        // ---- generic K (IC % 256 != 0): per-block scalar loads, one block of register prefetch
        auto meta_idx = [&](int kb, int& hw_idx, int& b_idx) {
          const int chunk = kb >> 6, bp = kb & 63;
          const int word = chunk * 32 + (bp & 31), ph = bp >> 5;
          hw_idx = word * 2 + ph;
          b_idx = word * 4 + ph;
        };
        uint4 wq = __ldg(wrow);
        uint32_t wl = (uint32_t)__ldg(wlrow);
        int hi, bi;
        meta_idx(0, hi, bi);
        uint32_t zs = __ldg(zsrow + hi);
        uint32_t z2 = __ldg(z2row + bi);
        float s2v[3] = {__half2float(__ldg(s2row)), __half2float(__ldg(s2row + 1)), __half2float(__ldg(s2row + 2))};
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint4 cwq = wq;
          const uint32_t cwl = wl, czs = zs, cz2 = z2;
          const float c0 = s2v[0], c1 = s2v[1], c2 = s2v[2];
          if (kb + 1 < num_kb) {
            wq = __ldg(wrow + kb + 1);
            wl = (uint32_t)__ldg(wlrow + kb + 1);
            meta_idx(kb + 1, hi, bi);
            zs = __ldg(zsrow + hi);
            z2 = __ldg(z2row + bi);
            const __half* sp = s2row + (size_t)(kb + 1) * 3;
            s2v[0] = __half2float(__ldg(sp)); s2v[1] = __half2float(__ldg(sp + 1)); s2v[2] = __half2float(__ldg(sp + 2));
          }
          produce(kb, cwq, cwl, czs, cz2, c0, c1, c2);
        }
      }
    }
    // ===== epilogue: TMEM -> registers -> fp16 -> global =====
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int half = dw >> 2;                 // accumulator (token half)
    const int quad = warp & 3;                // TMEM lane quarter this warp may access
    const int token = m0 + half * 128 + quad * 32 + lane;
    const size_t yoff = (size_t)token * p.ldy + p.col0 + n0;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN + cb * 32, v);
      tmem_ld_wait();
      if (token < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = n0 + cb * 32 + q * 8;
          if (col + 8 <= p.OC) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              oh[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
            // local store, or one store per peer buffer (st.global on NVLink-mapped addresses)
            for (int pe = 0; pe < p.npeers; ++pe)
              *reinterpret_cast<uint4*>(p.y[pe] + yoff + cb * 32 + q * 8) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// =============================================================================================
// CTA-pair variant (cta_group::2).  Measured on B200 (profiles/probes/probe_gemm_smem.py): a
// single-CTA tcgen05.mma with both operands in shared memory fetches them at 64 B/cycle, i.e. a
// 128 x 256 x 16 MMA takes (4 KB + 8 KB) / 64 = 192 cycles instead of its 128-cycle floor -- the
// single-CTA kernel above tops out near 70 % of the tensor peak however fast the dequantizers are.
// A CTA pair issues ONE M = 256 MMA per K step: each SM fetches its own 128 x 16 A slice and HALF
// of B (128 weight rows x 16), 8 KB per 128 cycles = the floor, and each CTA dequantizes only 128
// weight rows per K block.
//
//   cluster (2,1,1) along M: rank r owns tokens [m0 + 256 r, +256) (two M = 128 halves, 512 TMEM
//   columns) and dequantizes weight rows [n0 + 128 r, +128) of the pair's 256;
//   warp 0      : TMA producer of its CTA's A tile; transaction bytes of BOTH CTAs land on the
//                 leader's full_a barrier (cp.async.bulk.tensor .cta_group::2)
//   warp 1      : TMEM alloc (cta_group::2, both CTAs); the leader's lane 0 issues every MMA and
//                 multicasts tcgen05.commit to both CTAs' empty / tmem_full barriers
//   warps 2..9  : dequantizers, thread = weight row, two sets of 4 warps taking alternate groups
//                 of 4 K blocks (each set has two groups' worth of MMA time per group); they arrive
//                 on their own CTA's full_b barrier, rank 1's idle warp-1 thread relays that to the
//                 leader's full_b_peer with a cluster-scope release; afterwards the epilogue.
// =============================================================================================
namespace pair {

constexpr int BM = 256;        // tokens per CTA
constexpr int BN = 256;        // weight rows per CTA pair (UMMA N)
constexpr int BNH = 128;       // weight rows dequantized per CTA
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;    // 32 KB
constexpr int B_STAGE_BYTES = BNH * BK * 2;   // 16 KB
constexpr int SET_WARPS = NUM_DEQ_WARPS / 2;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/ + STG_BYTES;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of rank 1 -> same offset in rank 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as low words + the shared constant high word
__device__ __forceinline__ void umma2_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of the pair have completed) on the barrier at this offset in both CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA load whose completion bytes are reported to the barrier at shared::cluster address `bar_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                 uint32_t bar_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
// Arrive on the barrier at shared::cluster address `bar_addr` (the leader's) with the default
// .release.cta semantics.  The stores being published are st.shared to the issuing CTA's own shared
// memory, already pushed to the async proxy by fence.proxy.async.shared::cta: CTA-scope
// performed-ness of a shared-memory store IS its presence in that SM's shared memory, which is
// what the pair's tensor cores read.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// remote arrive with cluster-scope release (compiles to MEMBAR.ALL.GPU + arrive: issue it from a
// thread without loads in flight)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Hang guard: a barrier that never completes is a programming error; after kHangCycles (~20 s at 2 GHz --
// far beyond any time-slice, preemption or debugger pause a healthy launch can see; 0.5 s in MXQ_DEBUG
// builds) the kernel traps instead of wedging the GPU.  (A guard-free `while (!try_wait) {}` was measured
// too: the tighter polling of the 8 producer warps slows the MMA-issuing thread that shares their
// schedulers -- 1262 -> 1049 TFLOP/s on 4096^2 -- so the counted loop stays.)
#ifdef MXQ_DEBUG
constexpr long long kHangCycles = 1000000000LL;
#else
constexpr long long kHangCycles = 40000000000LL;
#endif
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && clock64() - t0 > kHangCycles) __trap();
  }
}
// debugging: waits that leave a record {site, kb, block, rank} in pinned host memory before trapping
__device__ __noinline__ void wd_record(unsigned long long* h, int site, int kb) {
  if (h) {
    const unsigned long long v = ((unsigned long long)site << 48) | ((unsigned long long)(kb & 0xFFFF) << 32) |
                                 ((unsigned long long)(blockIdx.x & 0xFFFF) << 16) |
                                 (threadIdx.x & 0xFFFF);
    const unsigned slot = atomicAdd((unsigned*)h, 1u);
    if (slot < 60) h[1 + slot] = v;
    __threadfence_system();
  }
}
// The record in pinned host memory (`h`, MXQ_GEMM_DBG_PTR) exists in MXQ_DEBUG builds only.
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, unsigned long long* h, int site, int kb,
                                              bool cluster) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!(cluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if ((++spins & 0x3FFu) == 0 && clock64() - t0 > kHangCycles) {
      wd_record(h, site, kb);
      // let the other stuck threads record too before the context dies
      const long long t1 = clock64();
      while (clock64() - t1 < 200000000LL) {}
      __trap();
    }
  }
}

// Second pass of the K-split tail tiles (Plan): the clusters that ran a K slice left their fp32
// accumulators in the exchange buffer, thread-major -- 16-byte chunk c (columns 4 c .. 4 c + 3 of
// the tile) of epilogue thread t at [c * 256 + t] of the CTA's slot -- so that the writes of the
// GEMM epilogue and the reads here are coalesced.  A block takes the 32 tokens of one epilogue
// warp: warp w adds the slices of columns 32 w .. 32 w + 31 in slice order (deterministic), the
// fp16 rows are transposed through shared memory and leave as whole 512-byte row segments -- to
// the local buffer, to every peer, or once to the multicast address (fused column all-gather).
constexpr int SLOT_F4 = (BN / 4) * 256;                // float4 per CTA partial (256 KB)
struct OutPtrs {
  __half* y[MAX_PEERS];
  int npeers, mc;
};
__global__ void __launch_bounds__(256) gemm_split_reduce_kernel(const float4* __restrict__ partial, const OutPtrs out,
                                                                int split, int full, int mt, int M, int OC, int ldy,
                                                                int col0) {
  constexpr int PITCH = BN * 2 + 16;
  __shared__ __align__(16) uint8_t rows[32 * PITCH];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dwi = blockIdx.x & 7, rank = (blockIdx.x >> 3) & 1, tail = blockIdx.x >> 4;
  const int tile = full + tail;
  const int m0 = (2 * (tile % mt) + rank) * BM, n0 = (tile / mt) * BN;
  // epilogue thread t = 32 * dw + lane of the GEMM: token half dw >> 2, TMEM lane quarter (dw + 2) & 3
  const int tok0 = m0 + (dwi >> 2) * 128 + ((dwi + 2) & 3) * 32;
  const float4* src = partial + ((size_t)(tail * split) * 2 + rank) * SLOT_F4 + (8 * w) * 256 + dwi * 32 + lane;
  float4 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldcs(src + j * 256);
  for (int sl = 1; sl < split; ++sl) {
    src += 2 * SLOT_F4;
    float4 f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __ldcs(src + j * 256);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j].x += f[j].x; acc[j].y += f[j].y; acc[j].z += f[j].z; acc[j].w += f[j].w;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
    oh[0] = __floats2half2_rn(acc[2 * q].x, acc[2 * q].y);
    oh[1] = __floats2half2_rn(acc[2 * q].z, acc[2 * q].w);
    oh[2] = __floats2half2_rn(acc[2 * q + 1].x, acc[2 * q + 1].y);
    oh[3] = __floats2half2_rn(acc[2 * q + 1].z, acc[2 * q + 1].w);
    *reinterpret_cast<uint4*>(rows + lane * PITCH + w * 64 + q * 16) = o;
  }
  __syncthreads();
  const int col = n0 + lane * 8;
  if (col + 8 > OC) return;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = 4 * w + rr;
    if (tok0 + r >= M) break;
    const uint4 o = *reinterpret_cast<const uint4*>(rows + r * PITCH + lane * 16);
    const size_t off = (size_t)(tok0 + r) * ldy + col0 + col;
    if (out.mc) {
      multimem_st_16(out.y[0] + off, o);
    } else {
      for (int pe = 0; pe < out.npeers; ++pe) *reinterpret_cast<uint4*>(out.y[pe] + off) = o;
    }
  }
}

template <bool kDenseB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
    gemm_mxq_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_a = bars;                 // [STAGES]  used in the leader only
  uint64_t* full_b = bars + STAGES;        // [STAGES]  per CTA: this CTA's half of B is in place
  // empty[set][stage]: the K blocks that pass through a stage alternate between the two dequantizer
  // sets, and a set only ever waits for the OTHER set's K block to be consumed.  One barrier per
  // (set, stage) lets every waiter observe every phase of the barrier it waits on -- with a single
  // barrier per stage a set would skip phases and the parity test could not tell "not yet" from
  // "two phases ago".
  uint64_t* empty = bars + 2 * STAGES;     // [2][STAGES]  per CTA
  uint64_t* full_b_peer = bars + 4 * STAGES;   // [STAGES]  leader only: rank 1's half is in place
  uint64_t* tmem_full = bars + 5 * STAGES; // [1]       per CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 * STAGES + 1);
  uint8_t* stage_base = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // tile schedule: 1-D grid of clusters, M tile fastest (consecutive clusters share weight rows)
  const int cid = blockIdx.x >> 1;
  int tile = cid, slice = 0, kb_begin = 0, num_kb = p.IC / BK;
  const bool partial = cid >= p.full;
  if (partial) {
    const int u = cid - p.full;
    const int ng = num_kb >> 2;                    // split only when IC % 256 == 0
    tile = p.full + u / p.split;
    slice = u % p.split;
    kb_begin = 4 * ((ng * slice) / p.split);
    num_kb = 4 * ((ng * (slice + 1)) / p.split) - kb_begin;
  }
  const int m0 = (2 * (tile % p.mt) + (int)rank) * BM, n0 = (tile / p.mt) * BN;
  const int nrow0 = n0 + (int)rank * BNH;          // first weight row this CTA dequantizes
  // K blocks are numbered locally (kb = 0 is global block kb_begin).  Local block kb is dequantized
  // by set (kb >> 2) & 1 and is the (kb >> 3)-th block of that set in its
  // stage; its consumption completes phase (kb >> 3) of empty[set][kb % STAGES]
  const bool direct = !(p.dbg & 16);   // MXQ_GEMM_DBG=16: publish rank 1's half through the relay thread
  auto empty_of = [&](int kb) { return &empty[((kb >> 2) & 1) * STAGES + (kb % STAGES)]; };
  auto wait_stage_free = [&](int kb, int site) {    // before K block kb may overwrite its stage
    if (kb >= STAGES) mbar_wait_dbg(empty_of(kb - STAGES), ((kb - STAGES) >> 3) & 1, p.dbg_host, site, kb, false);
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    if (kDenseB) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_a[s], 1);                                  // the leader's arrive.expect_tx
      // one warp set of this CTA (relay mode) or of both CTAs (direct mode: rank 1 arrives remotely)
      mbar_init(&full_b[s], kDenseB ? 1 : (direct ? 2 * SET_WARPS : SET_WARPS));
      mbar_init(&full_b_peer[s], 1);                             // rank 1's relay thread
      mbar_init(&empty[s], 1);
      mbar_init(&empty[STAGES + s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();            // barriers of both CTAs are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: this CTA's 256 tokens; completion reported to the leader =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        wait_stage_free(kb, 1);
        if (rank == 0) {
          mbar_arrive_expect_tx(&full_a[s], 2 * A_STAGE_BYTES);
          if (kDenseB) mbar_arrive_expect_tx(&full_b[s], 2 * B_STAGE_BYTES);
        }
        tma_load_2d_pair(smem_a + s * A_STAGE_BYTES, &tmap_x, (kb_begin + kb) * BK, m0, smem_u32(&full_a[s]) & kPeerBitMask);
        if (kDenseB)
          tma_load_2d_pair(smem_b + s * B_STAGE_BYTES, &tmap_w, (kb_begin + kb) * BK, nrow0, smem_u32(&full_b[s]) & kPeerBitMask);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(256, BN);
      // The issuing thread shares its scheduler with two dequantizer warps, so its own instruction
      // count is on the critical path: descriptors are {low word = address >> 4, constant high
      // word}; per MMA only the low words get an immediate added.
      const uint32_t desc_hi = (uint32_t)(make_smem_desc(0) >> 32);
      const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFF) >> 4, b_lo0 = (smem_u32(smem_b) & 0x3FFFF) >> 4;
      const bool prof = p.dbg_host != nullptr;
      long long wait_a = 0, wait_b = 0;          // debugging: cycles the issuer spent on each barrier
      const long long c_begin = prof ? clock64() : 0;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (prof) {
          const long long c0 = clock64();
          mbar_wait_dbg(&full_a[s], ph, p.dbg_host, 2, kb, true);
          const long long c1 = clock64();
          mbar_wait_dbg(&full_b[s], ph, p.dbg_host, 3, kb, direct);
          wait_a += c1 - c0; wait_b += clock64() - c1;
        } else {
          mbar_wait_cluster(&full_a[s], ph);
          if (direct) mbar_wait_cluster(&full_b[s], ph); else mbar_wait(&full_b[s], ph);
        }
        if (!kDenseB && !direct) mbar_wait_dbg(&full_b_peer[s], ph, p.dbg_host, 4, kb, true);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (A_STAGE_BYTES >> 4), b_lo = b_lo0 + s * (B_STAGE_BYTES >> 4);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma2_f16_lohi(tmem_base + h * BN, a_lo + ((h * (128 * BK * 2) + k * (UMMA_K * 2)) >> 4),
                           b_lo + ((k * (UMMA_K * 2)) >> 4), desc_hi, idesc, k == 0 ? accumulate : 1u);
          }
        }
        accumulate = 1;
        umma2_commit_both(empty_of(kb));   // frees the stage in both CTAs
      }
      umma2_commit_both(tmem_full);        // accumulators of both CTAs complete
      if (p.dbg_host && blockIdx.x == 0) {
        volatile unsigned long long* hv = p.dbg_host;
        hv[61] = (unsigned long long)wait_a;
        hv[62] = (unsigned long long)wait_b;
        hv[63] = (unsigned long long)(clock64() - c_begin);   // meaningful with MXQ_GEMM_DBG_PTR set
        __threadfence_system();
      }
    } else if (!kDenseB && !direct && rank == 1 && lane == 0) {
      // Relay: the dequantizers publish their half of B on this CTA's own barrier (cheap .cta
      // release); this otherwise idle thread forwards it to the leader with a cluster-scope
      // release -- a MEMBAR.ALL.GPU that would stall a dequantizer on its prefetch loads.
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait_dbg(&full_b[s], (kb / STAGES) & 1, p.dbg_host, 5, kb, false);
        mbar_arrive_cluster(smem_u32(&full_b_peer[s]) & kPeerBitMask);
      }
    }
  } else {
    // ===== dequant producers (thread = weight row, two alternating warp sets), then epilogue =====
    const int dw = warp - 2;                      // 0..7
    const int set = dw >> 2;                      // groups g with (g & 1) == set
    const int row_local = (dw & 3) * 32 + lane;   // 0..127
    const int oc = nrow0 + row_local;
    if (!kDenseB) {
      const bool row_ok = oc < p.OC;
      const int ocs = row_ok ? oc : 0;
      const int nblk = p.IC / BK;                   // blocks per weight row (num_kb of them are this cluster's)
      const int nchunk = (nblk + 63) >> 6;
      const uint4* wrow = reinterpret_cast<const uint4*>(p.w.weight + (size_t)ocs * nblk * 4);
      const int32_t* wlrow = p.w.weight_last + (size_t)ocs * nblk;
      const uint16_t* zsrow = reinterpret_cast<const uint16_t*>(p.w.zeros_and_scales) + (size_t)ocs * 64 * nchunk;
      const uint8_t* z2row = reinterpret_cast<const uint8_t*>(p.w.zeros_2nd) + (size_t)(ocs >> 2) * 128 * nchunk;
      const __half* s2row = reinterpret_cast<const __half*>(p.w.scales_2nd) + (size_t)(ocs >> 2) * nblk * 3;
      const float s4 = __half2float(reinterpret_cast<const __half*>(p.w.scales_4b)[ocs]);
      const uint32_t z4 = ((uint32_t)p.w.zeros_4b[ocs >> 3] >> (4 * (ocs & 7))) & 0xF;
      const uint32_t z4magic = (0x6400u | z4) * 0x00010001u;
      const uint32_t s4h2 = h2_bcast(row_ok ? s4 : 0.f);

      auto produce = [&](int kb, const uint4& cwq, uint32_t cwl, uint32_t czs, uint32_t cz2,
                         float s20, float s21, float s22) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        uint4 ch[8];
        const uint32_t ws[3] = {cwq.x, cwq.y, cwq.z};
        const float cs2[3] = {s20, s21, s22};
        if (p.dbg & 32) {          // profiling: no dequant arithmetic (garbage operand)
#pragma unroll
          for (int c = 0; c < 8; ++c) ch[c] = make_uint4(cwq.x + c, cwq.y, czs, cwl);
        } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint32_t z1 = (czs >> (2 * k)) & 3;
          const float c = (float)((czs >> (8 + 2 * k)) & 3);
          const float zz = (float)((cz2 >> (2 * k)) & 3);
          const float scale = row_ok ? cs2[k] * (c - zz) : 0.f;
          dequant_2b(ws[k], (0x6400u | z1) * 0x00010001u, h2_bcast(scale), ch[2 * k], ch[2 * k + 1]);
        }
        ch[6] = dequant_4b(cwq.w, z4magic, s4h2);
        ch[7] = dequant_4b(cwl, z4magic, s4h2);
        }
        if (p.dbg & 512) {         // profiling: every lane polls the barrier
          wait_stage_free(kb, 6);
        } else {                   // one lane polls, __syncwarp orders the others' stores after its acquire
          if (lane == 0) wait_stage_free(kb, 6);
          __syncwarp();
        }
        uint8_t* brow = smem_b + s * B_STAGE_BYTES + row_local * 128;
        const int sw = row_local & 7;
        if (!(p.dbg & 256)) {     // dbg 256: profiling, no operand stores
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(brow + ((c ^ sw) << 4)) = ch[c];
        }
        // generic stores to this CTA's shared memory -> async proxy (the pair's tensor cores).  The
        // unqualified fence.proxy.async compiles to MEMBAR.ALL.GPU and stalls on the prefetch loads.
        if (!(p.dbg & 128)) fence_proxy_async();   // dbg 128: profiling, no proxy fence (wrong results)
        __syncwarp();
        if (lane == 0) {
          if (direct) mbar_arrive_remote(smem_u32(&full_b[s]) & kPeerBitMask);
          else mbar_arrive(&full_b[s]);
        }
      };

      if ((nblk & 3) == 0) {
        uint8_t* stg = stage_base + dw * (32 * STG_PITCH);
        const int ngroups = num_kb >> 2;          // local groups; global group = g0 + g
        const int g0 = kb_begin >> 2;
        const int q = lane & 3, r8 = lane >> 2;
        const uint4* wbase[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int orow = nrow0 + (dw & 3) * 32 + 8 * i + r8;
          orow = orow < p.OC ? orow : 0;
          wbase[i] = reinterpret_cast<const uint4*>(p.w.weight + (size_t)orow * nblk * 4) + q;
        }
        const uint4* wl4 = reinterpret_cast<const uint4*>(wlrow);
        const uint32_t* zsw = reinterpret_cast<const uint32_t*>(p.w.zeros_and_scales) + (size_t)ocs * 32 * nchunk;
        const uint32_t* z2w = reinterpret_cast<const uint32_t*>(p.w.zeros_2nd) + (size_t)(ocs >> 2) * 32 * nchunk;
        const uint2* s2v2 = reinterpret_cast<const uint2*>(s2row);
        uint4 pw[4], pwl, pzs, pz2;
        uint2 ps2[3];
        auto fetch = [&](int gl) {
          const int g = g0 + gl;
#pragma unroll
          for (int i = 0; i < 4; ++i) pw[i] = __ldg(wbase[i] + 4 * g);
          pwl = __ldg(wl4 + g);
          const int kb0 = 4 * g;
          const int mword = (kb0 >> 6) * 32 + (kb0 & 31);
          pzs = __ldg(reinterpret_cast<const uint4*>(zsw + mword));
          pz2 = __ldg(reinterpret_cast<const uint4*>(z2w + mword));
#pragma unroll
          for (int i = 0; i < 3; ++i) ps2[i] = __ldg(s2v2 + 3 * g + i);
        };
        if (set < ngroups) fetch(set);
        for (int g = set; g < ngroups; g += 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(stg + (8 * i + r8) * STG_PITCH + q * 16) = pw[i];
          const uint4 cwl = pwl, czs = pzs, cz2 = pz2;
          const uint2 cs2[3] = {ps2[0], ps2[1], ps2[2]};
          __syncwarp();
          uint4 wq[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) wq[kk] = *reinterpret_cast<const uint4*>(stg + lane * STG_PITCH + kk * 16);
          __syncwarp();
          if (g + 2 < ngroups && !(p.dbg & 64)) fetch(g + 2);   // dbg 64: profiling, reuse the first group's words
          const int hsh = (((4 * (g0 + g)) & 63) >> 5) * 16;
          const uint32_t wlw[4] = {cwl.x, cwl.y, cwl.z, cwl.w};
          const uint32_t zsv[4] = {czs.x, czs.y, czs.z, czs.w};
          const uint32_t z2v[4] = {cz2.x, cz2.y, cz2.z, cz2.w};
          const uint32_t s2w[6] = {cs2[0].x, cs2[0].y, cs2[1].x, cs2[1].y, cs2[2].x, cs2[2].y};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            float sv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const int hidx = 3 * kk + k;
              const uint32_t wsel = s2w[hidx >> 1];
              sv[k] = __half2float(__ushort_as_half((unsigned short)((hidx & 1) ? (wsel >> 16) : (wsel & 0xFFFF))));
            }
            produce(4 * g + kk, wq[kk], wlw[kk], (zsv[kk] >> hsh) & 0xFFFF, (z2v[kk] >> (hsh >> 1)) & 0xFF,
                    sv[0], sv[1], sv[2]);
          }
        }
      } else {
        // generic K (IC % 256 != 0): per-block scalar loads, K blocks [4g, 4g+4) go to set g & 1
        for (int kb = 0; kb < num_kb; ++kb) {
          if (((kb >> 2) & 1) != set) continue;
          const int chunk = kb >> 6, bp = kb & 63;
          const int word = chunk * 32 + (bp & 31), hw = bp >> 5;
          const uint4 cwq = __ldg(wrow + kb);
          const uint32_t cwl = (uint32_t)__ldg(wlrow + kb);
          const uint32_t czs = __ldg(zsrow + word * 2 + hw);
          const uint32_t cz2 = __ldg(z2row + word * 4 + hw);
          const __half* sp = s2row + (size_t)kb * 3;
          produce(kb, cwq, cwl, czs, cz2, __half2float(__ldg(sp)), __half2float(__ldg(sp + 1)),
                  __half2float(__ldg(sp + 2)));
        }
      }
    }
    // ===== epilogue: TMEM -> registers -> fp16 -> global (K slice: fp32 -> exchange buffer) =====
    mbar_wait_dbg(tmem_full, 0, p.dbg_host, 7, 0, true);
    tc_fence_after();
    const int half = dw >> 2;                 // accumulator (token half)
    const int quad = warp & 3;                // TMEM lane quarter this warp may access
    const int token = m0 + half * 128 + quad * 32 + lane;
    const size_t yoff = (size_t)token * p.ldy + p.col0 + n0;
    float4* mine = partial ? p.partial + ((size_t)((tile - p.full) * p.split + slice) * 2 + rank) * SLOT_F4 + (threadIdx.x - 64)
                           : nullptr;
    if (!partial && (p.mc || p.npeers > 1)) {
      // Exchange epilogue (fused column all-gather).  A thread owns a token row, so direct stores
      // are 64-byte pieces, one NVLink packet per 16 bytes.  The operand ring is free once
      // tmem_full has fired: each warp transposes its 32 rows x 512 B through it and stores whole
      // 512-byte row segments -- to every peer, or once to the multicast address.
      constexpr int PITCH = BN * 2 + 16;
      uint8_t* stg = smem + dw * (32 * PITCH);
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN + cb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            oh[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
          *reinterpret_cast<uint4*>(stg + lane * PITCH + cb * 64 + q * 16) = o;
        }
      }
      __syncwarp();
      const int tok0 = m0 + half * 128 + quad * 32;
      const int col = n0 + lane * 8;
      if (col + 8 <= p.OC) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          if (tok0 + r >= p.M) break;
          const uint4 o = *reinterpret_cast<const uint4*>(stg + r * PITCH + lane * 16);
          const size_t off = (size_t)(tok0 + r) * p.ldy + p.col0 + col;
          if (p.mc) {
            multimem_st_16(p.y[0] + off, o);
          } else {
            for (int pe = 0; pe < p.npeers; ++pe) *reinterpret_cast<uint4*>(p.y[pe] + off) = o;
          }
        }
      }
    } else
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN + cb * 32, v);
      tmem_ld_wait();
      if (partial) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          __stcg(mine + (cb * 8 + j) * 256, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
      } else if (token < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = n0 + cb * 32 + q * 8;
          if (col + 8 <= p.OC) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              oh[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
            *reinterpret_cast<uint4*>(p.y[0] + yoff + cb * 32 + q * 8) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();            // neither CTA exits (or frees TMEM) while the pair's MMAs may read it
  if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace pair

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

// fp16 [rows, cols] row-major, box = [box_rows x 64 cols], 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MXQ_E_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MXQ_OK : MXQ_E_UNSUPPORTED;
}

// Tile schedule of the CTA-pair kernel.  Tiles are 512 tokens x 256 weight rows, one cluster each,
// dispatched M-fastest.  With T tiles on P SM pairs the last T % P tiles would run as a partly
// filled wave while the other pairs idle (11008 x 4096 at M = 2048: 172 tiles on 74 pairs = 2.3
// waves that cost 3).  Those tail tiles are cut along K into `split` slices so that the last wave
// is full and 1 / split as long; the slices leave fp32 partials in an exchange buffer (L2-resident)
// that a second, GPU-wide kernel adds up (an in-kernel "last arriver reduces" fix-up was measured
// first: one CTA reading (split - 1) x 256 KB is bound by its own SM's L2 bandwidth and cost as
// much as the wave it saved).
struct Plan {
  int mt, nt, tiles, full, split;
  size_t partial_bytes;
};
constexpr size_t kPartialSlotBytes = (size_t)pair::BM * pair::BN * 4;   // one CTA's fp32 accumulators
static Plan make_plan_for(int M, int IC, int OC, int sms, bool allow_split) {
  Plan pl{};
  pl.mt = ceil_div(M, 2 * pair::BM);
  pl.nt = ceil_div(OC, pair::BN);
  pl.tiles = pl.mt * pl.nt;
  pl.full = pl.tiles;
  pl.split = 1;
  if (allow_split && IC % (4 * BK) == 0 && sms >= 2) {
    const int pairs = sms / 2;
    const int tail = pl.tiles % pairs;
    const int groups = IC / (4 * BK);
    if (tail > 0) {
      int smax = groups / 4;                   // a slice keeps >= 16 K blocks of main loop
      if (smax > 8) smax = 8;
      int s = pairs / tail;
      if (s > smax) s = smax;
      // (More than half a wave of tiles, e.g. 43 on 74 pairs, cut 3 ways = 2 waves of a third was
      // measured too: 62-66 us against 52-55 us whole -- 66 MB of partials cost more than the idle
      // pairs.  Only tails whose slices fit ONE wave are cut.)
      if (const char* e = getenv("MXQ_GEMM_SPLIT")) s = atoi(e) < s ? atoi(e) : s;   // profiling knob
      if (s >= 2) {
        pl.split = s;
        pl.full = pl.tiles - tail;
        pl.partial_bytes = (size_t)tail * s * 2 * kPartialSlotBytes;
      }
    }
  }
  return pl;
}
static Plan make_plan(int M, int IC, int OC, bool allow_split) {
  int dev = 0, sms = 0;
  if (!allow_split || cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    sms = 0;
  return make_plan_for(M, IC, OC, sms, allow_split);
}
static size_t plan_workspace_bytes(const Plan& pl) {
  return pl.split > 1 ? 256 + pl.partial_bytes : 256;
}

// CTA-pair kernel: 1-D grid of clusters of 2 (see Plan); workspace (optional) enables the K split
// force_split > 0: EVERY tile is cut into `force_split` K slices (phased exchange, see mxq_gemm_partials)
// and only the pair kernel runs; the second pass is launched separately (mxq_gemm_reduce_store).
static Plan make_plan_forced(int M, int OC, int split) {
  Plan pl{};
  pl.mt = ceil_div(M, 2 * pair::BM);
  pl.nt = ceil_div(OC, pair::BN);
  pl.tiles = pl.mt * pl.nt;
  pl.full = 0;
  pl.split = split;
  pl.partial_bytes = (size_t)pl.tiles * split * 2 * kPartialSlotBytes;
  return pl;
}

template <bool kDenseB>
static int launch_pair(const void* x, const Params& p, cudaStream_t st, void* workspace = nullptr,
                       size_t workspace_bytes = 0, int force_split = 0) {
  CUtensorMap mx, mw;
  int rc = make_map(&mx, x, p.M, p.IC, pair::BM);
  if (rc) return rc;
  if (kDenseB) {
    rc = make_map(&mw, p.wdense, p.OC, p.IC, pair::BNH);
    if (rc) return rc;
  } else {
    mw = mx;
  }
  auto k = pair::gemm_mxq_pair_kernel<kDenseB>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  Params pp = p;
  Plan pl = force_split > 0 ? make_plan_forced(p.M, p.OC, force_split) : make_plan(p.M, p.IC, p.OC, workspace != nullptr);
  if (pl.split > 1 || force_split > 0) {
    const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
    if (base + pl.partial_bytes > reinterpret_cast<uintptr_t>(workspace) + workspace_bytes) {
      if (force_split > 0) return MXQ_E_WORKSPACE;
      pl = make_plan(p.M, p.IC, p.OC, false);          // workspace too small: whole tiles only
    } else {
      pp.partial = reinterpret_cast<float4*>(base);
    }
  }
  pp.mt = pl.mt; pp.full = pl.full; pp.split = pl.split;
  const unsigned clusters = (unsigned)(pl.full + (pl.tiles - pl.full) * pl.split);
  dim3 grid(2u * clusters);
  pp.dbg = 0;
  pp.dbg_host = nullptr;
  if (const char* e = getenv("MXQ_GEMM_DBG")) pp.dbg = atoi(e);
#ifdef MXQ_DEBUG
  if (const char* e = getenv("MXQ_GEMM_DBG_PTR")) pp.dbg_host = (unsigned long long*)strtoull(e, nullptr, 0);
#endif
  k<<<grid, THREADS, pair::SMEM_BYTES, st>>>(mx, mw, pp);
  if (pl.split > 1 && force_split == 0) {
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    pair::OutPtrs out{};
    for (int i = 0; i < pp.npeers; ++i) out.y[i] = pp.y[i];
    out.npeers = pp.npeers; out.mc = pp.mc;
    pair::gemm_split_reduce_kernel<<<(unsigned)(pl.tiles - pl.full) * 16u, 256, 0, st>>>(
        pp.partial, out, pl.split, pl.full, pl.mt, pp.M, pp.OC, pp.ldy, pp.col0);
  }
  MXQ_LAUNCH_RESULT();
}

template <bool kDenseB>
static int launch(const void* x, const Params& p, cudaStream_t st, void* workspace = nullptr,
                  size_t workspace_bytes = 0) {
  // More than one M tile: CTA pairs (M = 256 MMAs).  MXQ_GEMM_SINGLE forces the one-CTA kernel.
  // M <= 256 has only OC / 256 tiles (16 for a 4096-row linear): with a workspace the pair kernel cuts
  // them along K and fills the machine (half of each M = 256 MMA is padding there, the K split is
  // worth more: profiles/r1_probe_gemm_smallm.txt).
  if (!getenv("MXQ_GEMM_SINGLE")) {
    if (p.M > BM) return launch_pair<kDenseB>(x, p, st, workspace, workspace_bytes);
    if (!kDenseB && workspace && make_plan(p.M, p.IC, p.OC, true).split >= 2)
      return launch_pair<kDenseB>(x, p, st, workspace, workspace_bytes);
  }
  CUtensorMap mx, mw;
  int rc = make_map(&mx, x, p.M, p.IC, BM);
  if (rc) return rc;
  if (kDenseB) {
    rc = make_map(&mw, p.wdense, p.OC, p.IC, BN);
    if (rc) return rc;
  } else {
    mw = mx;
  }
  auto k = gemm_mxq_kernel<kDenseB>;
  Params pp = p;
  pp.dbg = 0;
  if (const char* e = getenv("MXQ_GEMM_DBG")) pp.dbg = atoi(e);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)ceil_div(p.OC, BN), (unsigned)ceil_div(p.M, BM));
  k<<<grid, THREADS, SMEM_BYTES, st>>>(mx, mw, pp);
  MXQ_LAUNCH_RESULT();
}

}  // namespace gemm
}  // namespace mxq

using namespace mxq;

extern "C" size_t mxq_gemm_workspace_bytes(int64_t M, int64_t IC, int64_t OC) {
  if (M <= 0 || IC <= 0 || OC <= 0 || M > INT32_MAX || OC > INT32_MAX || IC > (1 << 24)) return 256;
  return gemm::plan_workspace_bytes(gemm::make_plan((int)M, (int)IC, (int)OC, true));
}

extern "C" int mxq_gemm_plan(int64_t M, int64_t IC, int64_t OC, int sms, int32_t* out5, size_t* workspace_bytes) {
  if (M <= 0 || IC <= 0 || OC <= 0 || IC % 64 || M > INT32_MAX || OC > INT32_MAX || IC > (1 << 24) || sms < 0) return MXQ_E_SHAPE;
  if (!out5) return MXQ_E_NULL;
  const gemm::Plan pl = gemm::make_plan_for((int)M, (int)IC, (int)OC, sms, true);
  out5[0] = pl.mt; out5[1] = pl.nt; out5[2] = pl.tiles; out5[3] = pl.full; out5[4] = pl.split;
  if (workspace_bytes) *workspace_bytes = gemm::plan_workspace_bytes(pl);
  return MXQ_OK;
}

extern "C" int mxq_gemm(const void* x, mxq_packed_t w, void* y, int64_t M, int64_t IC, int64_t OC,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (M < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(y);
  MXQ_CHECK_PTR(w.weight);
  if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b || !w.zeros_4b)
    return MXQ_E_NULL;
  if (IC % 64 || IC == 0 || OC % 8 || M > INT32_MAX || OC > INT32_MAX || IC > (1 << 24)) return MXQ_E_SHAPE;
  gemm::Params p{};
  p.w = w; p.y[0] = (__half*)y; p.npeers = 1; p.ldy = (int)OC; p.col0 = 0;
  p.M = (int)M; p.IC = (int)IC; p.OC = (int)OC;
  return gemm::launch<false>(x, p, as_stream(stream), workspace, workspace_bytes);
}

extern "C" int mxq_gemm_scatter(const void* x, mxq_packed_t w, void* const* y_peers, int npeers,
                                int64_t M, int64_t IC, int64_t OC, int64_t ldy, int64_t col0,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (M < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(w.weight);
  if (!y_peers || !w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b || !w.zeros_4b)
    return MXQ_E_NULL;
  if (npeers < 1 || npeers > gemm::MAX_PEERS) return MXQ_E_SHAPE;
  if (IC % 64 || IC == 0 || OC % 8 || ldy % 8 || col0 % 8 || col0 + OC > ldy || ldy > INT32_MAX) return MXQ_E_SHAPE;
  if (M > INT32_MAX || IC > (1 << 24)) return MXQ_E_SHAPE;
  gemm::Params p{};
  p.w = w;
  for (int i = 0; i < npeers; ++i) {
    MXQ_CHECK_PTR(y_peers[i]);
    p.y[i] = (__half*)y_peers[i];
  }
  p.npeers = npeers; p.ldy = (int)ldy; p.col0 = (int)col0;
  p.M = (int)M; p.IC = (int)IC; p.OC = (int)OC;
  return gemm::launch<false>(x, p, as_stream(stream), workspace, workspace_bytes);
}

extern "C" int mxq_gemm_multicast(const void* x, mxq_packed_t w, void* y_multicast, int64_t M, int64_t IC,
                                  int64_t OC, int64_t ldy, int64_t col0, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (M < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(y_multicast);
  MXQ_CHECK_PTR(w.weight);
  if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b || !w.zeros_4b)
    return MXQ_E_NULL;
  if (IC % 64 || IC == 0 || OC % 8 || ldy % 8 || col0 % 8 || col0 + OC > ldy || ldy > INT32_MAX) return MXQ_E_SHAPE;
  if (M > INT32_MAX || IC > (1 << 24)) return MXQ_E_SHAPE;
  gemm::Params p{};
  p.w = w; p.y[0] = (__half*)y_multicast; p.npeers = 1; p.mc = 1; p.ldy = (int)ldy; p.col0 = (int)col0;
  p.M = (int)M; p.IC = (int)IC; p.OC = (int)OC;
  // the CTA-pair kernel carries the multicast epilogue; short M runs it too (rows beyond M are masked)
  return gemm::launch_pair<false>(x, p, as_stream(stream), workspace, workspace_bytes);
}

// ---- phased exchange (column-sharded GEMM whose shard is a single partly filled wave) ----------------
// At 8 ranks a 70B shard has 16-56 tiles on 74 SM pairs: every tile finishes at the same time and the
// exchange stores of the fused epilogue (114 us of NVLink ingress per rank for gate/up) follow 93 us of
// tensor work instead of overlapping it.  The caller cuts the shard's weight rows into groups of N tiles
// and runs, per group, (1) mxq_gemm_partials: the pair kernel with every tile cut into `split` K slices
// so that the group alone fills the machine -- fp32 partials go to the workspace -- and (2)
// mxq_gemm_reduce_store on a second stream: slices added in slice order, fp16 rows stored to the peers /
// the multicast mapping.  Group g's exchange then runs under group g+1's tensor work.
extern "C" size_t mxq_gemm_partials_workspace_bytes(int64_t M, int64_t OC, int split) {
  if (M <= 0 || OC <= 0 || split < 1 || M > INT32_MAX || OC > INT32_MAX) return 256;
  return 256 + gemm::make_plan_forced((int)M, (int)OC, split).partial_bytes;
}

extern "C" int mxq_gemm_partials(const void* x, mxq_packed_t w, int64_t M, int64_t IC, int64_t OC, int split,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (M < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(w.weight);
  MXQ_CHECK_PTR(workspace);
  if (!w.weight_last || !w.zeros_and_scales || !w.zeros_2nd || !w.scales_2nd || !w.scales_4b || !w.zeros_4b)
    return MXQ_E_NULL;
  if (IC % (4 * gemm::BK) || IC == 0 || OC % 8 || M > INT32_MAX || OC > INT32_MAX || IC > (1 << 24)) return MXQ_E_SHAPE;
  if (split < 1 || split > IC / (4 * gemm::BK)) return MXQ_E_SHAPE;
  gemm::Params p{};
  p.w = w; p.y[0] = nullptr; p.npeers = 1; p.ldy = (int)OC; p.col0 = 0;
  p.M = (int)M; p.IC = (int)IC; p.OC = (int)OC;
  return gemm::launch_pair<false>(x, p, as_stream(stream), workspace, workspace_bytes, split);
}

extern "C" int mxq_gemm_reduce_store(const void* workspace, size_t workspace_bytes, void* const* y_peers, int npeers,
                                     void* y_multicast, int64_t M, int64_t OC, int split, int64_t ldy, int64_t col0,
                                     void* stream) {
  if (M < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(workspace);
  if (M > INT32_MAX || OC > INT32_MAX || OC % 8 || ldy % 8 || col0 % 8 || col0 + OC > ldy || ldy > INT32_MAX || split < 1)
    return MXQ_E_SHAPE;
  const gemm::Plan pl = gemm::make_plan_forced((int)M, (int)OC, split);
  const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  if (base + pl.partial_bytes > reinterpret_cast<uintptr_t>(workspace) + workspace_bytes) return MXQ_E_WORKSPACE;
  gemm::pair::OutPtrs out{};
  if (y_multicast) {
    if (reinterpret_cast<uintptr_t>(y_multicast) & 15) return MXQ_E_ALIGN;
    out.y[0] = (__half*)y_multicast; out.npeers = 1; out.mc = 1;
  } else {
    if (!y_peers || npeers < 1 || npeers > gemm::MAX_PEERS) return MXQ_E_SHAPE;
    for (int i = 0; i < npeers; ++i) {
      MXQ_CHECK_PTR(y_peers[i]);
      out.y[i] = (__half*)y_peers[i];
    }
    out.npeers = npeers; out.mc = 0;
  }
  gemm::pair::gemm_split_reduce_kernel<<<(unsigned)pl.tiles * 16u, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(base), out, split, 0, pl.mt, (int)M, (int)OC, (int)ldy, (int)col0);
  MXQ_LAUNCH_RESULT();
}

extern "C" int mxq_gemm_dense(const void* x, const void* W, void* y, int64_t M, int64_t IC, int64_t OC,
                              void* stream) {
  if (M < 0 || IC < 0 || OC < 0) return MXQ_E_SHAPE;
  if (M == 0 || OC == 0) return MXQ_OK;
  MXQ_CHECK_PTR(x);
  MXQ_CHECK_PTR(W);
  MXQ_CHECK_PTR(y);
  if (IC % 64 || IC == 0 || OC % 8) return MXQ_E_SHAPE;
  gemm::Params p{};
  p.wdense = (const __half*)W; p.y[0] = (__half*)y; p.npeers = 1; p.ldy = (int)OC; p.col0 = 0;
  p.M = (int)M; p.IC = (int)IC; p.OC = (int)OC;
  return gemm::launch<true>(x, p, as_stream(stream));
}
