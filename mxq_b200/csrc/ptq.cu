// Fused MXQGPT.fasterquant(blocksize=16) + Quantizer (sm_100a).
//
// Replaces mxq_quant/lib/mxqgpt.py:387-448 and mxq_quant/lib/quantizer.py:5-20,61-121,149-155:
// 192 Quantizer() constructions and ~30 ATen calls per 64-column block become two kernels:
//   (A) pool_minmax : per row min/max over the pooled (4-bit) columns of W with dead columns
//                     zeroed (reads one 32-byte sector in four);
//   (B) ptq_tile    : one warp per (16-row tile, 4-group block): lane = (row, half-group),
//                     group min/max by one xor-shuffle, the 16-row second-level scale
//                     quantisation (quantizer.py:114-121) by xor-shuffles 2,4,8,16, then
//                     quantize/dequantize and 128-bit stores of the fp16 fake-quant weight.
// HBM traffic: read W 1.25x, write Wq 1x (algorithmic 4 B / weight).
// All arithmetic is op-by-op IEEE fp32 like the reference; x/scale uses a correctly-rounded
// reciprocal + Markstein step (bit-identical to division, oracle/div_check.c).
#include "common.cuh"

namespace mxq {

__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// quantizer.py:81-99: scale, zero from (xmin, xmax) with the degenerate fix
__device__ __forceinline__ void find_params(float xmin, float xmax, float maxq, float& scale,
                                            float& zero) {
  if (xmin == xmax) { xmin = -1.f; xmax = 1.f; }
  scale = fdiv(__fsub_rn(xmax, xmin), maxq);
  zero = fdiv(-xmin, scale);
}

// quantizer.py:114-121 over 16 lanes that differ in lane bits 1..4 (same bit 0)
__device__ __forceinline__ float qq_scale_16rows(float scale, int qq_maxq) {
  float smin = scale, smax = scale;
#pragma unroll
  for (int o = 2; o <= 16; o <<= 1) {
    smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
  }
  float s2, z2;
  find_params(smin, smax, (float)qq_maxq, s2, z2);
  float q = rintf(__fadd_rn(fdiv(scale, fmaxf(s2, 1e-9f)), z2));
  q = fminf(fmaxf(q, 0.f), (float)qq_maxq);
  return __fmul_rn(s2, __fsub_rn(q, z2));
}

__global__ void dead_mask_kernel(const float* __restrict__ colstat, uint8_t* __restrict__ dead,
                                 int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) dead[c] = colstat ? (colstat[c] == 0.f) : 0;
}

__device__ __forceinline__ void load_chunk_f16(const __half* W, const uint8_t* dead, size_t row_off,
                                               int col, float* f) {
  const uint4 ch = *reinterpret_cast<const uint4*>(W + row_off + col);
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

// (A) one warp per row
template <bool kRef>
__global__ void __launch_bounds__(256) pool_minmax_kernel(const __half* __restrict__ W,
                                                          const uint8_t* __restrict__ dead,
                                                          const uint8_t* __restrict__ gbits,
                                                          float2* __restrict__ out, int rows,
                                                          int cols, int group) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int cpg = group / 8;  // 16-byte chunks per group
  const int nchunks = cols / 8;
  float mn = INFINITY, mx = -INFINITY;
  const size_t roff = (size_t)row * cols;
  if (kRef) {
    const int npc = nchunks / 4;
    for (int m = lane; m < npc; m += 32) {
      const int c = ((m / cpg) * 4 + 3) * cpg + (m % cpg);
      float f[8];
      load_chunk_f16(W, dead, roff, c * 8, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
    }
  } else {
    for (int c = lane; c < nchunks; c += 32) {
      if (gbits[c / cpg] & MXQ_POOL_FLAG) {
        float f[8];
        load_chunk_f16(W, dead, roff, c * 8, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { mn = fminf(mn, f[e]); mx = fmaxf(mx, f[e]); }
      }
    }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if (lane == 0) out[row] = make_float2(mn, mx);
}

// (B) group == 16 (2 chunks per group): warp unit = 16 rows x 4 groups (64 columns)
template <bool kRef>
__global__ void __launch_bounds__(256) ptq_tile_kernel(const __half* __restrict__ W,
                                                       __half* __restrict__ Wq,
                                                       uint8_t* __restrict__ codes,
                                                       const uint8_t* __restrict__ dead,
                                                       const uint8_t* __restrict__ gbits,
                                                       const float2* __restrict__ pool_mm,
                                                       int rows, int cols, int low_bits,
                                                       int pool_bits) {
  const int lane = threadIdx.x & 31;
  const int r = lane >> 1, h = lane & 1;
  const int nblk = cols / 64;
  const int64_t units = (int64_t)(rows / 16) * nblk;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t wstride = (int64_t)gridDim.x * 8;
  const float maxq_low = (float)((1 << low_bits) - 1);
  const float maxq_pool = (float)((1 << pool_bits) - 1);
  int cur_tile = -1;
  float pool_scale = 0.f, pool_zero = 0.f;
  for (int64_t u = warp0; u < units; u += wstride) {
    const int tile = (int)(u / nblk), blk = (int)(u % nblk);
    const int row = tile * 16 + r;
    const size_t roff = (size_t)row * cols;
    if (tile != cur_tile) {  // per-row pool parameters + their 16-row second level
      cur_tile = tile;
      const float2 mm = pool_mm[row];
      find_params(mm.x, mm.y, maxq_pool, pool_scale, pool_zero);
      pool_scale = qq_scale_16rows(pool_scale, 15);
    }
    float f[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) load_chunk_f16(W, dead, roff, blk * 64 + k * 16 + h * 8, f[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = blk * 4 + k;
      const bool pooled = kRef ? (k == 3) : ((gbits[g] & MXQ_POOL_FLAG) != 0);
      float mn = f[k][0], mx = f[k][0];
#pragma unroll
      for (int e = 1; e < 8; ++e) { mn = fminf(mn, f[k][e]); mx = fmaxf(mx, f[k][e]); }
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      float scale, zero, maxq;
      find_params(mn, mx, maxq_low, scale, zero);
      scale = qq_scale_16rows(scale, 15);  // executed by all lanes (shuffles), selected below
      maxq = maxq_low;
      if (pooled) { scale = pool_scale; zero = pool_zero; maxq = maxq_pool; }
      const float sc = fmaxf(scale, 1e-9f);
      const float rc = __frcp_rn(sc);
      float o[8];
      uint32_t cw[2] = {0, 0};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float q = rintf(__fadd_rn(div_rn_by(f[k][e], sc, rc), zero));
        q = fminf(fmaxf(q, 0.f), maxq);
        o[e] = __fmul_rn(scale, __fsub_rn(q, zero));
        cw[e >> 2] |= (uint32_t)q << (8 * (e & 3));
      }
      const size_t off = roff + blk * 64 + k * 16 + h * 8;
      *reinterpret_cast<uint4*>(Wq + off) = DT<__half>::pack(o);
      if (codes) *reinterpret_cast<uint2*>(codes + off) = make_uint2(cw[0], cw[1]);
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

extern "C" size_t mxq_ptq_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows < 0 || cols < 0) return 0;
  return align16((size_t)cols) + align16((size_t)rows * sizeof(float2)) + 16;
}

extern "C" int mxq_ptq_quant(const void* W, void* Wq, uint8_t* codes, const float* colstat,
                             int64_t rows, int64_t cols, int group, int low_bits,
                             const uint8_t* group_bits, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (rows < 0 || cols < 0) return MXQ_E_SHAPE;
  if (rows == 0 || cols == 0) return MXQ_OK;
  MXQ_CHECK_PTR(W);
  MXQ_CHECK_PTR(Wq);
  MXQ_CHECK_PTR(workspace);
  if (group != 16) return MXQ_E_UNSUPPORTED;  // fasterquant is only ever called with blocksize=16
  if (rows % 16 || cols % 64) return MXQ_E_SHAPE;
  if (low_bits < 1 || low_bits > 8) return MXQ_E_SHAPE;
  if (rows > INT32_MAX || cols > (1 << 24)) return MXQ_E_SHAPE;
  if (workspace_bytes < mxq_ptq_workspace_bytes(rows, cols)) return MXQ_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  uint8_t* dead = (uint8_t*)workspace;
  float2* pool_mm = (float2*)((uint8_t*)workspace + align16((size_t)cols));
  dead_mask_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, st>>>(colstat, dead, (int)cols);
  const unsigned gridA = (unsigned)ceil_div(rows, 8);
  const int64_t units = (rows / 16) * (cols / 64);
  int64_t gridB = ceil_div(units, 8);
  if (gridB > kNumSMs * 8) gridB = kNumSMs * 8;
  const __half* w = (const __half*)W;
  if (group_bits == nullptr) {
    pool_minmax_kernel<true><<<gridA, 256, 0, st>>>(w, dead, nullptr, pool_mm, (int)rows, (int)cols, group);
    ptq_tile_kernel<true><<<(unsigned)gridB, 256, 0, st>>>(w, (__half*)Wq, codes, dead, nullptr, pool_mm,
                                                           (int)rows, (int)cols, low_bits, 4);
  } else {
    pool_minmax_kernel<false><<<gridA, 256, 0, st>>>(w, dead, group_bits, pool_mm, (int)rows, (int)cols, group);
    ptq_tile_kernel<false><<<(unsigned)gridB, 256, 0, st>>>(w, (__half*)Wq, codes, dead, group_bits, pool_mm,
                                                            (int)rows, (int)cols, low_bits, 4);
  }
  MXQ_LAUNCH_RESULT();
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
