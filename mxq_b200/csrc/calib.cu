// Calibration statistics for the mxq PTQ pass (sm_100a).
//
// Replaces MXQGPT.add_batch (mxq_quant/lib/mxqgpt.py:369-383), whose K x K fp32 Hessian
// (68.7 GFLOP per 2048-token sample at K=4096) is only ever read as diag(H) == 0
// (mxqgpt.py:399-403), and WrappedGPT.add_batch (mxq_quant/lib/layerwrapper.py:22-35), by one
// HBM-bound column sum-of-squares: X is read exactly once, 2 bytes per element.
//
// Layout: X[tokens, cols] row-major.  A thread owns one 16-byte column chunk and walks down a
// slab of rows with 8 independent 128-bit loads in flight; a warp reads 512 contiguous bytes per
// row.  Partial sums go to workspace[slab][col] and a second tiny kernel folds the slabs in a
// fixed order (deterministic, no atomics).
#include <cstdlib>

#include "common.cuh"

namespace mxq {

constexpr int kCSThreads = 256;

// kMinBlocks is a register budget (launch bounds).  8 (32 registers): ptxas keeps ~3 loads in flight
// per thread and the kernel relies on full occupancy (8 CTAs / SM) for its memory parallelism.
// 4 or 2: all kCSUnroll loads are front-loaded (56 registers at 8, 96 at 16), so 3 CTAs / SM keep as
// many bytes in flight and leave room for a co-resident kernel.
template <typename T, int kCSUnroll = 8, int kMinBlocks = 8>
__global__ void __launch_bounds__(kCSThreads, kMinBlocks) colsumsq_partial_kernel(
    const uint8_t* __restrict__ X, float* __restrict__ partial, int64_t tokens, int cols, int cpr,
    int rows_per_slab) {
  using D = DT<T>;
  constexpr int EPC = D::EPC;
  const int chunk = blockIdx.x * kCSThreads + threadIdx.x;
  if (chunk >= cpr) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  int64_t r1 = r0 + rows_per_slab;
  if (r1 > tokens) r1 = tokens;
  const size_t row_bytes = (size_t)cols * sizeof(T);
  const uint8_t* p = X + (size_t)r0 * row_bytes + (size_t)chunk * 16;
  float acc[EPC];
#pragma unroll
  for (int e = 0; e < EPC; ++e) acc[e] = 0.f;
  int64_t r = r0;
  for (; r + kCSUnroll <= r1; r += kCSUnroll) {
    uint4 v[kCSUnroll];
#pragma unroll
    for (int u = 0; u < kCSUnroll; ++u) v[u] = ld_stream(p + (size_t)u * row_bytes);
    p += (size_t)kCSUnroll * row_bytes;
#pragma unroll
    for (int u = 0; u < kCSUnroll; ++u) {
      float f[EPC];
      D::unpack(v[u], f);
#pragma unroll
      for (int e = 0; e < EPC; ++e) acc[e] = fmaf(f[e], f[e], acc[e]);
    }
  }
  for (; r < r1; ++r) {
    float f[EPC];
    D::unpack(ld_stream(p), f);
    p += row_bytes;
#pragma unroll
    for (int e = 0; e < EPC; ++e) acc[e] = fmaf(f[e], f[e], acc[e]);
  }
  float* dst = partial + (size_t)blockIdx.y * cols + (size_t)chunk * EPC;
#pragma unroll
  for (int e = 0; e < EPC; e += 4)
    *reinterpret_cast<float4*>(dst + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
}

// ---------------------------------------------------------------------------------------------
// Same statistic, bytes in flight held in SHARED memory instead of registers: one CTA per SM takes a
// slab of whole rows; R consecutive rows are one contiguous byte range, fetched by ONE 1-D TMA bulk
// copy (cp.async.bulk + mbarrier tx-count) into a ring of up to 8 stages (<= 192 KB per SM in
// flight).  Thread t owns 16-byte column chunks t, t + 256, ... (conflict-free LDS.128) and keeps
// their sums in registers.  The CTA needs ~10 K registers, so the issue-bound quantize+pack tile
// kernel of the previous layer can run ALONGSIDE on the same SMs (two of its CTAs fit).  Measured: it
// does, but this kernel is then -- and even alone, 5.0 TB/s -- bound by its 8 consumer warps
// (LDS -> convert -> FMA chains, 2 warps per scheduler), not by the ring; kept as the measured
// alternative (mxq_colsumsq_ex ctas_per_sm = 0), not used by default.
// ---------------------------------------------------------------------------------------------
constexpr int kRingStageMax = 32 * 1024;
constexpr int kRingBytesMax = 192 * 1024;
constexpr int kRingStagesMax = 8;

template <typename T, int CH>
__global__ void __launch_bounds__(kCSThreads, 1) colsumsq_ring_kernel(
    const uint8_t* __restrict__ X, float* __restrict__ partial, int64_t tokens, int cols, int cpr,
    int rows_per_slab, int rows_per_stage, int nstages) {
  using D = DT<T>;
  constexpr int EPC = D::EPC;
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ uint64_t full[kRingStagesMax], empty[kRingStagesMax];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_slab;
  const int64_t r1 = r0 + rows_per_slab < tokens ? r0 + rows_per_slab : tokens;
  const size_t row_bytes = (size_t)cols * sizeof(T);
  const uint32_t stage_bytes = (uint32_t)(rows_per_stage * row_bytes);
  const int nfill = r1 > r0 ? (int)((r1 - r0 + rows_per_stage - 1) / rows_per_stage) : 0;
  auto fill = [&](int f) {      // thread 0: rows [r0 + f R, ...) -> stage f % nstages
    const int64_t ra = r0 + (int64_t)f * rows_per_stage;
    const int64_t rb = ra + rows_per_stage < r1 ? ra + rows_per_stage : r1;
    const uint32_t bytes = (uint32_t)((rb - ra) * row_bytes);
    uint64_t* bar = &full[f % nstages];
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(ring + (size_t)(f % nstages) * stage_bytes, X + (size_t)ra * row_bytes, bytes, bar);
  };
  if (tid == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kCSThreads / 32); }
    mbar_fence_init();
    for (int f = 0; f < nstages && f < nfill; ++f) fill(f);
  }
  __syncthreads();
  float acc[CH][EPC];
#pragma unroll
  for (int j = 0; j < CH; ++j)
#pragma unroll
    for (int e = 0; e < EPC; ++e) acc[j][e] = 0.f;
  for (int f = 0; f < nfill; ++f) {
    const int s = f % nstages;
    const uint32_t ph = (uint32_t)(f / nstages) & 1u;
    mbar_wait(&full[s], ph);
    const int64_t ra = r0 + (int64_t)f * rows_per_stage;
    const int nr = (int)((ra + rows_per_stage < r1 ? ra + rows_per_stage : r1) - ra);
    const uint8_t* st = ring + (size_t)s * stage_bytes + (size_t)tid * 16;
    for (int r = 0; r < nr; ++r) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (tid + j * kCSThreads < cpr) {
          float v[EPC];
          D::unpack(*reinterpret_cast<const uint4*>(st + (size_t)r * row_bytes + (size_t)j * (kCSThreads * 16)), v);
#pragma unroll
          for (int e = 0; e < EPC; ++e) acc[j][e] = fmaf(v[e], v[e], acc[j][e]);
        }
      }
    }
    // this warp is done with the stage; the producer refills it once all 8 warps are
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (tid == 0 && f + nstages < nfill) {
      mbar_wait(&empty[s], ph);
      fill(f + nstages);
    }
  }
  float* dst = partial + (size_t)blockIdx.x * cols;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int c = tid + j * kCSThreads;
    if (c < cpr) {
#pragma unroll
      for (int e = 0; e < EPC; e += 4)
        *reinterpret_cast<float4*>(dst + (size_t)c * EPC + e) =
            make_float4(acc[j][e], acc[j][e + 1], acc[j][e + 2], acc[j][e + 3]);
    }
  }
}

// fold the slabs: block = 32 columns x 32 slab lanes (coalesced 128-byte rows of `partial`), fixed
// summation order -> deterministic
__global__ void __launch_bounds__(1024) colsumsq_final_kernel(const float* __restrict__ partial,
                                                              float* __restrict__ out, int cols,
                                                              int slabs, float prev_scale,
                                                              float add_scale, int accumulate) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < cols) {
    int k = threadIdx.y;
    for (; k + 96 < slabs; k += 128) {
      const float a0 = partial[(size_t)k * cols + c], a1 = partial[(size_t)(k + 32) * cols + c];
      const float a2 = partial[(size_t)(k + 64) * cols + c], a3 = partial[(size_t)(k + 96) * cols + c];
      s += (a0 + a1) + (a2 + a3);
    }
    for (; k < slabs; k += 32) s += partial[(size_t)k * cols + c];
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    float v = add_scale * t;
    if (accumulate) v = fmaf(prev_scale, out[c], v);
    out[c] = v;
  }
}

template <typename T>
__global__ void wanda_metric_kernel(const T* __restrict__ W, const float* __restrict__ sr,
                                    float* __restrict__ out, int64_t rows, int cols) {
  const int64_t n = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % cols);
    out[i] = __fmul_rn(fabsf((float)W[i]), __fsqrt_rn(sr[c]));
  }
}


// ---------------------------------------------------------------------------------------------
// Importance-driven 2/4-bit allocation (the reference hard-wires "the last 16
// of every 64 columns are 4-bit", utils_quant.py:340-385 / mxqgpt.py:404-419, and computes the
// Wanda metric |W| * sqrt(scaler_row) only on its pruning paths, prune.py:177).
//   importance[g] = sum over rows and the group's columns of |W[r,c]| * sqrt(scaler_row[c])
// Of every 4 consecutive groups the most important one becomes the pooled 4-bit group.
// The column sums of |W| are EXACT: fp16 magnitudes are integer multiples of 2^-24, so they are
// accumulated as 64-bit integers (order-independent, atomics allowed); the per-group combination
// runs in fp64 in a fixed order.  Oracle and kernel therefore agree bit for bit and the mask never
// depends on the launch geometry.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colabs_i64_kernel(const __half* __restrict__ W,
                                                         unsigned long long* __restrict__ colabs,
                                                         int64_t rows, int cols, int rows_per_slab) {
  const int chunk = blockIdx.x * 256 + threadIdx.x;          // 8 columns
  if (chunk * 8 >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = min(rows, r0 + rows_per_slab);
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t r = r0; r < r1; ++r) {
    float f[8];
    DT<__half>::unpack(ld_stream(W + (size_t)r * cols + (size_t)chunk * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float m = fabsf(f[e]) * 16777216.0f;             // exact: |fp16| * 2^24 < 2^40
      acc[e] += (m <= 1.1e12f) ? (unsigned long long)__float2ll_rn(m) : 0ull;   // inf / nan weights: ignored
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) atomicAdd(colabs + (size_t)chunk * 8 + e, acc[e]);
}

__global__ void alloc_bits_kernel(const unsigned long long* __restrict__ colabs,
                                  const float* __restrict__ scaler_row, int cols, int group, int low_bits,
                                  uint8_t* __restrict__ group_bits, double* __restrict__ importance) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;     // 4 consecutive groups
  const int nblk = cols / (4 * group);
  if (blk >= nblk) return;
  double best = -1.0;
  int arg = 0;
  for (int k = 0; k < 4; ++k) {
    const int c0 = (blk * 4 + k) * group;
    double acc = 0.0;
    for (int j = 0; j < group; ++j) {
      const double a = __dmul_rn((double)colabs[c0 + j], 1.0 / 16777216.0);
      const double w = scaler_row ? __dsqrt_rn((double)scaler_row[c0 + j]) : 1.0;
      acc = __dadd_rn(acc, __dmul_rn(a, w));                  // never contracted: the oracle rounds twice
    }
    if (importance) importance[blk * 4 + k] = acc;
    if (acc > best) { best = acc; arg = k; }                 // ties -> lowest index
  }
  for (int k = 0; k < 4; ++k) group_bits[blk * 4 + k] = (uint8_t)(k == arg ? (MXQ_POOL_FLAG | 4) : low_bits);
}

static int slabs_for(int64_t tokens, int cpr, int ctas_per_sm = 8) {
  const int col_tiles = (int)ceil_div(cpr, kCSThreads);
  int slabs = (kNumSMs * ctas_per_sm) / col_tiles;  // ~8 CTAs of 256 threads per SM (one wave)
  if (slabs < 1) slabs = 1;
  const int64_t max_slabs = ceil_div(tokens, 64);
  if (slabs > max_slabs) slabs = (int)(max_slabs > 0 ? max_slabs : 1);
  return slabs;
}

}  // namespace mxq

using namespace mxq;

extern "C" size_t mxq_colsumsq_workspace_bytes(int64_t tokens, int64_t cols) {
  if (tokens <= 0 || cols <= 0) return 16;
  const int cpr_min = (int)(cols / 8) > 0 ? (int)(cols / 8) : 1;  // fp32 has more chunks -> fewer slabs
  // sized for the largest slab count mxq_colsumsq_ex may use (ctas_per_sm = 32)
  return (size_t)slabs_for(tokens, cpr_min, 32) * (size_t)cols * sizeof(float) + 16;
}

// ctas_per_sm = 0: the shared-memory ring kernel (one CTA per SM; made to share the SMs with the PTQ
// tile kernel of another stream), falling back to 8 for shapes it does not cover.
// ctas_per_sm >= 1: CTAs of the register-buffered partial kernel per SM.  8 (= mxq_colsumsq) is one full wave.  Smaller
// values are enforced with an unused dynamic shared-memory reservation and leave registers / thread
// slots free; larger values (up to 32) split the rows into more, shorter slabs = several waves, so
// CTAs retire continuously and a concurrent kernel on another (higher-priority) stream -- the PTQ
// tile kernel of the previous layer -- keeps getting SM slots.
extern "C" int mxq_colsumsq_ex(const void* X, int64_t tokens, int64_t cols, int dtype, float* out,
                               float prev_scale, float add_scale, int accumulate, int ctas_per_sm,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (tokens < 0 || cols <= 0 || ctas_per_sm < 0 || ctas_per_sm > 32) return MXQ_E_SHAPE;
  MXQ_CHECK_PTR(out);
  if (dtype != MXQ_F32 && dtype != MXQ_F16 && dtype != MXQ_BF16) return MXQ_E_DTYPE;
  const int esize = dtype == MXQ_F32 ? 4 : 2;
  if ((cols * esize) % 16 || cols > (1 << 24)) return MXQ_E_SHAPE;
  cudaStream_t st = as_stream(stream);
  const int cpr = (int)(cols * esize / 16);
  int slabs = 0;
  if (tokens > 0) {
    MXQ_CHECK_PTR(X);
    MXQ_CHECK_PTR(workspace);
    const size_t row_bytes = (size_t)cols * esize;
    const int ch = (int)ceil_div(cpr, kCSThreads);
    const char* re = getenv("MXQ_STAT_RING");                 // profiling knob: 0 forces the register kernel
    const bool ring = ctas_per_sm == 0 ? true : (re && atoi(re) == 1);
    const int ring_rows_per_slab = (int)ceil_div(tokens, kNumSMs);
    const int ring_slabs = (int)ceil_div(tokens, ring_rows_per_slab);
    if (ring && row_bytes <= (size_t)kRingStageMax * 2 && ch <= 8 && tokens >= kNumSMs && !(re && atoi(re) == 0) &&
        workspace_bytes >= (size_t)ring_slabs * cols * sizeof(float)) {
      // one CTA per SM, whole rows, TMA bulk ring (colsumsq_ring_kernel)
      const int rps = row_bytes >= (size_t)kRingStageMax ? 1 : (int)(kRingStageMax / row_bytes);
      const size_t stage_bytes = (size_t)rps * row_bytes;
      int nst = (int)(kRingBytesMax / stage_bytes);
      if (nst > kRingStagesMax) nst = kRingStagesMax;
      slabs = ring_slabs;
      const int rows_per_slab = ring_rows_per_slab;
      const size_t smem = (size_t)nst * stage_bytes;
      float* part = (float*)workspace;
      const uint8_t* Xb = (const uint8_t*)X;
      auto launch_ring = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)slabs, kCSThreads, smem, st>>>(Xb, part, tokens, (int)cols, cpr, rows_per_slab, rps, nst);
        return cudaSuccess;
      };
      auto by_ch = [&](auto tag) -> cudaError_t {
        using TT = decltype(tag);
        return ch <= 1 ? launch_ring(colsumsq_ring_kernel<TT, 1>) : ch <= 2 ? launch_ring(colsumsq_ring_kernel<TT, 2>)
             : ch <= 4 ? launch_ring(colsumsq_ring_kernel<TT, 4>) : ch <= 6 ? launch_ring(colsumsq_ring_kernel<TT, 6>)
                       : launch_ring(colsumsq_ring_kernel<TT, 8>);
      };
      cudaError_t e = dtype == MXQ_F32 ? by_ch(float{}) : dtype == MXQ_F16 ? by_ch(__half{}) : by_ch(__nv_bfloat16{});
      if (e != cudaSuccess) return (int)e;
    } else {
    if (ctas_per_sm == 0) ctas_per_sm = 8;
    slabs = slabs_for(tokens, cpr, ctas_per_sm);
    if (workspace_bytes < (size_t)slabs * cols * sizeof(float)) return MXQ_E_WORKSPACE;
    const int rows_per_slab = (int)ceil_div(tokens, slabs);
    slabs = (int)ceil_div(tokens, rows_per_slab);
    dim3 grid((unsigned)ceil_div(cpr, kCSThreads), (unsigned)slabs);
    float* part = (float*)workspace;
    const uint8_t* Xb = (const uint8_t*)X;
    // n CTAs per SM: n * (pad + 1 KB) <= 220 KB < (n + 1) * (pad + 1 KB); the last 8 KB stay free for
    // the CTAs (1 KB of reserved shared memory each) of a kernel that is meant to run alongside
    size_t pad = 0;
    if (ctas_per_sm < 8) pad = (((size_t)233472 - 8192) / ctas_per_sm - 1024) & ~(size_t)1023;
    auto launch = [&](auto kern) -> cudaError_t {
      if (pad > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        if (e != cudaSuccess) return e;
      }
      kern<<<grid, kCSThreads, pad, st>>>(Xb, part, tokens, (int)cols, cpr, rows_per_slab);
      return cudaSuccess;
    };
    // capped occupancy -> the variant whose threads really keep 16 loads in flight (96 registers: two
    // CTAs resident per SM hold 128 KB in flight); MXQ_STAT_UNROLL = 0 / 8 / 16 overrides (profiling)
    const char* ue = getenv("MXQ_STAT_UNROLL");
    const int variant = ue ? atoi(ue) : (ctas_per_sm < 8 ? 16 : 0);
    cudaError_t e;
    if (dtype == MXQ_F32) e = launch(colsumsq_partial_kernel<float>);
    else if (dtype == MXQ_BF16) e = variant == 16  ? launch(colsumsq_partial_kernel<__nv_bfloat16, 16, 2>)
                                    : variant == 8 ? launch(colsumsq_partial_kernel<__nv_bfloat16, 8, 4>)
                                                   : launch(colsumsq_partial_kernel<__nv_bfloat16>);
    else e = variant == 16  ? launch(colsumsq_partial_kernel<__half, 16, 2>)
             : variant == 8 ? launch(colsumsq_partial_kernel<__half, 8, 4>)
                            : launch(colsumsq_partial_kernel<__half>);
    if (e != cudaSuccess) return (int)e;
  }
    }
  colsumsq_final_kernel<<<(unsigned)ceil_div(cols, 32), dim3(32, 32), 0, st>>>(
      (const float*)workspace, out, (int)cols, slabs, prev_scale, add_scale, accumulate);
  MXQ_LAUNCH_RESULT();
}

extern "C" int mxq_colsumsq(const void* X, int64_t tokens, int64_t cols, int dtype, float* out,
                            float prev_scale, float add_scale, int accumulate, void* workspace,
                            size_t workspace_bytes, void* stream) {
  // 3 = sixteen loads in flight per thread, two CTAs resident per SM, 1.5 waves: 6.8 TB/s on a 2.1 GB
  // fp16 input against 6.4 TB/s for the full-occupancy variant (8); fp32 inputs only have that one
  return mxq_colsumsq_ex(X, tokens, cols, dtype, out, prev_scale, add_scale, accumulate, dtype == MXQ_F32 ? 8 : 3,
                         workspace, workspace_bytes, stream);
}

extern "C" int mxq_wanda_metric(const void* W, const float* scaler_row, float* out, int64_t rows,
                                int64_t cols, int dtype, void* stream) {
  if (rows < 0 || cols < 0) return MXQ_E_SHAPE;
  if (rows == 0 || cols == 0) return MXQ_OK;
  if (!W || !scaler_row || !out) return MXQ_E_NULL;
  cudaStream_t st = as_stream(stream);
  int64_t blocks = ceil_div(rows * cols, 256 * 4);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (dtype == MXQ_F32)
    wanda_metric_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)W, scaler_row, out, rows, (int)cols);
  else if (dtype == MXQ_F16)
    wanda_metric_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)W, scaler_row, out, rows, (int)cols);
  else if (dtype == MXQ_BF16)
    wanda_metric_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)W, scaler_row, out, rows, (int)cols);
  else
    return MXQ_E_DTYPE;
  MXQ_LAUNCH_RESULT();
}

extern "C" size_t mxq_allocate_bits_workspace_bytes(int64_t cols) {
  return cols > 0 ? (size_t)cols * sizeof(unsigned long long) : 16;
}

extern "C" int mxq_allocate_bits(const void* W, const float* scaler_row, int64_t rows, int64_t cols, int group,
                                 int low_bits, uint8_t* group_bits, double* importance, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (rows < 0 || cols <= 0 || group <= 0) return MXQ_E_SHAPE;
  if (cols % (4 * group) || cols % 8 || cols > (1 << 24) || low_bits < 1 || low_bits > 8) return MXQ_E_SHAPE;
  MXQ_CHECK_PTR(W);
  MXQ_CHECK_PTR(workspace);
  if (!group_bits) return MXQ_E_NULL;
  if (workspace_bytes < (size_t)cols * sizeof(unsigned long long)) return MXQ_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  unsigned long long* colabs = (unsigned long long*)workspace;
  cudaError_t e = cudaMemsetAsync(colabs, 0, (size_t)cols * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return (int)e;
  if (rows > 0) {
    const int col_tiles = (int)ceil_div(cols / 8, 256);
    int slabs = (kNumSMs * 8) / col_tiles;
    if (slabs < 1) slabs = 1;
    if (slabs > rows) slabs = (int)rows;
    const int rows_per_slab = (int)ceil_div(rows, slabs);
    slabs = (int)ceil_div(rows, rows_per_slab);
    colabs_i64_kernel<<<dim3((unsigned)col_tiles, (unsigned)slabs), 256, 0, st>>>((const __half*)W, colabs, rows,
                                                                               (int)cols, rows_per_slab);
  }
  const int nblk = (int)(cols / (4 * group));
  alloc_bits_kernel<<<(unsigned)ceil_div(nblk, 128), 128, 0, st>>>(colabs, scaler_row, (int)cols, group, low_bits,
                                                                   group_bits, importance);
  MXQ_LAUNCH_RESULT();
}
