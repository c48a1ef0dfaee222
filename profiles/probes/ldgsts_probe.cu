// Probe: can per-warp private cp.async (LDGSTS) rings stream HBM at full rate?  Each warp owns a
// contiguous slice of a big buffer, copies it unit by unit (UNIT bytes, 16 B per lane-instruction)
// into its own R-slot shared-memory ring, and consumes each unit with LDS.128 + xor.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int UNIT, int R>
__global__ void __launch_bounds__(256) k(const uint4* __restrict__ src, size_t units_per_warp, uint32_t* out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int CH = UNIT / 16;                 // 16-byte chunks per unit
  unsigned char* ring = smem + (size_t)warp * R * UNIT;
  const size_t gw = (size_t)blockIdx.x * 8 + warp;
  const uint4* base = src + gw * units_per_warp * CH;
  auto issue = [&](size_t u) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (u % R) * UNIT);
    for (int c = lane; c < CH; c += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(base + u * CH + c) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int u = 0; u < R - 1 && u < (int)units_per_warp; ++u) issue(u);
  uint32_t acc = 0;
  for (size_t u = 0; u < units_per_warp; ++u) {
    if (u + R - 1 < units_per_warp) issue(u + R - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");
    __syncwarp();
    const uint4* s = reinterpret_cast<const uint4*>(ring + (u % R) * UNIT);
    for (int c = lane; c < CH; c += 32) { uint4 v = s[c]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    __syncwarp();
  }
  if (acc == 0x12345678u) out[0] = acc;
}
template <int UNIT, int R>
void run(const uint4* src, size_t bytes, uint32_t* out, int ctas_per_sm) {
  const int grid = 148 * ctas_per_sm;
  const size_t upw = bytes / ((size_t)grid * 8 * UNIT);
  const size_t smem = (size_t)8 * R * UNIT;
  cudaFuncSetAttribute(k<UNIT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(a);
    k<UNIT, R><<<grid, 256, smem>>>(src, upw, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  const double moved = (double)upw * grid * 8 * UNIT;
  printf("UNIT %5d R %d ctas/SM %d smem/CTA %6zu: %.1f us  %.0f GB/s  (%s)\n", UNIT, R, ctas_per_sm, smem, best * 1e3, moved / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const size_t bytes = (size_t)1200 << 20;
  uint4* src; uint32_t* out;
  cudaMalloc(&src, bytes); cudaMalloc(&out, 4);
  cudaMemset(src, 1, bytes);
  run<1536, 2>(src, bytes, out, 1); run<1536, 4>(src, bytes, out, 1); run<1536, 8>(src, bytes, out, 1);
  run<1536, 2>(src, bytes, out, 2); run<1536, 4>(src, bytes, out, 2); run<1536, 6>(src, bytes, out, 2);
  run<3072, 2>(src, bytes, out, 2); run<3072, 4>(src, bytes, out, 2);
  run<512, 8>(src, bytes, out, 2); run<512, 16>(src, bytes, out, 2);
  // small: one GEMV worth (6.3 MB) to see ramp cost
  const size_t small = (size_t)6300000;
  run<1536, 4>(src, small, out, 2); run<1536, 4>(src, small, out, 1);
  return 0;
}
