// Issue-rate probe: FHFMA (fma.rn.f32.f16) vs FFMA vs HFMA2 vs LOP3, 8 independent chains per thread,
// 1 CTA of 512 threads per SM (4 warps per scheduler).  Prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int MODE>
__global__ void k(float* out, uint32_t seed, long long* cyc) {
  float a[8]; uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; h[i] = seed + i * 0x00010001u + threadIdx.x; }
  unsigned short xs = (unsigned short)(seed & 0x3fff), ws = (unsigned short)((seed >> 3) & 0x3fff);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < N; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(ws), "h"(xs));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
      if (MODE == 2) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(seed), "r"(seed + 1));
      if (MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h[i]) : "r"(seed), "r"(seed + 7));
      if (MODE == 4) { asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(ws), "h"(xs)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h[i]) : "r"(seed), "r"(seed + 7)); }
      if (MODE == 6) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(seed), "r"(seed + 7));
      if (MODE == 7) { asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(seed), "r"(seed + 7)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h[(i + 4) & 7]) : "r"(seed), "r"(seed + 7)); }
      if (MODE == 8) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(seed), "r"(seed + 7)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h[(i + 4) & 7]) : "r"(seed), "r"(seed + 7)); }
      if (MODE == 9) { asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(seed), "r"(seed + 1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h[(i + 4) & 7]) : "r"(seed), "r"(seed + 7)); }
      if (MODE == 5) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h[i]) : "r"(seed), "r"(seed + 7)); }
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t x = 0;
  for (int i = 0; i < 8; ++i) { s += a[i]; x ^= h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"FHFMA", "FFMA", "HFMA2", "LOP3", "FHFMA+LOP3", "FFMA+LOP3", "IDP4A", "IDP4A+LOP3", "IMAD+LOP3", "HFMA2+LOP3"};
  for (int m = 0; m < 10; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (m) { case 0: k<0><<<148, 512>>>(out, 123, cyc); break; case 1: k<1><<<148, 512>>>(out, 123, cyc); break;
        case 2: k<2><<<148, 512>>>(out, 123, cyc); break; case 3: k<3><<<148, 512>>>(out, 123, cyc); break;
        case 4: k<4><<<148, 512>>>(out, 123, cyc); break; case 5: k<5><<<148, 512>>>(out, 123, cyc); break; case 6: k<6><<<148, 512>>>(out, 123, cyc); break; case 7: k<7><<<148, 512>>>(out, 123, cyc); break; case 8: k<8><<<148, 512>>>(out, 123, cyc); break; case 9: k<9><<<148, 512>>>(out, 123, cyc); break; }
      cudaDeviceSynchronize();
    }
    const double instr_per_smsp = (double)N * 8 * ((m == 4 || m == 5 || m >= 7) ? 2 : 1) * 4;   // 4 warps per SMSP
    printf("%-12s %8lld cycles  %.2f cycles per warp-instr per SMSP\n", names[m], *cyc, *cyc / instr_per_smsp);
  }
  return 0;
}
