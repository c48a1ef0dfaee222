#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE, int CHAINS>
__global__ void k(int iters, int* out, long long* cyc) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
  int di[CHAINS][4] = {};
  float df[CHAINS][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (MODE == 0) imma(di[c], a, b); else hmma(df[c], a, b);
    }
  }
  long long t1 = clock64();
  int s = 0;
  for (int c = 0; c < CHAINS; ++c) for (int j = 0; j < 4; ++j) s += di[c][j] + (int)df[c][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  int* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) k<0, 4><<<148, warps * 32>>>(iters, out, cyc); else k<1, 4><<<148, warps * 32>>>(iters, out, cyc);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = (double)h[0] / (iters * 4.0 * warps);
      printf("%s warps/SM=%d: %.2f SM-cycles per mma (per-warp %.2f)  err=%s\n", mode == 0 ? "IMMA m16n8k32 u8s8" : "HMMA m16n8k16 f16", warps, c, c * warps, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
