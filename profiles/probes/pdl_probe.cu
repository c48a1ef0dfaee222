// PDL mechanics probe: a chain of kernels that (1) TMA-load their 43 KB share before
// griddepcontrol.wait, (2) wait, (3) "compute" for a fixed number of cycles, (4) exit.
// Variants: trigger (launch_dependents) at the top / after the wait / never; CTA size.
// Prints per-kernel time and, from %globaltimer stamps, when each phase happened.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smid() { uint32_t s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }

struct Stamp { unsigned long long start, loaded, waited, done; uint32_t sm, pad; };

// mode: 0 = no trigger, 1 = trigger after wait, 2 = trigger at top
__global__ void __launch_bounds__(448, 2) k(const unsigned char* __restrict__ p, size_t per, int compute_cycles, int mode, Stamp* stamps, uint32_t* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  unsigned long long t_start = gtimer();
  if (mode == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)per) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)), "l"(p + (size_t)blockIdx.x * per), "r"((uint32_t)per), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long t_waited = gtimer();
  if (mode == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  unsigned long long t_loaded = gtimer();
  uint32_t acc = 0;
  long long c0 = clock64();
  int i = threadIdx.x;
  while (clock64() - c0 < compute_cycles) { acc ^= reinterpret_cast<const uint32_t*>(sm)[i % (per / 4)]; i += 449; }
  if (acc == 0x12345678u) out[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && stamps) { Stamp s; s.start = t_start; s.loaded = t_loaded; s.waited = t_waited; s.done = gtimer(); s.sm = smid(); s.pad = 0; stamps[blockIdx.x] = s; }
}

int main() {
  const int NK = 64, G = 147;
  const size_t per = 43008;
  uint32_t* out; CK(cudaMalloc(&out, 4));
  std::vector<unsigned char*> bufs(NK);
  for (auto& b : bufs) { CK(cudaMalloc(&b, per * G)); CK(cudaMemset(b, 1, per * G)); }
  Stamp* stamps; CK(cudaMalloc(&stamps, sizeof(Stamp) * G * NK));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per));
  for (int threads : {448, 256}) for (int compute : {0, 4000}) for (int mode : {0, 1, 2}) for (int pdl : {0, 1}) {
    if (!pdl && mode) continue;
    cudaStream_t st; CK(cudaStreamCreate(&st));
    auto launch = [&](int i) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(G); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = per; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl;
      CK(cudaLaunchKernelEx(&cfg, k, (const unsigned char*)bufs[i], per, compute, mode, stamps + (size_t)i * G, out));
    };
    for (int i = 0; i < NK; ++i) launch(i);
    CK(cudaStreamSynchronize(st));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < NK; ++i) launch(i);
    CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, st));
    for (int r = 0; r < 5; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    std::vector<Stamp> h(G * NK);
    CK(cudaMemcpy(h.data(), stamps, sizeof(Stamp) * G * NK, cudaMemcpyDeviceToHost));
    // kernel 40: phases relative to kernel 39's last "done"
    auto last_done = [&](int kx) { unsigned long long m = 0; for (int c = 0; c < G; ++c) m = std::max(m, h[kx * G + c].done); return m; };
    const int kx = 40;
    const unsigned long long ref = last_done(kx - 1);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int smcount[160] = {0}; int maxper = 0;
    for (int c = 0; c < G; ++c) { const Stamp& s = h[kx * G + c]; s0 += (double)s.start - ref; s1 += (double)s.waited - ref; s2 += (double)s.loaded - ref; s3 += (double)s.done - ref; smcount[s.sm % 160]++; }
    for (int i = 0; i < 160; ++i) maxper = std::max(maxper, smcount[i]);
    printf("thr %3d compute %4d mode %d pdl %d: %6.2f us/kernel | k40 avg rel. to k39 end (ns): start %7.0f waited %6.0f loaded %6.0f done %6.0f | span %5.0f | max CTAs/SM %d\n",
           threads, compute, mode, pdl, ms * 1e3 / 5 / NK, s0 / G, s1 / G, s2 / G, s3 / G, (double)last_done(kx) - ref, maxper);
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaStreamDestroy(st));
  }
  return 0;
}
