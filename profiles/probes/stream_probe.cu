// Ceiling probe: how fast can a short kernel stream a 6-17 MB buffer from HBM inside a CUDA-graph
// chain (distinct buffers, > L2 in total)?  Variants: LDG.128, cp.async ring, TMA bulk ring; each
// with and without programmatic dependent launch.   nvcc -arch=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int UNROLL, bool PDL>
__global__ void __launch_bounds__(512) k_ldg(const uint4* __restrict__ p, size_t n16, uint32_t* out) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n16; i += stride) { uint4 v = p[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (PDL) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  if (acc == 0x12345678u) out[0] = acc;
}

// TMA bulk: each CTA owns a contiguous range, one thread issues CHUNK-byte bulk copies into a ring
template <int CHUNK, int STAGES, bool PDL>
__global__ void __launch_bounds__(256) k_bulk(const unsigned char* __restrict__ p, size_t bytes, uint32_t* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar[STAGES];
  const size_t per = ((bytes / gridDim.x) / CHUNK) * CHUNK;
  const unsigned char* base = p + (size_t)blockIdx.x * per;
  const int nchunks = (int)(per / CHUNK);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < STAGES && s < nchunks; ++s) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + s * CHUNK)), "l"(base + (size_t)s * CHUNK), "r"(CHUNK), "r"(smem_u32(&bar[s])) : "memory");
    }
  }
  __syncthreads();
  if (PDL) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  uint32_t acc = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % STAGES;
    const uint32_t parity = (c / STAGES) & 1;
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
    const uint4* q = reinterpret_cast<const uint4*>(sm + s * CHUNK);
    for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) { uint4 v = q[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    __syncthreads();
    if (threadIdx.x == 0 && c + STAGES < nchunks) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + s * CHUNK)), "l"(base + (size_t)(c + STAGES) * CHUNK), "r"(CHUNK), "r"(smem_u32(&bar[s])) : "memory");
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
float time_chain(F launch, int nbuf, int reps) {
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaGraph_t g; cudaGraphExec_t ge;
  for (int i = 0; i < nbuf; ++i) launch(i, st);
  CK(cudaStreamSynchronize(st));
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < nbuf; ++i) launch(i, st);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, st));
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaStreamDestroy(st));
  return ms * 1e3f / reps / nbuf;
}

template <typename K, typename... A>
void launch_k(K k, dim3 grid, dim3 block, size_t smem, bool pdl, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, k, args...));
}

int main() {
  uint32_t* out; CK(cudaMalloc(&out, 4));
  for (size_t mb : {6, 17, 64}) {
    const size_t bytes = mb << 20;
    const int nbuf = (int)((600u << 20) / bytes);
    std::vector<unsigned char*> bufs(nbuf);
    for (auto& b : bufs) { CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 1, bytes)); }
    auto report = [&](const char* name, float us) { printf("%2zu MiB %-34s %7.2f us  %7.0f GB/s\n", mb, name, us, bytes / us / 1e3); };
    for (int cps : {1, 2, 4}) {
      char nm[64];
      snprintf(nm, 64, "ldg u4 %dx148x512", cps);
      report(nm, time_chain([&](int i, cudaStream_t st) { launch_k(k_ldg<4, false>, dim3(148 * cps), dim3(512), 0, false, st, (const uint4*)bufs[i], bytes / 16, out); }, nbuf, 5));
      snprintf(nm, 64, "ldg u8 %dx148x512", cps);
      report(nm, time_chain([&](int i, cudaStream_t st) { launch_k(k_ldg<8, false>, dim3(148 * cps), dim3(512), 0, false, st, (const uint4*)bufs[i], bytes / 16, out); }, nbuf, 5));
      snprintf(nm, 64, "ldg u8 %dx148x512 pdl", cps);
      report(nm, time_chain([&](int i, cudaStream_t st) { launch_k(k_ldg<8, true>, dim3(148 * cps), dim3(512), 0, true, st, (const uint4*)bufs[i], bytes / 16, out); }, nbuf, 5));
    }
    {
      CK(cudaFuncSetAttribute(k_bulk<8192, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
      CK(cudaFuncSetAttribute(k_bulk<8192, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
      CK(cudaFuncSetAttribute(k_bulk<16384, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 6));
      report("bulk 8K x8 1x148", time_chain([&](int i, cudaStream_t st) { launch_k(k_bulk<8192, 8, false>, dim3(148), dim3(256), 8192 * 8, false, st, (const unsigned char*)bufs[i], bytes, out); }, nbuf, 5));
      report("bulk 8K x8 1x148 pdl", time_chain([&](int i, cudaStream_t st) { launch_k(k_bulk<8192, 8, true>, dim3(148), dim3(256), 8192 * 8, true, st, (const unsigned char*)bufs[i], bytes, out); }, nbuf, 5));
      report("bulk 8K x8 2x148 pdl", time_chain([&](int i, cudaStream_t st) { launch_k(k_bulk<8192, 8, true>, dim3(296), dim3(256), 8192 * 8, true, st, (const unsigned char*)bufs[i], bytes, out); }, nbuf, 5));
      report("bulk 16K x6 1x148 pdl", time_chain([&](int i, cudaStream_t st) { launch_k(k_bulk<16384, 6, true>, dim3(148), dim3(256), 16384 * 6, true, st, (const unsigned char*)bufs[i], bytes, out); }, nbuf, 5));
    }
    for (auto& b : bufs) CK(cudaFree(b));
  }
  return 0;
}
