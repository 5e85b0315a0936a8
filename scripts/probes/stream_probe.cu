// Probe: what rate can 148 persistent CTAs pull through a TMA-bulk shared-memory ring when the consumers do nothing?
// Usage: stream_probe <total_MB> <stage_KB> <stages> <copies_per_stage> <consumer_warps> <touch>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(544, 1) probe(const unsigned char* base, size_t per_cta, int stage_bytes, int S, int copies, int cw, int touch, float* sink, int loops, int hint, int pad) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], cw); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const unsigned char* src = base + (size_t)blockIdx.x * per_cta;
  const int nst1 = (int)(per_cta / stage_bytes), nst = nst1 * loops;
  if (warp == 16) {
    uint64_t pol = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    for (int st = 0; st < nst; ++st) {
      const int slot = st % S;
      if (st >= S) mbar_wait(&empty[slot], ((st / S) - 1) & 1);
      if (lane == 0) mbar_expect(&full[slot], stage_bytes);
      __syncwarp();
      const int cb = stage_bytes / copies;
      if (lane < copies) {
        unsigned char* d = ring + (size_t)slot * (stage_bytes + 16 * pad * copies) + lane * (cb + 16 * pad);
        const unsigned char* g = src + (size_t)(st % nst1) * stage_bytes + lane * cb;
        if (hint) bulk_hint(d, g, cb, &full[slot], pol); else bulk(d, g, cb, &full[slot]);
      }
    }
  } else if (warp < cw) {
    float acc = 0.f;
    for (int st = 0; st < nst; ++st) {
      const int slot = st % S;
      mbar_wait(&full[slot], (st / S) & 1);
      if (touch) {
        const uint4* p = reinterpret_cast<const uint4*>(ring + (size_t)slot * stage_bytes);
        for (int i = warp * 32 + lane; i < stage_bytes / 16; i += cw * 32) { uint4 v = p[i]; acc += __uint_as_float(v.x) + __uint_as_float(v.w); }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 123.456f) *sink = acc;
  }
}
int main(int argc, char** argv) {
  const size_t total_mb = argc > 1 ? atol(argv[1]) : 3072;
  const int stage_kb = argc > 2 ? atoi(argv[2]) : 32, S = argc > 3 ? atoi(argv[3]) : 5, copies = argc > 4 ? atoi(argv[4]) : 1;
  const int cw = argc > 5 ? atoi(argv[5]) : 16, touch = argc > 6 ? atoi(argv[6]) : 0;
  const int reps = argc > 7 ? atoi(argv[7]) : 1;
  const int loops = argc > 8 ? atoi(argv[8]) : 1;
  const int hint = argc > 9 ? atoi(argv[9]) : 0, pad = argc > 10 ? atoi(argv[10]) : 0;
  const int grid = 148, stage = stage_kb * 1024;
  size_t per_cta = (total_mb << 20) / grid / stage * stage;
  unsigned char* buf; float* sink;
  CK(cudaMalloc(&buf, per_cta * grid)); CK(cudaMemset(buf, 1, per_cta * grid)); CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, S * (stage + 16 * pad * copies)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) probe<<<grid, 544, S * (stage + 16 * pad * copies)>>>(buf, per_cta, stage, S, copies, cw, touch, sink, loops, hint, pad);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("total %zu MB stage %d KB x %d, %d copies/stage, %d consumer warps, touch %d, reps %d hint %d pad %d: %.3f ms  %.1f GB/s\n", total_mb, stage_kb, S, copies, cw, touch, reps, hint, pad,
         best, (double)per_cta * grid * reps * loops / best / 1e6);
  return 0;
}
