// How fast does one thread get tcgen05.mma instructions with a SMALL N through the tensor pipe?
//   mma_rate_probe <M> <N> <count> <nacc> <distinct>
// nacc accumulators (TMEM column ranges) are used round robin; distinct = 1: every MMA reads a different k slice / tile.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I zonos_b200/csrc -I include scripts/probes/mma_rate_probe.cu -o /tmp/mma_rate_probe
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"

__global__ void __launch_bounds__(128, 1) probe(int M, int N, int count, int nacc, int distinct, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // bf16 ~0.0078
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  fence_proxy_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
  if (distinct >= 4 && warp_u == 1) {
    // the whole warp runs the loop on uniform values; one elected lane issues (CUTLASS / DeepGEMM pattern)
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 32 * 1024;
    const uint32_t bstep = (uint32_t)N * 128u >> 4;
    const int wrap = max(1, min(16, (128 * 1024) / (N * 128)));     // k blocks before the descriptors wrap
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      uint64_t da = make_smem_desc(a_base), db = make_smem_desc(b_base);
      for (int kb = 0; kb < count / 4; ++kb) {
        if (elect_one()) {
          tc_mma(tmem, da, db, idesc, 1u);
          tc_mma(tmem, da + 2, db + 2, idesc, 1u);
          tc_mma(tmem, da + 4, db + 4, idesc, 1u);
          tc_mma(tmem, da + 6, db + 6, idesc, 1u);
        }
        __syncwarp();
        da += 64; db += bstep;
        if ((kb + 1) % wrap == 0) { da -= 64 * wrap; db -= bstep * wrap; }
      }
      const long long t1 = clock64();
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, (uint32_t)(rep & 1));
      const long long t2 = clock64();
      if (blockIdx.x == 0 && threadIdx.x == 32) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
    }
  }
  if (distinct < 4 && threadIdx.x == 32) {
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 64 * 1024;       // A: 64 KB region, B: 96 KB region
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      if (distinct == 2) {                                    // loop-invariant descriptors, unrolled: the bare issue rate
        const uint64_t da = make_smem_desc(a_base), db = make_smem_desc(b_base);
#pragma unroll 8
        for (int i = 0; i < count; ++i) tc_mma(tmem, da, db, idesc, 1u);
      } else if (distinct == 3) {                             // realistic tight loop: 4 k slices per k block, descriptors advanced by adds
        uint64_t da = make_smem_desc(a_base), db = make_smem_desc(b_base);
        const uint32_t bstep = (uint32_t)N * 128u >> 4;
        for (int kb = 0; kb < count / 4; ++kb) {
          tc_mma(tmem, da, db, idesc, 1u);
          tc_mma(tmem, da + 2, db + 2, idesc, 1u);
          tc_mma(tmem, da + 4, db + 4, idesc, 1u);
          tc_mma(tmem, da + 6, db + 6, idesc, 1u);
          da += 64; db += bstep;
          if ((kb & 15) == 15) { da -= 64 * 16; db -= bstep * 16; }
        }
      } else
      for (int i = 0; i < count; ++i) {
        const int kb = distinct ? (i / 4) % 32 : 0, kk = i % 4;
        const uint64_t da = make_smem_desc(a_base + kb * 1024) + (uint64_t)(kk * 2);
        const uint64_t db = make_smem_desc(b_base + kb * (uint32_t)N * 128u % (64 * 1024)) + (uint64_t)(kk * 2);
        tc_mma(tmem + (uint32_t)((i % nacc) * 64), da, db, idesc, i >= nacc ? 1u : 0u);
      }
      const long long t1 = clock64();
      tc_commit(&bar);
      mbar_wait(&bar, (uint32_t)(rep & 1));
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 128, N = argc > 2 ? atoi(argv[2]) : 16, count = argc > 3 ? atoi(argv[3]) : 128;
  const int nacc = argc > 4 ? atoi(argv[4]) : 1, distinct = argc > 5 ? atoi(argv[5]) : 1, grid = argc > 6 ? atoi(argv[6]) : 1;
  long long* out; cudaMallocManaged(&out, 64);
  const size_t smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<grid, 128, smem>>>(M, N, count, nacc, distinct, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  printf("M %3d N %3d count %4d nacc %d distinct %d grid %3d: issue %6lld cyc (%5.1f / mma), done %6lld cyc (%5.1f / mma)\n", M, N, count, nacc, distinct, grid,
         out[4], (double)out[4] / count, out[5], (double)out[5] / count);
  return 0;
}
