// Which TMEM lane holds accumulator row r of a cta_group::1 tcgen05.mma with M = 64 (and M = 128)?
// A row r carries the value r + 1 in its first k element, B rows carry 1 there: D[r][n] = r + 1.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I zonos_b200/csrc -I include scripts/probes/tmem_layout_probe.cu -o scripts/probes/tmem_layout_probe
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"

__global__ void __launch_bounds__(128, 1) probe(int M, int N, float* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  bf16* A = (bf16*)smem; bf16* B = (bf16*)(smem + 32 * 1024);
  for (int i = threadIdx.x; i < 64 * 1024 / 2; i += blockDim.x) A[i] = __float2bfloat16(0.f);
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += blockDim.x) {
    A[(r / 8) * 512 + (r % 8) * 64 + ((0 ^ (r % 8)) << 3)] = __float2bfloat16((float)(r + 1));
    B[(r / 8) * 512 + (r % 8) * 64 + ((0 ^ (r % 8)) << 3)] = __float2bfloat16(1.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  fence_proxy_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp_id_uniform() == 1) {
    if (elect_one()) {
      tc_mma(tmem, make_smem_desc(smem_u32(A)), make_smem_desc(smem_u32(B)), make_idesc(M, N), 0u);
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v[8];
  tc_ld8(tmem + ((uint32_t)(warp * 32) << 16), v);
  out[threadIdx.x * 2] = v[0]; out[threadIdx.x * 2 + 1] = v[1];
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
  (void)lane;
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 16;
  float* out; cudaMallocManaged(&out, 256 * 4);
  const size_t smem = 65 * 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(M, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  printf("M %d N %d: TMEM lane -> accumulator row + 1 (column 0 | column 1)\n", M, N);
  for (int l = 0; l < 128; ++l) printf("%s%3d:%3.0f|%3.0f", l % 8 ? "  " : "\n", l, out[2 * l], out[2 * l + 1]);
  printf("\n");
  return 0;
}
