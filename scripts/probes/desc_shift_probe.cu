// Can a K-major SWIZZLE_128B UMMA operand start at a row that is not a multiple of 8 (a convolution tap shifting the
// time rows of one resident halo tile)?  A row r holds r + 1 in its first k element, B rows hold 1: D[r][*] = first_row + r + 1.
//   desc_shift_probe <shift rows> <base_offset field>
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I zonos_b200/csrc -I include scripts/probes/desc_shift_probe.cu -o scripts/probes/desc_shift_probe
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"

__global__ void __launch_bounds__(128, 1) probe(int shift, int boff, float* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  bf16* A = (bf16*)smem; bf16* B = (bf16*)(smem + 32 * 1024);
  for (int i = threadIdx.x; i < 48 * 1024 / 2; i += blockDim.x) A[i] = __float2bfloat16(0.f);
  __syncthreads();
  for (int r = threadIdx.x; r < 160; r += blockDim.x)     // the swizzle is a function of the ABSOLUTE row index (address bits 7..9)
    A[(r / 8) * 512 + (r % 8) * 64 + ((0 ^ (r % 8)) << 3)] = __float2bfloat16((float)(r + 1));
  for (int r = threadIdx.x; r < 16; r += blockDim.x) B[(r / 8) * 512 + (r % 8) * 64 + ((0 ^ (r % 8)) << 3)] = __float2bfloat16(1.f);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  fence_proxy_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp_id_uniform() == 1) {
    if (elect_one()) {
      uint64_t da = make_smem_desc(smem_u32(A) + (uint32_t)shift * 128u) | ((uint64_t)(boff & 7) << 49);
      tc_mma(tmem, da, make_smem_desc(smem_u32(B)), make_idesc(128, 16), 0u);
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[8];
  tc_ld8(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
  out[threadIdx.x] = v[0];
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
  }
}

int main(int argc, char** argv) {
  const int shift = argc > 1 ? atoi(argv[1]) : 1, boff = argc > 2 ? atoi(argv[2]) : 0;
  float* out; cudaMallocManaged(&out, 128 * 4);
  const size_t smem = 49 * 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(shift, boff, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  int ok = 0;
  for (int r = 0; r < 128; ++r) ok += out[r] == (float)(r + shift + 1);
  printf("shift %2d base_offset %d: %3d of 128 rows as expected; rows 0..11 ->", shift, boff, ok);
  for (int r = 0; r < 12; ++r) printf(" %g", out[r]);
  printf("\n");
  return 0;
}
