// Probe: HBM rate of 148 persistent CTAs streaming a K-major weight matrix through a shared-memory ring, by access
// pattern.  The question behind it: does a 2-D tensor TMA box {64 k, RB rows} (128-byte pieces, row stride K*2 bytes -
// the layout tcgen05 wants) stream as fast as contiguous 1-D bulk copies, or must the weights be pre-tiled in HBM?
//   mode 0  1-D bulk, one contiguous run of RB*128*kps bytes per stage (the upper bound)
//   mode 1  2-D tensor map (SWIZZLE_128B), kps boxes {64, RB} per stage, a CTA walks K for its row block
//   mode 2  1-D bulk, pre-tiled emulation: RB/8 copies of kps KB per stage ([row group of 8][k block][8][64] layout)
//   mode 3  3-D tensor map {64, rows, K/64}, one box {64, RB, kps} per stage
// Usage: tma_tensor_probe <mode> <RB> <kps> <stages> <K> <total_MB> <mma N (0 = no tensor-core consumer)> <l2promo 0..3>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma2(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Args { const unsigned char* base; long long rows; int K, RB, kps, S, mode, nmma, nblocks, loops; float* sink; };

__global__ void __launch_bounds__(96, 1) probe(const __grid_constant__ CUtensorMap map, const Args a) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  __shared__ uint32_t tmem_s;
  unsigned char* ring = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = a.RB * 128 * a.kps;
  unsigned char* btile = ring + (size_t)a.S * stage_bytes;     // 16 KB of whatever: the B operand
  if (threadIdx.x == 0) { for (int s = 0; s < a.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1 && a.nmma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int nkb = a.K / 64, spb = nkb / a.kps;                 // stages per row block
  if (warp == 0 && lane == 0) {
    int st = 0;
    for (int lp = 0; lp < a.loops; ++lp)
    for (int blk = blockIdx.x; blk < a.nblocks; blk += gridDim.x) {
      const long long row0 = (long long)blk * a.RB;
      for (int j = 0; j < spb; ++j, ++st) {
        const int slot = st % a.S;
        if (st >= a.S) mbar_wait(&empty[slot], ((st / a.S) - 1) & 1);
        unsigned char* dst = ring + (size_t)slot * stage_bytes;
        mbar_expect(&full[slot], stage_bytes);
        const int kb = j * a.kps;
        if (a.mode == 0) {
          bulk(dst, a.base + ((size_t)blk * spb + j) * stage_bytes, stage_bytes, &full[slot]);
        } else if (a.mode == 1) {
          for (int i = 0; i < a.kps; ++i) tma2(dst + (size_t)i * a.RB * 128, &map, (kb + i) * 64, (int)row0, &full[slot]);
        } else if (a.mode == 2) {
          for (int g = 0; g < a.RB / 8; ++g)
            bulk(dst + (size_t)g * a.kps * 1024, a.base + (((size_t)(row0 / 8 + g)) * nkb + kb) * 1024, a.kps * 1024, &full[slot]);
        } else {
          tma3(dst, &map, 0, (int)row0, kb, &full[slot]);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    int st = 0;
    const uint32_t tmem = a.nmma ? tmem_s : 0;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.nmma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int lp = 0; lp < a.loops; ++lp)
    for (int blk = blockIdx.x; blk < a.nblocks; blk += gridDim.x) {
      for (int j = 0; j < spb; ++j, ++st) {
        const int slot = st % a.S;
        mbar_wait(&full[slot], (st / a.S) & 1);
        if (a.nmma) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(ring + (size_t)slot * stage_bytes);
          for (int i = 0; i < a.kps; ++i) {
            // mode 2 layout: [group][kps][1 KB] -> 8-row groups are kps KB apart; others: [kps][RB][128 B] -> 1 KB apart
            const uint32_t abase = a.mode == 2 ? sa + i * 1024 : sa + i * a.RB * 128;
            const uint64_t da = make_desc(abase, a.mode == 2 ? a.kps * 1024 : 1024), db = make_desc(smem_u32(btile), 1024);
            for (int kk = 0; kk < 4; ++kk) mma(tmem, da + kk * 2, db + kk * 2, idesc, (st | i | kk) ? 1u : 0u);
          }
          commit(&empty[slot]);
        } else {
          mbar_arrive(&empty[slot]);
        }
      }
    }
    if (a.nmma) {      // drain: wait for the last commit
      const int last = (st - 1) % a.S;
      mbar_wait(&empty[last], ((st - 1) / a.S) & 1);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1 && a.nmma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(128) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  Args a;
  a.mode = argc > 1 ? atoi(argv[1]) : 1; a.RB = argc > 2 ? atoi(argv[2]) : 128; a.kps = argc > 3 ? atoi(argv[3]) : 1; a.S = argc > 4 ? atoi(argv[4]) : 8;
  a.K = argc > 5 ? atoi(argv[5]) : 2048;
  const size_t total_mb = argc > 6 ? atol(argv[6]) : 1024;
  a.nmma = argc > 7 ? atoi(argv[7]) : 0;
  const int promo = argc > 8 ? atoi(argv[8]) : 2;
  const int grid = argc > 9 ? atoi(argv[9]) : 148;
  a.loops = argc > 10 ? atoi(argv[10]) : 1;
  a.rows = (long long)((total_mb << 20) / ((size_t)a.K * 2)) / a.RB * a.RB;
  a.nblocks = (int)(a.rows / a.RB);
  a.nblocks = a.nblocks / grid * grid;                       // equal work per CTA
  a.rows = (long long)a.nblocks * a.RB;
  unsigned char* buf;
  const size_t bytes = (size_t)a.rows * a.K * 2;
  CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes)); CK(cudaMalloc(&a.sink, 4));
  a.base = buf;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (a.mode == 1 || a.mode == 3) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUresult r;
    if (a.mode == 1) {
      cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.rows}; cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)a.RB}; cuuint32_t es[2] = {1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[3] = {64, (cuuint64_t)a.rows, (cuuint64_t)(a.K / 64)}; cuuint64_t strides[2] = {(cuuint64_t)a.K * 2, 128};
      cuuint32_t box[3] = {64, (cuuint32_t)a.RB, (cuuint32_t)a.kps}; cuuint32_t es[3] = {1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  const size_t smem = (size_t)a.S * a.RB * 128 * a.kps + 16384 + 1024;
  if (smem > 227 * 1024) { printf("smem %zu too large\n", smem); return 1; }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    probe<<<grid, 96, smem>>>(map, a);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("mode %d RB %3d kps %d stages %2d (%3zu KB in flight) K %5d mmaN %3d promo %d grid %d loops %d: %.3f ms  %.0f GB/s\n", a.mode, a.RB, a.kps, a.S,
         (size_t)a.S * a.RB * 128 * a.kps / 1024, a.K, a.nmma, promo, grid, a.loops, best, (double)bytes * a.loops / best / 1e6);
  return 0;
}
