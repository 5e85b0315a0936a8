// Warp-level attention tile on mma.sync.m16n8k16 (bf16 in, fp32 accumulate), shared by the persistent decode step
// (decode_tc.cu) and the prefill attention kernel (attn_prefill.cu).
//
// Orientation: the 64 tokens of a K/V tile sit on the M dimension, the query columns on N = 8:
//     S^T[64 tok x 8 col] = K[64 x 128] q^T          (A = K tile via ldmatrix, B = q fragments kept in registers)
//     O^T[128 d x 8 col] += V^T[128 x 64] P^T        (A = V tile via ldmatrix.trans, B = P^T via movmatrix from S^T)
// so the accumulator fragments of S^T and O^T share their column ownership (thread t owns columns 2 (t % 4) + {0, 1})
// and the online-softmax rescale is thread-local.  A column is one (query token, query head) pair: 8 / G tokens of the
// G heads that share the kv head.  The tile lives in shared memory as two [64 tok][64 d] halves per tensor with the
// 128-byte TMA swizzle (16-byte chunk c of row r at c ^ (r & 7)), which makes every ldmatrix conflict-free.
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t movm_t(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// address of the 16-byte chunk c (0..15, 8 bf16 each) of token row r in a [64 tok][128 d] tile stored as two 128B-swizzled halves
__device__ __forceinline__ uint32_t kv_chunk_addr(uint32_t base, int r, int c) {
  return base + (uint32_t)((c >> 3) * 8192 + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}

// One 64-token tile.  kb / vb: shared addresses of the K and V tiles; qf: B fragments of the 8 query columns
// (b0 = q[col = lane / 4][d = 16 ks + 2 (lane % 4) + {0, 1}], b1 = ... + 8); lim[j]: tokens of the tile visible to this
// thread's column j (tokens >= lim[j] are masked; <= 0 masks the whole tile).  V rows the mask hides must be finite.
// State: o[dt][e] = O^T[d = 16 dt + lane / 4 + 8 (e >> 1)][col 2 (lane % 4) + (e & 1)], mrun / lrun per column (lrun is the
// thread's partial over its fragment rows: reduce over lane / 4 at the end).
__device__ __forceinline__ void attn_tile64(uint32_t kb, uint32_t vb, const uint32_t (&qf)[8][2], float scale, const int (&lim)[2], float (&o)[8][4],
                                            float (&mrun)[2], float (&lrun)[2], int lane) {
  const int lq = lane >> 2;
  float sc[4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    sc[mt][0] = sc[mt][1] = sc[mt][2] = sc[mt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t af[4];
      const int row = 16 * mt + ((lane >> 3) & 1) * 8 + (lane & 7);
      ldsm4(kv_chunk_addr(kb, row, 2 * ks + (lane >> 4)), af);
      mma16816(sc[mt], af, qf[ks][0], qf[ks][1]);
    }
  }
  // online softmax per column; element (mt, e): token 16 mt + lq + 8 (e >> 1), column (e & 1)
  float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int tok = 16 * mt + lq + 8 * (e >> 1);
      sc[mt][e] = tok < lim[e & 1] ? sc[mt][e] * scale : -INFINITY;
      tmax[e & 1] = fmaxf(tmax[e & 1], sc[mt][e]);
    }
  float fac[2], mref[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    tmax[j] = fmaxf(tmax[j], __shfl_xor_sync(0xffffffffu, tmax[j], 4));
    tmax[j] = fmaxf(tmax[j], __shfl_xor_sync(0xffffffffu, tmax[j], 8));
    tmax[j] = fmaxf(tmax[j], __shfl_xor_sync(0xffffffffu, tmax[j], 16));
    const float mn = fmaxf(mrun[j], tmax[j]);
    mref[j] = mn == -INFINITY ? 0.f : mn;
    fac[j] = __expf(mrun[j] - mref[j]);                    // first tile: exp(-inf) = 0
    mrun[j] = mn;
    lrun[j] *= fac[j];
  }
  uint32_t pf[4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const float p0 = __expf(sc[mt][0] - mref[0]), p1 = __expf(sc[mt][1] - mref[1]);
    const float p2 = __expf(sc[mt][2] - mref[0]), p3 = __expf(sc[mt][3] - mref[1]);
    lrun[0] += p0 + p2; lrun[1] += p1 + p3;
    pf[mt][0] = movm_t(pack_bf16(p0, p1));                // B fragment of O^T += V^T P^T: k = token, n = column
    pf[mt][1] = movm_t(pack_bf16(p2, p3));
  }
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    o[dt][0] *= fac[0]; o[dt][1] *= fac[1]; o[dt][2] *= fac[0]; o[dt][3] *= fac[1];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t af[4];
      const int row = 16 * kk + (lane >> 4) * 8 + (lane & 7);
      ldsm4t(kv_chunk_addr(vb, row, 2 * dt + ((lane >> 3) & 1)), af);
      mma16816(o[dt], af, pf[kk][0], pf[kk][1]);
    }
  }
}

}  // namespace
