// tcgen05 / TMEM GEMM for the dense contractions of the path (prefill, decode with more than 4 activation rows):
//     Y[M, N] = X[M, K] * W[N, K]^T        bf16 operands, fp32 accumulation in tensor memory
// "Swap-AB" tiling: the 128 MMA rows are WEIGHT rows (a CTA owns 128 output features and streams their K-major
// rows once), the MMA N dimension is a block of BN <= 256 activation rows, so one UMMA shape serves M = 8 ... 256+
// activation rows and the accumulator D[feature, row] lives in 128 lanes x BN columns of TMEM.
//   warp 4   : TMA producer  (cp.async.bulk.tensor.2d, 128-byte swizzle, mbarrier complete_tx)
//   warp 5   : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, cta_group::1, UMMA 128 x BN x 16)
//   warps 0-3: epilogue: tcgen05.ld 32 lanes x 16 columns -> registers -> fused epilogue (same rounding points as
//              the GEMV path: bf16 Linear output, then residual / SiLU gate / RoPE + paged KV append / fp32 CFG mix)
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA).
#include <stdlib.h>

#include "tc.cuh"

namespace {

constexpr int TC_BM = 128;        // weight rows per CTA (UMMA M)
constexpr int TC_BK = 64;         // k per stage: 64 bf16 = one 128-byte swizzle row
constexpr int TC_THREADS = 192;

enum { TEPI_STORE = 0, TEPI_RESID = 1, TEPI_QKV = 2, TEPI_SILU = 3, TEPI_HEADS = 4 };

struct TcArgs {
  CUtensorMap map_w;    // weights  [N, K], box {64 k, 64 rows}
  CUtensorMap map_x;    // activations [M, K], box {64 k, BN rows}
  int M, N, K, BN, stages, epi;
  int F;                // SILU: value rows [0,F), gate rows [F,2F)
  bf16* y; long long ldy; const bf16* resid; long long ldr;
  // QKV
  int T, Hq, Hkv, hd, rope_interleaved, rope_len, max_pages;
  const float* rope; const int32_t* lengths; const int32_t* page_table; bf16* kv_layer; bf16* q_out;
  // HEADS
  int B; float cfg_scale; float* logits; int QV;
  // split-K: grid.z CTAs share one output tile; fp32 partials meet in `ws`, the last CTA to arrive reduces them in
  // split order (deterministic) and runs the epilogue
  int ksplit; float* ws; int32_t* counters;
};

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcArgs a) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = a.BN;
  const int a_bytes = TC_BM * TC_BK * 2;                 // 16 KB
  const int b_bytes = BN * TC_BK * 2;
  const int stage_bytes = a_bytes + ((b_bytes + 1023) / 1024) * 1024;
  // 1024-byte aligned base (swizzle atoms)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_tc) + 1023) & ~(uintptr_t)1023);
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int m0 = m_tile * BN;
  // weight rows of this tile: two 64-row halves (SILU: value rows | gate rows)
  int row_lo, row_hi;
  if (a.epi == TEPI_SILU) { row_lo = n_tile * 64; row_hi = a.F + n_tile * 64; }
  else { row_lo = n_tile * TC_BM; row_hi = row_lo + 64; }
  const int nk_total = a.K / TC_BK;
  const int nk_per = (nk_total + a.ksplit - 1) / a.ksplit;
  const int kb0 = blockIdx.z * nk_per;
  const int nk = max(0, min(nk_total, kb0 + nk_per) - kb0);          // k-blocks of this split

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  uint32_t ncols = 32;
  while ((int)ncols < BN) ncols <<= 1;
  if (warp == 5) {                                       // TMEM allocation by one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % a.stages;
        if (kb >= a.stages) mbar_wait(&empty_bar[s], ((kb / a.stages) - 1) & 1);
        unsigned char* sa = base + (size_t)s * stage_bytes;
        unsigned char* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + b_bytes));
        tma_load_2d(sa, &a.map_w, (kb0 + kb) * TC_BK, row_lo, &full_bar[s]);
        tma_load_2d(sa + a_bytes / 2, &a.map_w, (kb0 + kb) * TC_BK, row_hi, &full_bar[s]);
        tma_load_2d(sb, &a.map_x, (kb0 + kb) * TC_BK, m0, &full_bar[s]);
      }
    }
  } else if (warp_id_uniform() == 5) {
    // ===== MMA issuer: the whole warp walks the ring on uniform values, one elected lane issues (see elect_one) =====
    const uint32_t idesc = make_idesc(TC_BM, BN);
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % a.stages;
      mbar_wait(&full_bar[s], (kb / a.stages) & 1);
      tc_fence_after();
      const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
      const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + a_bytes);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)                    // UMMA K = 16 bf16 = 32 bytes: advance the start address
          tc_mma(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb | kk) ? 1u : 0u);
        tc_commit(&empty_bar[s]);                        // frees the smem slot when these MMAs retire
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&tmem_full_bar);          // accumulator complete
    __syncwarp();
  } else {
    // ===== epilogue warps 0..3: TMEM lanes [32*warp, 32*warp+32) = weight rows of the tile =====
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int lr = warp * 32 + lane;                                   // row inside the 128-row tile
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    float* ws_tile = a.ws + (size_t)tile_id * a.ksplit * TC_BM * BN;   // [ksplit][128][BN]
    if (a.ksplit > 1) {
      __shared__ int s_last_tc;
      float* mine = ws_tile + ((size_t)blockIdx.z * TC_BM + lr) * BN;
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(mine + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) {
        int prev;
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(prev) : "l"(a.counters + tile_id) : "memory");
        s_last_tc = (prev == a.ksplit - 1);
        if (s_last_tc) a.counters[tile_id] = 0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (!s_last_tc) { tc_fence_before(); goto tc_done; }
    }
    // accumulator chunk: straight from TMEM, or (split-K) the sum of all partials in split order
    auto load_chunk = [&](int c0, float (&v)[16]) {
      if (a.ksplit == 1) { tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v); return; }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
      for (int sp = 0; sp < a.ksplit; ++sp) {
        const float* src = ws_tile + ((size_t)sp * TC_BM + lr) * BN + c0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 t4 = __ldcg(reinterpret_cast<const float4*>(src + j));
          v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
        }
      }
    };
    const int n = (lr < 64) ? row_lo + lr : row_hi + (lr - 64);        // weight row == output feature
    const bool n_ok = n < a.N;
    float* xchg = reinterpret_cast<float*>(base);                      // SILU: gate values [64][BN] (ring is idle now)
    if (a.epi == TEPI_SILU) {
      // pass 1: gate rows (warps 2,3) park their bf16-rounded values in shared memory
      if (warp >= 2) {
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float v[16];
          load_chunk(c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) xchg[(size_t)(lr - 64) * BN + c0 + j] = rbf(v[j]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp < 2) {
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float v[16];
          load_chunk(c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < a.M && n_ok) {
              const float g = xchg[(size_t)lr * BN + c0 + j];
              const float sg = rbf(g / (1.0f + expf(-g)));
              a.y[(size_t)m * a.ldy + n] = f2bf(__fmul_rn(rbf(v[j]), sg));
            }
          }
        }
      }
    } else if (a.epi == TEPI_QKV && !a.rope_interleaved) {
      // rotate-half RoPE (flash_attn's non-interleaved convention, hybrid variant): feature i of a head pairs with i + 64.  A
      // 128-row tile is exactly one head, so the partner sits 64 lanes away: all rows park their bf16-rounded values in
      // shared memory (the ring is idle now), then every row reads its partner's.
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        load_chunk(c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) xchg[(size_t)lr * BN + c0 + j] = rbf(v[j]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int qn = a.Hq * a.hd, kn = a.Hkv * a.hd;
      const bool rot = n < qn + kn;
      const int fi = (n % a.hd) % (a.hd / 2);                           // frequency index of this row and of its partner
      for (int c = 0; c < BN; ++c) {
        const int m = m0 + c;
        if (m >= a.M || !n_ok) continue;
        const int pos = a.lengths[m / a.T] + m % a.T;
        float o = xchg[(size_t)lr * BN + c];
        if (rot) {
          const float other = xchg[(size_t)(lr ^ 64) * BN + c];
          const float2 cs = *reinterpret_cast<const float2*>(a.rope + ((size_t)min(pos, a.rope_len - 1) * (a.hd / 2) + fi) * 2);
          o = (lr & 64) ? __fadd_rn(__fmul_rn(o, cs.x), __fmul_rn(other, cs.y))      // x[i+64]*c + x[i]*s
                        : __fsub_rn(__fmul_rn(o, cs.x), __fmul_rn(other, cs.y));     // x[i]*c - x[i+64]*s
        }
        if (n < qn) {
          a.q_out[(size_t)m * qn + n] = f2bf(o);
        } else {
          const int kvsel = n < qn + kn ? 0 : 1;
          const int ci = n - qn - kvsel * kn;
          const int page = a.page_table[(size_t)(m / a.T) * a.max_pages + pos / ZB_PAGE_TOKENS];
          bf16* pb = a.kv_layer + ((size_t)page * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;
          pb[((size_t)(ci / a.hd) * ZB_PAGE_TOKENS + pos % ZB_PAGE_TOKENS) * a.hd + (ci % a.hd)] = f2bf(o);
        }
      }
    } else {
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        load_chunk(c0, v);
        if (a.epi == TEPI_HEADS && a.cfg_scale != 1.0f) {
          // columns [0,B) are cond rows, [B,2B) uncond rows (single m tile): park the bf16-rounded row in shared memory
          // (the ring is idle now) and mix in a second pass, so no unaligned TMEM column reads are needed
#pragma unroll
          for (int j = 0; j < 16; ++j) xchg[(size_t)lr * BN + c0 + j] = rbf(v[j]);
          continue;
        }
        if (a.epi == TEPI_QKV) {
          // RoPE pairs (2i, 2i+1) sit in adjacent lanes (interleaved convention, _torch.py:57-68).  Positions, pages and
          // cos/sin of the 16 columns are fetched up front (independent loads) instead of one dependent chain per element.
          const int qn = a.Hq * a.hd, kn = a.Hkv * a.hd;
          const bool rot = n < qn + kn;
          int pos[16];
          float2 cs[16];
          int page[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = min(m0 + c0 + j, a.M - 1);
            pos[j] = a.lengths[m / a.T] + m % a.T;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = min(m0 + c0 + j, a.M - 1);
            cs[j] = rot && n_ok ? *reinterpret_cast<const float2*>(a.rope + ((size_t)min(pos[j], a.rope_len - 1) * (a.hd / 2) + (n % a.hd) / 2) * 2)
                                : make_float2(1.f, 0.f);
            page[j] = (n >= qn && n_ok) ? a.page_table[(size_t)(m / a.T) * a.max_pages + pos[j] / ZB_PAGE_TOKENS] : 0;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            float o = rbf(v[j]);
            const float other = __shfl_xor_sync(0xffffffffu, o, 1);
            if (m < a.M && n_ok) {
              if (rot) o = (n & 1) ? __fadd_rn(__fmul_rn(o, cs[j].x), __fmul_rn(other, cs[j].y))      // x1*c + x0*s
                                   : __fsub_rn(__fmul_rn(o, cs[j].x), __fmul_rn(other, cs[j].y));     // x0*c - x1*s
              if (n < qn) {
                a.q_out[(size_t)m * qn + n] = f2bf(o);
              } else {
                const int kvsel = n < qn + kn ? 0 : 1;
                const int ci = n - qn - kvsel * kn;
                bf16* pb = a.kv_layer + ((size_t)page[j] * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;
                pb[((size_t)(ci / a.hd) * ZB_PAGE_TOKENS + pos[j] % ZB_PAGE_TOKENS) * a.hd + (ci % a.hd)] = f2bf(o);
              }
            }
          }
          continue;
        }
        if (a.epi == TEPI_RESID) {
          float rv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {                     // 16 independent loads in flight, then 16 stores
            const int m = m0 + c0 + j;
            rv[j] = (m < a.M && n_ok) ? __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(a.resid + (size_t)m * a.ldr + n)) << 16) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < a.M && n_ok) a.y[(size_t)m * a.ldy + n] = f2bf(rv[j] + rbf(v[j]));
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = m0 + c0 + j;
          if (m < a.M && n_ok) {
            if (a.epi == TEPI_STORE) a.y[(size_t)m * a.ldy + n] = f2bf(v[j]);
            else if (a.epi == TEPI_HEADS) a.logits[(size_t)m * a.QV + n] = rbf(v[j]);
          }
        }
      }
    }
    if (a.epi == TEPI_HEADS && a.cfg_scale != 1.0f && n_ok) {
      for (int b = 0; b < a.B; ++b) {                      // u + (c - u) * s in fp32 (model.py:230-232)
        const float cc = xchg[(size_t)lr * BN + b], uu = xchg[(size_t)lr * BN + a.B + b];
        a.logits[(size_t)b * a.QV + n] = __fadd_rn(uu, __fmul_rn(__fsub_rn(cc, uu), a.cfg_scale));
      }
    }
    tc_fence_before();
  }
tc_done:
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// bf16 row-major [rows, cols] (cols contiguous), box {64 cols, box_rows}, 128-byte swizzle, zero fill out of bounds
zb_status make_map_2d(zb_ctx* ctx, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  ZB_REQUIRE(ctx, enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ZB_REQUIRE(ctx, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] ld %llu box %u", (int)r, (unsigned long long)rows,
             (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
  return ZB_OK;
}

}  // namespace

// Y = X W^T with a fused epilogue.  X rows must be 16-byte aligned and K a multiple of 64.
zb_status zb_launch_gemm_tc(zb_ctx* ctx, const zb_gemm_tc& g, cudaStream_t stream) {
  ZB_REQUIRE(ctx, g.K % TC_BK == 0 && g.K >= TC_BK, "gemm_tc: K=%d must be a multiple of 64", g.K);
  ZB_REQUIRE(ctx, g.M >= 1 && g.N >= 1, "gemm_tc: empty problem");
  TcArgs a;
  memset(&a, 0, sizeof(a));
  // activation rows per tile: a multiple of 16, at most 256, balanced over the tiles
  int mtiles = (g.M + 255) / 256;
  int BN = ((g.M + mtiles - 1) / mtiles + 15) / 16 * 16;
  if (g.epi == TEPI_HEADS && g.cfg_scale != 1.0f) {
    ZB_REQUIRE(ctx, g.M <= 256, "gemm_tc: CFG heads need all rows in one tile (M=%d)", g.M);
    mtiles = 1;
    BN = (g.M + 15) / 16 * 16;
  }
  if (BN < 16) BN = 16;
  a.M = g.M; a.N = g.N; a.K = g.K; a.BN = BN; a.epi = g.epi; a.F = g.F;
  a.y = g.y; a.ldy = g.ldy; a.resid = g.resid; a.ldr = g.ldr;
  a.T = g.T; a.Hq = g.Hq; a.Hkv = g.Hkv; a.hd = g.hd; a.rope_interleaved = g.rope_interleaved; a.rope_len = g.rope_len;
  a.max_pages = g.max_pages; a.rope = g.rope; a.lengths = g.lengths; a.page_table = g.page_table; a.kv_layer = g.kv_layer; a.q_out = g.q_out;
  a.B = g.B; a.cfg_scale = g.cfg_scale; a.logits = g.logits; a.QV = g.QV;
  ZB_REQUIRE(ctx, g.epi != TEPI_QKV || g.rope_interleaved || g.hd == TC_BM, "gemm_tc: the rotate-half RoPE epilogue needs one head per 128-row tile");
  if (zb_status st = make_map_2d(ctx, &a.map_w, g.W, g.N, g.K, g.K, 64)) return st;
  if (zb_status st = make_map_2d(ctx, &a.map_x, g.x, g.M, g.K, g.ldx, BN)) return st;
  const int a_bytes = TC_BM * TC_BK * 2, b_bytes = ((BN * TC_BK * 2 + 1023) / 1024) * 1024;
  int stages = (200 * 1024) / (a_bytes + b_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  // epilogue exchange buffer (reuses the ring): gate rows for SiLU, whole tiles for the CFG mix and the rotate-half RoPE
  const size_t xchg = g.epi == TEPI_SILU ? (size_t)64 * BN * 4
                    : (g.epi == TEPI_HEADS && g.cfg_scale != 1.0f) || (g.epi == TEPI_QKV && !g.rope_interleaved) ? (size_t)128 * BN * 4 : 0;
  // Many m-tiles (prefill): TWO CTAs per SM, so that the epilogue of one tile (RoPE + KV append, SiLU: as long as its main
  // loop) overlaps the MMAs of another - 2 x <= 113 KB of shared memory, 2 x <= 256 TMEM columns
  static const int two_cta = getenv("ZB_TC_TWO_CTA") ? atoi(getenv("ZB_TC_TWO_CTA")) : 1;
  if (two_cta && mtiles >= 4 && xchg + 1024 <= 112 * 1024) {
    const int s2 = (int)((112 * 1024 - 1024) / (a_bytes + b_bytes));
    if (s2 >= 2) stages = std::min(stages, s2);
  }
  a.stages = stages;
  size_t smem = (size_t)stages * (a_bytes + b_bytes) + 1024;
  if (xchg + 1024 > smem) smem = xchg + 1024;
  ZB_CUDA(ctx, zb_ensure_smem(ctx, gemm_tc_kernel, smem));
  const int ntiles = (g.epi == TEPI_SILU) ? (g.F + 63) / 64 : (g.N + TC_BM - 1) / TC_BM;
  // split-K for the bandwidth-bound small-M regime: enough CTAs to pull weights with every SM
  int ksplit = 1;
  // ZB_TC_SPLITK: 0 never, 1 whenever the grid is smaller than the GPU, default -1 = only for few activation rows and
  // few weight tiles (batch 3..8 decode: the N=2048 matrices are 16 tiles, i.e. 16 SMs would stream them alone; at
  // batch 64 split-K measured slower: 1.75 s vs 1.52 s per pass)
  static const int splitk_env = getenv("ZB_TC_SPLITK") ? atoi(getenv("ZB_TC_SPLITK")) : -1;
  static const int splitk_maxm = getenv("ZB_TC_SPLITK_MAXM") ? atoi(getenv("ZB_TC_SPLITK_MAXM")) : 32;
  static const int splitk_maxtiles = getenv("ZB_TC_SPLITK_MAXTILES") ? atoi(getenv("ZB_TC_SPLITK_MAXTILES")) : 48;
  const bool splitk_on = splitk_env > 0 || (splitk_env < 0 && g.decode && g.M <= splitk_maxm && ntiles <= splitk_maxtiles);
  if (splitk_on && mtiles == 1 && ntiles < ctx->num_sms) {
    ksplit = (ctx->num_sms + ntiles - 1) / ntiles;
    const int nkb = g.K / TC_BK;
    if (ksplit > nkb / 4) ksplit = nkb / 4;              // at least 4 k-blocks (256 k) per split
    if (ksplit > 16) ksplit = 16;
    if (ksplit < 1) ksplit = 1;
  }
  a.ksplit = ksplit;
  if (ksplit > 1) {
    const size_t ws_bytes = (size_t)ntiles * ksplit * TC_BM * BN * sizeof(float);
    if (zb_status st = zb_tc_workspace_reserve(ctx, ws_bytes)) return st;
    ZB_REQUIRE(ctx, (size_t)ntiles + 65536 <= ZB_NUM_COUNTERS, "gemm_tc: too many tiles for the split-K counters");
    a.ws = (float*)ctx->tc_ws;
    a.counters = ctx->counters + 65536 * 8;              // disjoint from the attention-merge counters
  }
  ZB_CUDA(ctx, zb_launch_pdl(gemm_tc_kernel, dim3(ntiles, mtiles, ksplit), dim3(TC_THREADS), smem, stream, a));
  ctx->launches++;
  return ZB_OK;
}

// diagnostics: plain Y = X W^T (bf16) for tests
extern "C" ZB_API zb_status zb_debug_gemm(zb_ctx* ctx, const void* x, const void* w, void* y, int32_t M, int32_t N, int32_t K, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  zb_gemm_tc g;
  g.W = (const bf16*)w; g.x = (const bf16*)x; g.ldx = K; g.M = M; g.N = N; g.K = K; g.epi = TEPI_STORE; g.y = (bf16*)y; g.ldy = N;
  return zb_launch_gemm_tc(ctx, g, (cudaStream_t)stream);
}
