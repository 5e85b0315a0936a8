// Backbone kernels for few activation rows (decode, and prefill as row tiles): HBM-bound weight and KV
// streaming with 16-byte coalesced loads, fused prologues (LayerNorm/RMSNorm) and epilogues
// (RoPE + paged KV append, residual add, SiLU gate, heads + CFG mix), warp-shuffle reductions.
//
// Reference call sites replaced (zonos/backbone/_torch.py): :326-328 block wiring, :401 in_proj split,
// :57-68 RoPE, :105-106 KV write, :415 SDPA (GQA, causal), :419-420 out_proj twice, :473-474 gated MLP,
// :238 final norm; zonos/utilities/codec_utils.py:37 embedding sum, :68-79 heads; zonos/model.py:229-233 CFG.
// Rounding points are the reference's: every Linear / norm / SiLU / residual add rounds to bf16.
#include <stdlib.h>

#include <algorithm>

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "tc.cuh"

namespace {

constexpr int kWarps = 8;               // warps per GEMV CTA
constexpr int kThreads = kWarps * 32;
constexpr int kMaxMT = 8;               // activation rows per CTA tile

__device__ __forceinline__ bool loop_idle(const zb_loop_state* st, int T_delayed) {
  // device-driven generate: nothing to do once finished, or when this launch is the closing one (model.py:471-472)
  return st && (st->done || st->offset + 1 >= T_delayed);
}

// ------------------------------------------------------------------ embedding sum ------------
struct EmbedArgs {
  const bf16* tab[16];
  const int64_t* codes; int64_t sb, sq, st;
  int B, T, Q, D, vocab, repeat;
  bf16* out; int64_t out_rs;          // elements between consecutive output rows-of-T
  const zb_loop_state* loop; int T_delayed;
};

__global__ void __launch_bounds__(256) embed_kernel(EmbedArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (loop_idle(a.loop, a.T_delayed)) return;
  const int b = blockIdx.x / a.T, t = blockIdx.x % a.T;
  const int64_t col = a.loop ? (int64_t)a.loop->offset : (int64_t)t;
  long long ids[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    long long id = (k < a.Q) ? a.codes[b * a.sb + k * a.sq + col * a.st] : 0;     // all id loads in flight together
    ids[k] = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
  }
  for (int d0 = threadIdx.x * 8; d0 < a.D; d0 += blockDim.x * 8) {
    uint4 rows[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < a.Q) rows[k] = *reinterpret_cast<const uint4*>(a.tab[k] + (size_t)ids[k] * a.D + d0);   // then all row loads
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < a.Q) {
        const uint32_t w[4] = {rows[k].x, rows[k].y, rows[k].z, rows[k].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {      // sequential bf16 adds: ((0 + E0) + E1) + ...  (codec_utils.py:37)
          acc[2 * i] = rbf(acc[2 * i] + bf16lo(w[i]));
          acc[2 * i + 1] = rbf(acc[2 * i + 1] + bf16hi(w[i]));
        }
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
    for (int rep = 0; rep < a.repeat; ++rep)
      *reinterpret_cast<uint4*>(a.out + (size_t)(rep * a.B + b) * a.out_rs + (size_t)t * a.D + d0) = o;
  }
}

// ------------------------------------------------------------------ skinny GEMM (GEMV family) -
enum { PRO_NONE = 0, PRO_NORM = 1, PRO_GATED = 2 };           // GATED: mamba_ssm RMSNormGated of y * silu(z) (persistent kernel only)
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_QKV = 2, EPI_SILU = 3, EPI_HEADS = 4 };

struct GemvArgs {
  const bf16* W;            // [N, K]
  const bf16* x; int64_t ldx;   // activation rows; row m at x + m*ldx (HEADS: see row map)
  int M, N, K;
  // prologue
  const bf16* nw; const bf16* nb; float eps; int norm_kind;
  // EPI_STORE / EPI_RESID
  bf16* y; int64_t ldy; const bf16* resid; int64_t ldr;
  // EPI_QKV
  int T;                    // tokens per cache row (m = r*T + t)
  int Hq, Hkv, hd, rope_interleaved;
  const float* rope; int rope_len;
  const int32_t* lengths; const int32_t* page_table; int max_pages; int num_pages;
  bf16* kv_layer;           // this layer's pages [num_pages][2][Hkv][64][hd]
  bf16* q_out;              // [M, Hq*hd]
  // EPI_SILU
  int F;
  // EPI_HEADS
  int B; float cfg_scale; float* logits; int QV;
  const zb_loop_state* loop; int T_delayed;
  int ring_stages, prefetch_ahead;      // gemv3: stages in the shared-memory ring, stages prefetched into L2 beyond it
  unsigned long long* timeline;         // debug: 8 globaltimer stamps per launch (CTA 0), or null
  const float* wscale;                  // persistent kernel, FP8 mode: one power-of-two scale per weight row (W then holds e4m3 bytes)
};


__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

// Finishes one output pair (n0, n1) of activation row m from the fp32 dot products v0, v1.  Rounding points are the
// reference's (bf16 Linear output first, then the fused op).  HEADS with CFG: m is the utterance, (v0,v1) its cond
// row and (u0,u1) its uncond row.
template <int EPI>
__device__ __forceinline__ void gemv_epilogue(const GemvArgs& a, int m, int n0, int n1, bool has1, float v0, float v1,
                                              float u0, float u1) {
  if (EPI == EPI_STORE) {
    a.y[(size_t)m * a.ldy + n0] = f2bf(v0);
    if (has1) a.y[(size_t)m * a.ldy + n1] = f2bf(v1);
  } else if (EPI == EPI_RESID) {
    a.y[(size_t)m * a.ldy + n0] = f2bf(ldcg_bf16(a.resid + (size_t)m * a.ldr + n0) + rbf(v0));
    if (has1) a.y[(size_t)m * a.ldy + n1] = f2bf(ldcg_bf16(a.resid + (size_t)m * a.ldr + n1) + rbf(v1));
  } else if (EPI == EPI_SILU) {
    const float yv = rbf(v0), g = rbf(v1);
    const float sg = rbf(g / (1.0f + expf(-g)));            // F.silu on bf16: fp32 math, bf16 result
    a.y[(size_t)m * a.ldy + n0] = f2bf(__fmul_rn(yv, sg));
  } else if (EPI == EPI_QKV) {
    const int r = m / a.T, t = m % a.T;
    const int pos = a.lengths[r] + t;
    const int qn = a.Hq * a.hd, kn = a.Hkv * a.hd;
    float o0 = rbf(v0), o1 = rbf(v1);
    if (n0 < qn + kn) {                                      // q or k: rotate (_torch.py:57-68)
      const int i = a.rope_interleaved ? (n0 % a.hd) / 2 : (n0 % a.hd);
      const float2 cs = *reinterpret_cast<const float2*>(a.rope + ((size_t)min(pos, a.rope_len - 1) * (a.hd / 2) + i) * 2);
      // separate fp32 mul / sub / add like the reference's eager ops (no FMA contraction)
      const float r0 = __fsub_rn(__fmul_rn(o0, cs.x), __fmul_rn(o1, cs.y));
      const float r1 = __fadd_rn(__fmul_rn(o1, cs.x), __fmul_rn(o0, cs.y));
      o0 = r0; o1 = r1;
    }
    if (n0 < qn) {
      a.q_out[(size_t)m * qn + n0] = f2bf(o0);
      a.q_out[(size_t)m * qn + n1] = f2bf(o1);
    } else {
      const int kvsel = n0 < qn + kn ? 0 : 1;
      const int c0i = n0 - qn - kvsel * kn, c1i = n1 - qn - kvsel * kn;
      const int page = a.page_table[(size_t)r * a.max_pages + pos / ZB_PAGE_TOKENS];
      bf16* base = a.kv_layer + ((size_t)page * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;
      const int tk = pos % ZB_PAGE_TOKENS;
      base[((size_t)(c0i / a.hd) * ZB_PAGE_TOKENS + tk) * a.hd + (c0i % a.hd)] = f2bf(o0);
      base[((size_t)(c1i / a.hd) * ZB_PAGE_TOKENS + tk) * a.hd + (c1i % a.hd)] = f2bf(o1);
    }
  } else if (EPI == EPI_HEADS) {
    if (a.cfg_scale != 1.0f) {                               // u + (c - u) * s in fp32 (model.py:230-232)
      const float c0 = rbf(v0), c1 = rbf(v1), w0 = rbf(u0), w1 = rbf(u1);
      a.logits[(size_t)m * a.QV + n0] = __fadd_rn(w0, __fmul_rn(__fsub_rn(c0, w0), a.cfg_scale));
      if (has1) a.logits[(size_t)m * a.QV + n1] = __fadd_rn(w1, __fmul_rn(__fsub_rn(c1, w1), a.cfg_scale));
    } else {
      a.logits[(size_t)m * a.QV + n0] = rbf(v0);
      if (has1) a.logits[(size_t)m * a.QV + n1] = rbf(v1);
    }
  }
}

// Weight-row pair of unit p (see gemv kernels): SiLU pairs value row with gate row, rotate-half RoPE pairs (i, i+hd/2).
template <int EPI>
__device__ __forceinline__ void unit_rows(const GemvArgs& a, int p, int& n0, int& n1) {
  if (EPI == EPI_SILU) { n0 = p; n1 = p + a.F; }
  else if (EPI == EPI_QKV && !a.rope_interleaved && p < (a.Hq + a.Hkv) * (a.hd / 2)) {
    const int half = a.hd / 2;
    n0 = (p / half) * a.hd + (p % half); n1 = n0 + half;
  } else { n0 = 2 * p; n1 = 2 * p + 1; }
}

template <int MT, int PRO, int EPI>
__global__ void __launch_bounds__(kThreads) gemv_kernel(GemvArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (loop_idle(a.loop, a.T_delayed)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);          // [MT][K]
  __shared__ float red[kWarps][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K;
  const int m0 = blockIdx.y * MT;

  // row map: tile row i -> activation row m (HEADS with CFG pairs cond row b with uncond row B+b)
  auto row_of = [&](int i) -> int {
    if (EPI == EPI_HEADS && a.cfg_scale != 1.0f) {
      const int half = MT / 2;
      const int b = blockIdx.y * half + (i % half);
      return (b < a.B) ? b + (i / half) * a.B : -1;
    }
    const int m = m0 + i;
    return m < a.M ? m : -1;
  };

  // ---- stage activations (optionally normalised) into shared memory ----
#pragma unroll 1
  for (int i = 0; i < MT; ++i) {
    const int m = row_of(i);
    bf16* dst = xs + (size_t)i * K;
    if (m < 0) {
      for (int k = threadIdx.x * 8; k < K; k += kThreads * 8) *reinterpret_cast<uint4*>(dst + k) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const bf16* src = a.x + (size_t)m * a.ldx;
    if (PRO == PRO_NONE) {
      for (int k = threadIdx.x * 8; k < K; k += kThreads * 8)
        *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(src + k);
    } else {
      // two-pass fp32 statistics over the bf16 row (nn.LayerNorm / RMSNorm semantics), output rounded to bf16
      float s = 0.f;
      for (int k = threadIdx.x * 8; k < K; k += kThreads * 8) {
        uint4 v = *reinterpret_cast<const uint4*>(src + k);
        s += bf16lo(v.x) + bf16hi(v.x) + bf16lo(v.y) + bf16hi(v.y) + bf16lo(v.z) + bf16hi(v.z) + bf16lo(v.w) + bf16hi(v.w);
      }
      s = warp_sum(s);
      if (lane == 0) red[warp][0] = s;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += red[w][0];
      const float mean = (a.norm_kind == ZB_NORM_LAYERNORM) ? tot / (float)K : 0.f;
      float sq = 0.f;
      for (int k = threadIdx.x * 8; k < K; k += kThreads * 8) {
        uint4 v = *reinterpret_cast<const uint4*>(src + k);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d0 = bf16lo(w4[j]) - mean, d1 = bf16hi(w4[j]) - mean;
          sq += d0 * d0 + d1 * d1;
        }
      }
      sq = warp_sum(sq);
      if (lane == 0) red[warp][1] = sq;
      __syncthreads();
      float tsq = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tsq += red[w][1];
      const float rstd = rsqrtf(tsq / (float)K + a.eps);
      for (int k = threadIdx.x * 8; k < K; k += kThreads * 8) {
        uint4 v = *reinterpret_cast<const uint4*>(src + k);
        uint4 g = *reinterpret_cast<const uint4*>(a.nw + k);
        uint4 bb = a.nb ? *reinterpret_cast<const uint4*>(a.nb + k) : make_uint4(0, 0, 0, 0);
        const uint32_t xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g.x, g.y, g.z, g.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo = (bf16lo(xv[j]) - mean) * rstd * bf16lo(gv[j]) + bf16lo(bv[j]);
          float hi = (bf16hi(xv[j]) - mean) * rstd * bf16hi(gv[j]) + bf16hi(bv[j]);
          o[j] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<uint4*>(dst + k) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- weight streaming: each warp owns pairs of weight rows (n0, n1) ----
  const int npairs = (EPI == EPI_SILU) ? a.F : (a.N + 1) / 2;
  const int gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
  for (int p = gw; p < npairs; p += nw) {
    int n0, n1;
    unit_rows<EPI>(a, p, n0, n1);
    const bool has1 = n1 < a.N;
    const bf16* w0 = a.W + (size_t)n0 * K;
    const bf16* w1 = a.W + (size_t)(has1 ? n1 : n0) * K;
    float acc0[MT], acc1[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
    constexpr int U = 4;                     // 16-byte chunks in flight per weight row per lane
    for (int kb = 0; kb < K; kb += 256 * U) {
      uint4 a0[U], a1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * 256 + lane * 8;
        if (k < K) { a0[u] = ldg_stream(w0 + k); a1[u] = ldg_stream(w1 + k); }
        else { a0[u] = make_uint4(0, 0, 0, 0); a1[u] = make_uint4(0, 0, 0, 0); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * 256 + lane * 8;
        if (k < K) {
          const uint32_t c0[4] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w};
          const uint32_t c1[4] = {a1[u].x, a1[u].y, a1[u].z, a1[u].w};
#pragma unroll
          for (int i = 0; i < MT; ++i) {
            uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)i * K + k);
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float xl = bf16lo(xw[j]), xh = bf16hi(xw[j]);
              acc0[i] = fmaf(bf16lo(c0[j]), xl, acc0[i]);
              acc0[i] = fmaf(bf16hi(c0[j]), xh, acc0[i]);
              acc1[i] = fmaf(bf16lo(c1[j]), xl, acc1[i]);
              acc1[i] = fmaf(bf16hi(c1[j]), xh, acc1[i]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) { acc0[i] = warp_sum(acc0[i]); acc1[i] = warp_sum(acc1[i]); }

    // ---- epilogue: lane i finishes tile row i ----
    if (lane < MT) {
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int i = 0; i < MT; ++i) if (i == lane) { v0 = acc0[i]; v1 = acc1[i]; }
      const int m = row_of(lane);
      if (EPI == EPI_HEADS && a.cfg_scale != 1.0f) {
        // lanes [0,half) hold cond rows, lanes [half, MT) the matching uncond rows (model.py:230-232)
        constexpr int half = MT / 2;
        const float u0 = __shfl_down_sync((1u << MT) - 1, v0, half, 32);
        const float u1 = __shfl_down_sync((1u << MT) - 1, v1, half, 32);
        if (lane < half && m >= 0) gemv_epilogue<EPI>(a, m, n0, n1, has1, v0, v1, u0, u1);
      } else if (m >= 0) {
        gemv_epilogue<EPI>(a, m, n0, n1, has1, v0, v1, 0.f, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------ decode GEMV (M <= 4 rows) ---
// One CTA per SM-sized slice of the weight matrix; the slice streams through a shared-memory ring filled by the
// TMA engine (cp.async.bulk + mbarrier, producer warp) and drained by 16 consumer warps.  Consumers keep their
// k-slice of all R activation rows as fp32 pairs in REGISTERS; the inner loop is LDS.128 + bf16->fp32 shifts +
// packed FFMA2 only, every stage is full (no per-row branches) and V = RW*R lane-partials are reduced with one
// butterfly per stage.
// PDL: weights do not depend on the previous kernel, so the producer fills the whole ring BEFORE
// griddepcontrol.wait - the HBM stream of kernel N+1 overlaps the tail and launch latency of kernel N.
// Weights are read once per step: their loads carry an L2 evict_first policy so KV and activations stay in L2.
constexpr int kW3 = 16;                      // consumer warps
constexpr int kMaxStages = 24;               // ring stages are a launch parameter (a.ring_stages); stage = RG*RW weight rows

template <int EPI>
__device__ __forceinline__ int units_total(const GemvArgs& a) {
  return (EPI == EPI_SILU) ? a.F : (EPI == EPI_QKV) ? a.N / 2 : a.N;     // pairs only where the epilogue couples rows
}
// local row index -> weight row, for the CTA whose units are [u_begin, ...)
template <int EPI>
__device__ __forceinline__ int row_of_local(const GemvArgs& a, int u_begin, int lr) {
  if (EPI == EPI_SILU || EPI == EPI_QKV) {
    int n0, n1;
    unit_rows<EPI>(a, u_begin + (lr >> 1), n0, n1);
    return (lr & 1) ? n1 : n0;
  }
  return u_begin + lr;
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define ZB_STAMP(i) do { if (a.timeline && blockIdx.x == 0) a.timeline[i] = gtime(); } while (0)

template <int R, int NC, int RW, int PRO, int EPI>
__global__ void __launch_bounds__((kW3 + 1) * 32, (R <= 2) ? 2 : 1) gemv3_kernel(GemvArgs a) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) ZB_STAMP(0);                        // kernel start
  extern __shared__ __align__(128) unsigned char smem3[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ float red[2][kW3][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K;
  const int row_bytes = K * 2;
  constexpr int Kc = NC * 256;
  const int KS = K / Kc;                                       // k-slices; RG = kW3 / KS row groups
  const int RPS = (kW3 / KS) * RW;                             // weight rows per stage
  const int kStageBytes = RPS * row_bytes;
  const int kStages = a.ring_stages;
  const int kPrefetchAhead = a.prefetch_ahead;
  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);
  const int nunits = units_total<EPI>(a);
  const int u_begin = (int)((long long)blockIdx.x * nunits / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * nunits / gridDim.x);
  const int nrows = (u_end - u_begin) * (kPairs ? 2 : 1);
  const int nstage = (nrows + RPS - 1) / RPS;
  unsigned char* ring = smem3;
  float* part = reinterpret_cast<float*>(smem3 + (size_t)kStages * kStageBytes);   // [nstage*RPS][KS][R]

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kW3); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kW3) {
    // ===== producer warp drives the TMA engine; every stage is filled completely (tail rows repeat the last row,
    // their results are never read) so the consumers run branch-free.  Lane q issues the copy of row q of a stage
    // (the issue cost of a stage's copies is paid once per warp instruction); lane 0 arms the barrier first. =====
    const uint64_t pol = l2_evict_first_policy();
    for (int st = 0; st < nstage; ++st) {
      const int slot = st % kStages;
      if (st >= kStages) mbar_wait(&empty_bar[slot], ((st / kStages) - 1) & 1);
      const int r0 = st * RPS;
      if (lane == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)kStageBytes);
      if (lane == 0 && st == 0) ZB_STAMP(1);                   // first copy issued
      if (lane == 0 && st == nstage - 1) ZB_STAMP(2);          // last copy issued
      __syncwarp();
      unsigned char* dst = ring + (size_t)slot * kStageBytes;
      if (!kPairs && r0 + RPS <= nrows) {                     // contiguous rows: one copy per stage
        if (lane == 0) bulk_g2s(dst, a.W + (size_t)(u_begin + r0) * K, (uint32_t)kStageBytes, &full_bar[slot], pol);
      } else {
        for (int q = lane; q < RPS; q += 32) {
          const int lr = min(r0 + q, nrows - 1);
          bulk_g2s(dst + (size_t)q * row_bytes, a.W + (size_t)row_of_local<EPI>(a, u_begin, lr) * K, row_bytes, &full_bar[slot], pol);
        }
      }
      if (kPrefetchAhead > 0 && st >= kStages - 1) {           // optional rolling L2 prefetch beyond the ring
        const int first = (st == kStages - 1) ? kStages : st + kPrefetchAhead;
        const int last = (st == kStages - 1) ? kStages + kPrefetchAhead : st + kPrefetchAhead + 1;
        for (int ps = first; ps < last && ps < nstage; ++ps)
          for (int q = lane; q < RPS && ps * RPS + q < nrows; q += 32)
            prefetch_l2(a.W + (size_t)row_of_local<EPI>(a, u_begin, ps * RPS + q) * K, row_bytes);
      }
    }
    return;
  }

  // ===== consumers =====
  const int ks = warp % KS, rg = warp / KS;
  const size_t koff = (size_t)ks * Kc;
  uint4 nwr[NC], nbr[NC];                                      // norm weight / bias slice: parameters, loaded before the wait
  if (PRO == PRO_NORM) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const size_t k = koff + c * 256 + lane * 8;
      nwr[c] = *reinterpret_cast<const uint4*>(a.nw + k);
      nbr[c] = a.nb ? *reinterpret_cast<const uint4*>(a.nb + k) : make_uint4(0, 0, 0, 0);
    }
  }
  pdl_wait();                                                  // activations / loop state come from the predecessor
  if (threadIdx.x == 0) ZB_STAMP(3);                           // dependency satisfied
  const bool idle = loop_idle(a.loop, a.T_delayed);            // still drain the ring: the producer already filled it

  // this warp's k-slice of the R activation rows as fp32 pairs: lane holds k = koff + c*256 + lane*8 + [0,8)
  float xf[R][NC * 8];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const uint4 v = (i < a.M && !idle) ? *reinterpret_cast<const uint4*>(a.x + (size_t)i * a.ldx + koff + c * 256 + lane * 8)
                                         : make_uint4(0, 0, 0, 0);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { xf[i][c * 8 + 2 * j] = bf16lo(w4[j]); xf[i][c * 8 + 2 * j + 1] = bf16hi(w4[j]); }
    }
  if (PRO == PRO_NORM) {
    // fp32 statistics over the bf16 row (nn.LayerNorm / RMSNorm), result rounded to bf16 like the reference.  Sum and
    // sum of squares travel through ONE barrier (var = E[x^2] - mean^2; the rows are O(1) with |mean| << std, so the
    // cancellation error is far below the bf16 rounding that follows).
    float mean[R], rstd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float sacc = 0.f, qacc = 0.f;
#pragma unroll
      for (int e = 0; e < NC * 8; ++e) { sacc += xf[i][e]; qacc = fmaf(xf[i][e], xf[i][e], qacc); }
      sacc = warp_sum(sacc);
      qacc = warp_sum(qacc);
      if (rg == 0 && lane == 0) { red[0][ks][i] = sacc; red[1][ks][i] = qacc; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");      // consumer warps only
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float tot = 0.f, tsq = 0.f;
      for (int q = 0; q < KS; ++q) { tot += red[0][q][i]; tsq += red[1][q][i]; }
      const float mu = tot / (float)K;
      mean[i] = (a.norm_kind == ZB_NORM_LAYERNORM) ? mu : 0.f;
      const float var = (a.norm_kind == ZB_NORM_LAYERNORM) ? fmaxf(tsq / (float)K - mu * mu, 0.f) : tsq / (float)K;
      rstd[i] = rsqrtf(var + a.eps);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const uint32_t gv[4] = {nwr[c].x, nwr[c].y, nwr[c].z, nwr[c].w}, bv[4] = {nbr[c].x, nbr[c].y, nbr[c].z, nbr[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = bf16lo(gv[j]), g1 = bf16hi(gv[j]), b0 = bf16lo(bv[j]), b1 = bf16hi(bv[j]);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          xf[i][c * 8 + 2 * j] = rbf((xf[i][c * 8 + 2 * j] - mean[i]) * rstd[i] * g0 + b0);
          xf[i][c * 8 + 2 * j + 1] = rbf((xf[i][c * 8 + 2 * j + 1] - mean[i]) * rstd[i] * g1 + b1);
        }
      }
    }
  }
  unsigned long long x2[R][NC * 4];                            // (even k, odd k) pairs for FFMA2
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int e = 0; e < NC * 4; ++e) x2[i][e] = pack_f32x2(xf[i][2 * e], xf[i][2 * e + 1]);

  constexpr int V = RW * R;                                    // values each warp reduces per stage
  const int my_idx = multi_reduce_index<V>(lane);
  const bool writer = (lane & ((32 / V) - 1)) == 0;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  // lane's first weight byte inside a stage: row rg*RW, k = koff + lane*8
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(rg * RW) * row_bytes + (uint32_t)(koff + lane * 8) * 2;
  // where this lane stores its reduced value: part[(r0 + rg*RW + w)*KS + ks][i], w = my_idx / R, i = my_idx % R
  const uint32_t part_lane = smem_u32(part) + (uint32_t)((((rg * RW + my_idx / R) * KS + ks) * R + my_idx % R) * 4);
  const uint32_t part_stage = (uint32_t)(RPS * KS * R * 4);
  uint32_t slot = 0, phase = 0;
  if (threadIdx.x == 0) ZB_STAMP(4);                           // activations normalised, ready to consume
  for (int st = 0; st < nstage; ++st) {
    mbar_wait_u32(full0 + slot * 8, phase);
    if (threadIdx.x == 0 && st == 0) ZB_STAMP(5);              // first stage landed
    const uint32_t src = lane_base + slot * kStageBytes;
    unsigned long long acc2[V];
#pragma unroll
    for (int q = 0; q < V; ++q) acc2[q] = 0ull;
#pragma unroll
    for (int w = 0; w < RW; ++w) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const uint4 wv = lds128(src + w * row_bytes + c * 512);
        const uint32_t c0[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const unsigned long long w2 = pack_f32x2(bf16lo(c0[jj]), bf16hi(c0[jj]));
#pragma unroll
          for (int i = 0; i < R; ++i) ffma2(acc2[w * R + i], w2, x2[i][c * 4 + jj]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_u32(empty0 + slot * 8);           // weights are consumed: release the slot before reducing
    float acc[V];
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = sum_f32x2(acc2[q]);
    warp_reduce_multi<V>(acc);
    if (writer) sts32(part_lane + st * part_stage, acc[0]);
    if (++slot == (uint32_t)kStages) { slot = 0; phase ^= 1; }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
  if (threadIdx.x == 0) ZB_STAMP(6);                           // all stages consumed
  if (idle) return;

  // ---- combine the k-slices and finish: one thread per (unit, activation row) ----
  const bool cfg = (EPI == EPI_HEADS && a.cfg_scale != 1.0f);
  const int rows_out = cfg ? a.B : a.M;
  const int nu = u_end - u_begin;
  for (int t = threadIdx.x; t < nu * rows_out; t += kW3 * 32) {
    const int j = t / rows_out, i = t % rows_out;
    int n0, n1;
    float v0 = 0.f, v1 = 0.f, u0 = 0.f, u1 = 0.f;
    if (kPairs) {
      unit_rows<EPI>(a, u_begin + j, n0, n1);
      const float* s0 = part + (size_t)(2 * j) * KS * R;
      const float* s1 = s0 + (size_t)KS * R;
      for (int q = 0; q < KS; ++q) { v0 += s0[q * R + i]; v1 += s1[q * R + i]; }
      gemv_epilogue<EPI>(a, i, n0, n1, true, v0, v1, 0.f, 0.f);
    } else {
      n0 = u_begin + j; n1 = n0 + 1;
      const float* s0 = part + (size_t)j * KS * R;
      for (int q = 0; q < KS; ++q) { v0 += s0[q * R + i]; if (cfg) u0 += s0[q * R + a.B + i]; }
      gemv_epilogue<EPI>(a, i, n0, n1, false, v0, v1, u0, u1);
    }
  }
  if (threadIdx.x == 0) ZB_STAMP(7);                           // done
}

// ------------------------------------------------------------------ paged attention ----------
// grid (M, Hkv, nsplit); block = G warps (one per q head of the GQA group).  Each CTA handles the keys
// [split*CH, (split+1)*CH) of one (row, kv head); partials are merged by the last CTA to finish.
constexpr int kCH = ZB_PAGE_TOKENS;      // keys per split == page size, so a K (or V) tile is one contiguous 16 KB run
constexpr int kHD = 128;
constexpr int kKStride = kHD + 8;        // padded row stride (bf16) -> conflict-free 16 B column reads
constexpr int kPart = kHD + 4;           // floats per split partial: o[128], max, sum, pad (keeps float4 alignment)

struct AttnArgs {
  const bf16* q;            // [M, Hq*hd]
  const bf16* kv_layer;     // pages
  const int32_t* lengths; const int32_t* page_table; int max_pages;
  int T, Hq, Hkv, nsplit;
  float scale;
  float* part;              // [M, Hq, nsplit, hd + 2]
  int32_t* counters;        // [M, Hkv]
  bf16* y;                  // [M, Hq*hd]
  const zb_loop_state* loop; int T_delayed;
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(256) attn_kernel(AttnArgs a) {
  pdl_launch_dependents();
  __shared__ __align__(16) bf16 ks[kCH * kKStride];
  __shared__ __align__(16) bf16 vs[kCH * kHD];
  __shared__ __align__(16) float qs[8][kHD];
  __shared__ float ps[8][kCH];
  __shared__ int s_last;
  // Everything read before pdl_wait() was written by EARLIER steps / forwards (graph launches serialise): the loop
  // state, lengths, the page table and the K/V of tokens cached before this call.  Only q and the K/V of this call's
  // own tokens come from the producer kernel, so the bulk of the tile is already in flight when the wait returns.
  if (loop_idle(a.loop, a.T_delayed)) return;
  const int m = blockIdx.x, g = blockIdx.y, split = blockIdx.z;
  const int G = a.Hq / a.Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = m / a.T, t = m % a.T;
  const int old_len = a.lengths[r];
  const int kv_len = old_len + t + 1;                      // causal: keys 0..pos inclusive
  const int nact = (kv_len + kCH - 1) / kCH;               // splits that have keys
  if (split >= nact) return;
  const int k0 = split * kCH;
  const int nk = min(kCH, kv_len - k0);
  const int n_old = max(0, min(nk, old_len - k0));         // keys of this split that predate this call
  const int page = a.page_table[(size_t)r * a.max_pages + split];
  const bf16* kp = a.kv_layer + (((size_t)page * 2 + 0) * a.Hkv + g) * kCH * kHD;
  const bf16* vp = a.kv_layer + (((size_t)page * 2 + 1) * a.Hkv + g) * kCH * kHD;
  auto issue = [&](int lo, int hi) {
    for (int c = threadIdx.x; c < kCH * kHD / 8; c += blockDim.x) {
      const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
      if (tok >= lo && tok < hi) {
        cp_async16(ks + tok * kKStride + d8, kp + tok * kHD + d8);
        cp_async16(vs + tok * kHD + d8, vp + tok * kHD + d8);
      }
    }
  };
  issue(0, n_old);
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += blockDim.x) {          // keys beyond kv_len: zeros
    const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
    if (tok >= nk) {
      *reinterpret_cast<uint4*>(ks + tok * kKStride + d8) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vs + tok * kHD + d8) = make_uint4(0, 0, 0, 0);
    }
  }
  pdl_wait();
  issue(n_old, nk);
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  const int head = g * G + warp;
  {
    const bf16* qp = a.q + (size_t)m * a.Hq * kHD + (size_t)head * kHD;
    uint2 qv = *reinterpret_cast<const uint2*>(qp + lane * 4);
    qs[warp][lane * 4 + 0] = bf16lo(qv.x); qs[warp][lane * 4 + 1] = bf16hi(qv.x);
    qs[warp][lane * 4 + 2] = bf16lo(qv.y); qs[warp][lane * 4 + 3] = bf16hi(qv.y);
  }
  __syncthreads();
  // scores: lane <-> key (2 keys per lane)
  float sc[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int tok = lane + 32 * j;
    float s = 0.f;
#pragma unroll
    for (int d8 = 0; d8 < kHD; d8 += 8) {
      uint4 kv4 = *reinterpret_cast<const uint4*>(ks + tok * kKStride + d8);
      const float4 q0 = *reinterpret_cast<const float4*>(&qs[warp][d8]);
      const float4 q1 = *reinterpret_cast<const float4*>(&qs[warp][d8 + 4]);
      s = fmaf(bf16lo(kv4.x), q0.x, s); s = fmaf(bf16hi(kv4.x), q0.y, s);
      s = fmaf(bf16lo(kv4.y), q0.z, s); s = fmaf(bf16hi(kv4.y), q0.w, s);
      s = fmaf(bf16lo(kv4.z), q1.x, s); s = fmaf(bf16hi(kv4.z), q1.y, s);
      s = fmaf(bf16lo(kv4.w), q1.z, s); s = fmaf(bf16hi(kv4.w), q1.w, s);
    }
    sc[j] = (tok < nk) ? s * a.scale : -INFINITY;
  }
  const float mx = warp_max(fmaxf(sc[0], sc[1]));
  const float p0 = __expf(sc[0] - mx), p1 = __expf(sc[1] - mx);
  const float l = warp_sum(p0 + p1);
  ps[warp][lane] = p0; ps[warp][lane + 32] = p1;
  __syncwarp();
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tok = 0; tok < nk; ++tok) {
    const float p = ps[warp][tok];
    uint2 vv = *reinterpret_cast<const uint2*>(vs + tok * kHD + lane * 4);
    o[0] = fmaf(p, bf16lo(vv.x), o[0]); o[1] = fmaf(p, bf16hi(vv.x), o[1]);
    o[2] = fmaf(p, bf16lo(vv.y), o[2]); o[3] = fmaf(p, bf16hi(vv.y), o[3]);
  }
  float* part = a.part + (((size_t)m * a.Hq + head) * a.nsplit + split) * kPart;
  *reinterpret_cast<float4*>(part + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
  if (lane == 0) { part[kHD] = mx; part[kHD + 1] = l; }
  // ---- last CTA of this (row, kv head) merges the splits ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t* cnt = a.counters + (size_t)m * a.Hkv + g;
    const int prev = atomicAdd(cnt, 1);
    s_last = (prev == nact - 1);
    if (s_last) *cnt = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* base = a.part + (((size_t)m * a.Hq + head) * a.nsplit) * kPart;
  float M = -INFINITY;
  for (int s = 0; s < nact; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * kPart + kHD));
  float L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < nact; ++s) {
    const float* ps_ = base + (size_t)s * kPart;
    const float w = __expf(__ldcg(ps_ + kHD) - M);
    L = fmaf(__ldcg(ps_ + kHD + 1), w, L);
    const float4 ov = __ldcg(reinterpret_cast<const float4*>(ps_ + lane * 4));
    acc[0] = fmaf(ov.x, w, acc[0]); acc[1] = fmaf(ov.y, w, acc[1]);
    acc[2] = fmaf(ov.z, w, acc[2]); acc[3] = fmaf(ov.w, w, acc[3]);
  }
  const float inv = 1.0f / L;
  uint2 outv;
  outv.x = pack_bf16(acc[0] * inv, acc[1] * inv);
  outv.y = pack_bf16(acc[2] * inv, acc[3] * inv);
  *reinterpret_cast<uint2*>(a.y + (size_t)m * a.Hq * kHD + (size_t)head * kHD + lane * 4) = outv;
}


// ================================================================================================================
// Persistent decode step: ONE cooperative launch runs embed -> 26 x (in_proj, attention, out_proj x2, fc1, fc2) ->
// heads for up to 4 activation rows.  148 CTAs (one per SM) stay resident.  The producer warp of every CTA streams that
// CTA's slice of ALL the step's weight matrices back to back through the shared-memory ring - weights do not depend on
// activations, so HBM keeps streaming across phase boundaries.  The 8 consumer warps run the phases; a phase starts as
// soon as the tagged activation words of the previous one have arrived (no grid barrier, see "tagged activation words").
// out_proj is applied twice by the reference (_torch.py:419-420): its slice is held in the ring between the two
// passes, so it is read from HBM once.
// ================================================================================================================
struct MegaLayer {
  const bf16 *norm_w, *norm_b, *in_proj, *out_proj, *norm2_w, *norm2_b, *fc1, *fc2; bf16* kv_layer;
  // Mamba2 layers (kind 1): conv1d / SSM parameters, gated-norm weight, this layer's recurrent state of all rows
  const bf16 *conv_w, *conv_b, *dt_bias, *A_log, *Dp, *mnorm_w; bf16 *conv_state, *ssm_state; long long kind;
  const bf16 *in_t, *out_t, *fc1_t, *fc2_t;                   // the same matrices re-laid out for the tcgen05 consumer (see MegaTcGeo)
  const float *s_in, *s_out, *s_fc1, *s_fc2;                  // FP8 mode: row scales (in_proj ... fc2 then point at e4m3 bytes, [N, K] row-major)
};
// tcgen05 consumer: a matrix [N, K] is cut into units of RB weight rows (a multiple of 8; fc1: RBv value rows followed by
// the RBv gate rows of the same features), unit u belongs to CTA u.  K is cut into S segments of Ks = K / S elements and
// the unit is stored as Ks/64 consecutive tiles of [S * RB rows][64 k]: tile row s * RB + r holds W[row r][s * Ks + 64 t ..]
// ("segment-diagonal", see mega_consume_tc), in the 128-byte-swizzle shared-memory image the UMMA descriptor reads.  A
// unit is ONE contiguous run of RB * K * 2 bytes that 1-D bulk copies move through the ring, kps tiles per 32 KB stage.
struct MegaTcGeo { int RB, RBv, nunits, kps, S; };
enum { TG_QKV = 0, TG_OUT = 1, TG_FC1 = 2, TG_FC2 = 3, TG_HEADS = 4, TG_COUNT = 5 };
// what a thread's in_proj epilogue item needs besides the dot products: position, RoPE cos/sin, KV page.  The same in
// every layer of a step, so it is fetched once per step (three dependent global loads otherwise trail every in_proj)
struct MegaQkvPre { int pos, page; float2 cs; bool valid; };

struct MegaArgs {
  const MegaLayer* layers; int n_layer;
  int D, F, Hq, Hkv, hd; float eps; int norm_kind, rope_interleaved, out_proj_repeats;
  const bf16 *normf_w, *normf_b, *heads; int QV, B; float cfg_scale; float* logits;
  const float* rope; int rope_len;
  const int32_t* lengths; const int32_t* page_table; int max_pages;
  const bf16* emb[16]; int Q, vocab; const int64_t* delayed; int T_delayed;
  // activations exchanged between CTAs: one 32-bit word per element = bf16 value (high half) | 16-bit phase tag
  uint32_t *xt, *qt, *ayt, *y1t, *ht, *kvt;
  float* attn_part; int32_t* attn_counters; int nsplit; float scale;
  const zb_loop_state* loop;
  unsigned* sync;         // [1] = epoch: number of live steps this session's tagged buffers have seen
  int ring_stages, part_bytes, evict_first;
  unsigned long long* timeline;   // debug: globaltimer stamps of CTA 0 (2 per phase: inputs ready, work done)
  unsigned long long* steplog;    // debug: [2*step] start, [2*step+1] end of every step (CTA 0)
  // tcgen05 consumer (decode_step_kernel<R, true>)
  const bf16* heads_t; MegaTcGeo tg[TG_COUNT];
  const float* s_heads;   // FP8 mode (decode_step_kernel<R, 2>): row scales of the heads (`heads` then points at e4m3 bytes)
  // hybrid stacks: Mamba2 layers run as three phases (in_proj, conv1d step + state update, gated norm + out_proj)
  int nph, d_inner, m_nheads, ipo, conv_dim; uint32_t *zxt, *ygt;
};

constexpr int kMegaStageBytes = 32 * 1024, kMegaAttnBytes = 40 * 1024;
constexpr int kMegaTcImageBytes = 32 * 1024, kMegaTcPartBytes = 4 * 1024;   // tcgen05 consumer: activation image (32 k blocks x 1 KB), fp32 accumulator copy
// consumer warps of the persistent kernel (+ 1 producer warp).  Nine warps leave 168 registers per thread (17 warps: 96,
// with spills), and everything a warp does redundantly (norm statistics, index math, barrier and ring bookkeeping, the
// per-stage reduction tree) is issued half as often: the warps share 4 issue slots.
constexpr int kMW = 8;

// ---- tagged activation words -----------------------------------------------------------------------------------
// The phases of a step depend on each other all-to-all (every CTA needs the whole activation vector the previous
// phase produced).  Instead of a grid barrier followed by a load (release fence + atomic + poll + load = four L2
// round trips on the critical path), every activation element travels as a self-validating 32-bit word: the writer
// stores {bf16 value, tag of the writing phase} with one relaxed store, the readers spin on the very loads that
// fetch their operands until every word carries the expected tag.  No fence is needed (nothing but the word itself
// is published), a 32-bit store is single-copy atomic, and the tag sequence never repeats for a buffer (it is
// rewritten every live step; the buffers belong to one generate session and start zeroed, tag 0 is never used).
// A buffer is only rewritten by a phase that cannot start before every reader of the old contents is done: each
// phase needs ALL outputs of the phase before it, so no CTA is ever more than one phase ahead of the slowest.
__device__ __forceinline__ uint32_t mega_tag(unsigned epoch, int nph, int ph) {
  return (((epoch % 65535u) * (unsigned)nph + (unsigned)ph) % 65535u) + 1u;
}
__device__ __forceinline__ uint32_t tag_word(float v, uint32_t tag) {
  return ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)) << 16) | tag;
}
__device__ __forceinline__ float untag(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ bool tags_ok(const uint4& v, uint32_t tag) {
  return ((v.x & 0xffffu) == tag) & ((v.y & 0xffffu) == tag) & ((v.z & 0xffffu) == tag) & ((v.w & 0xffffu) == tag);
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_v4(uint32_t* p, const uint4& v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
constexpr unsigned kMegaSpinLimit = 1u << 21;                // a lost CTA must end in an error, not a hung GPU
// spin until the 4 words at p carry `tag`
__device__ __forceinline__ uint4 poll_v4(const uint32_t* p, uint32_t tag) {
  uint4 v = ld_relaxed_v4(p);
  for (unsigned spins = 0; !tags_ok(v, tag); ++spins) {
    if (spins > kMegaSpinLimit) asm volatile("trap;");
    v = ld_relaxed_v4(p);
  }
  return v;
}

template <int EPI>
__device__ __forceinline__ void mega_slice(const GemvArgs& a, int& u_begin, int& nrows) {
  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);
  const int nunits = units_total<EPI>(a);
  u_begin = (int)((long long)blockIdx.x * nunits / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * nunits / gridDim.x);
  nrows = (u_end - u_begin) * (kPairs ? 2 : 1);
}

// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst).  wsz = bytes per weight
// (2: bf16; 1: the e4m3 copy of the FP8 mode); rmult = rows of a stage relative to the bf16 stage (FP8 mode with R <= 2:
// twice the rows in the same 32 KB - a stage costs the consumer one latency chain whatever it holds; R = 4 keeps the bf16
// rows in 16 KB stages: 32 accumulators per lane would not fit the reduction tree).
template <int EPI>
__device__ __forceinline__ void mega_produce(const GemvArgs& a, unsigned char* ring, uint64_t* full_bar, uint64_t* empty_bar, int S, int& gst,
                                             uint64_t pol, int lane, int wsz = 2, int rmult = 1) {
  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);
  const int K = a.K, row_bytes = K * wsz, RPS = kMegaStageBytes / (K * 2) * rmult;
  const uint32_t stage_bytes = (uint32_t)(RPS * row_bytes);   // == kMegaStageBytes for bf16 and for e4m3 with twice the rows (rmult = 2)
  const unsigned char* W8 = reinterpret_cast<const unsigned char*>(a.W);
  int u_begin, nrows;
  mega_slice<EPI>(a, u_begin, nrows);
  const int nstage = (nrows + RPS - 1) / RPS;
  for (int st = 0; st < nstage; ++st, ++gst) {
    const int slot = gst % S;
    if (gst >= S) mbar_wait(&empty_bar[slot], ((gst / S) - 1) & 1);
    const int r0 = st * RPS;
    if (lane == 0) mbar_expect_tx(&full_bar[slot], stage_bytes);
    __syncwarp();
    unsigned char* dst = ring + (size_t)slot * stage_bytes;
    if (!kPairs && r0 + RPS <= nrows) {
      if (lane == 0) {
        if (pol) bulk_g2s(dst, W8 + (size_t)(u_begin + r0) * row_bytes, stage_bytes, &full_bar[slot], pol);
        else bulk_g2s_nohint(dst, W8 + (size_t)(u_begin + r0) * row_bytes, stage_bytes, &full_bar[slot]);
      }
    } else {
      for (int q = lane; q < RPS; q += 32) {
        const int lr = min(r0 + q, nrows - 1);
        const unsigned char* src = W8 + (size_t)row_of_local<EPI>(a, u_begin, lr) * row_bytes;
        if (pol) bulk_g2s(dst + (size_t)q * row_bytes, src, row_bytes, &full_bar[slot], pol);
        else bulk_g2s_nohint(dst + (size_t)q * row_bytes, src, row_bytes, &full_bar[slot]);
      }
    }
  }
}

// ---- FP8 mode of the FFMA consumer (decode_step_kernel<R, 2>, opt-in: ZB_FP8=1) ---------------
// Weights: e4m3 bytes [N, K] + one power-of-two fp32 scale per row with |w| / scale < 2, so that (a) the dequantised
// weight q * scale is exactly a bf16 number - the FP8 kernel can be checked against the bf16 kernel run on the
// dequantised weights - and (b) products stay far inside the f16 range.  An e4m3 pair becomes an f16x2 with ONE cvt
// (exact) and meets the activation pair in HFMA2: 4 cvt + 4 R HFMA2 per 8 weights against 8 shifts + 4 R FFMA2 for
// bf16 (converting e4m3 to fp32 would cost 12 and lose to the bf16 path per weight).  A lane's f16 chain is NC * 4
// products long (8 or 16) before it is widened to fp32; stage sums, the K-slice reduction and the epilogue are fp32 as before.
__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint32_t pair16) {
  uint32_t r;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"((unsigned short)pair16));   // low byte -> low half
  return r;
}
__device__ __forceinline__ void hfma2_acc(uint32_t& d, uint32_t a, uint32_t b) {
  asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float sum_f16x2(uint32_t v) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
  return f.x + f.y;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}

// consumers: one matrix phase.  wait_full / release control the out_proj "hold" (first pass keeps the slots).
template <int R, int NC, int RW, int PRO, int EPI, bool F8 = false>
__device__ __forceinline__ void mega_consume(const GemvArgs& a, unsigned char* ring, float* part, uint64_t* full_bar, uint64_t* empty_bar,
                                             float (*red)[kMW][4], int S, int& gst, bool release, int warp, int lane,
                                             const uint32_t* xt, uint32_t tag_in, uint32_t* yt, uint32_t tag_out, const uint32_t* rt, uint32_t* qt,
                                             uint32_t* kvt, unsigned long long* stamp, int norm_pending = 0, const MegaQkvPre* qkv_pre = nullptr, unsigned long long* dbg = nullptr,
                                             const uint32_t* zt = nullptr, int ldz = 0, uint32_t tag_z = 0u) {
#define DBG(i) do { if (dbg && threadIdx.x == 0) dbg[i] = gtime(); } while (0)
  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);
  constexpr int Kc = NC * 256;
  constexpr int kWsz = F8 ? 1 : 2;                             // bytes per weight
  const int K = a.K, row_bytes = K * kWsz;
  const int KS = K / Kc;
  const int RPS = (kMW / KS) * RW;
  const uint32_t kStage = F8 ? (uint32_t)(RPS * row_bytes) : (uint32_t)kMegaStageBytes;   // FP8: 16 KB (RW as bf16) or 32 KB (RW doubled)
  int u_begin, nrows;
  mega_slice<EPI>(a, u_begin, nrows);
  const int nstage = (nrows + RPS - 1) / RPS;
  const int ks = warp % KS, rg = warp / KS;
  const size_t koff = (size_t)ks * Kc;
  // activations: spin on the operand loads themselves until every word carries the producing phase's tag
  float xf[R][NC * 8];
  for (unsigned spins = 0;; ++spins) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const uint32_t* src = xt + (size_t)i * a.ldx + koff + c * 256 + lane * 8;
        const uint4 v0 = ld_relaxed_v4(src), v1 = ld_relaxed_v4(src + 4);
        ok = ok && tags_ok(v0, tag_in) && tags_ok(v1, tag_in);
        xf[i][c * 8 + 0] = untag(v0.x); xf[i][c * 8 + 1] = untag(v0.y); xf[i][c * 8 + 2] = untag(v0.z); xf[i][c * 8 + 3] = untag(v0.w);
        xf[i][c * 8 + 4] = untag(v1.x); xf[i][c * 8 + 5] = untag(v1.y); xf[i][c * 8 + 6] = untag(v1.z); xf[i][c * 8 + 7] = untag(v1.w);
      }
    if (ok) break;
    if (spins > kMegaSpinLimit) asm volatile("trap;");
  }
  if (PRO == PRO_GATED) {
    // mamba_ssm RMSNormGated (norm_before_gate = False): g = y * silu(z) stays fp32 up to xn = bf16(g * rsqrt(mean(g^2) + 1e-5) * w).
    // The scan phase publishes g as two tagged 16-bit halves (xt: high halves, already in xf; zt: low halves).
    for (unsigned spins = 0;; ++spins) {
      bool ok = true;
      uint32_t lo[R][NC * 8];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint32_t* src = zt + (size_t)i * ldz + koff + c * 256 + lane * 8;
          const uint4 v0 = ld_relaxed_v4(src), v1 = ld_relaxed_v4(src + 4);
          ok = ok && tags_ok(v0, tag_z) && tags_ok(v1, tag_z);
          lo[i][c * 8 + 0] = v0.x; lo[i][c * 8 + 1] = v0.y; lo[i][c * 8 + 2] = v0.z; lo[i][c * 8 + 3] = v0.w;
          lo[i][c * 8 + 4] = v1.x; lo[i][c * 8 + 5] = v1.y; lo[i][c * 8 + 6] = v1.z; lo[i][c * 8 + 7] = v1.w;
        }
      if (ok) {
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
          for (int e = 0; e < NC * 8; ++e) xf[i][e] = __uint_as_float(__float_as_uint(xf[i][e]) | (lo[i][e] >> 16));
        break;
      }
      if (spins > kMegaSpinLimit) asm volatile("trap;");
    }
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();
  DBG(0);
  if (PRO == PRO_NORM || PRO == PRO_GATED) {
    float mean[R], rstd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float sacc = 0.f, qacc = 0.f;
#pragma unroll
      for (int e = 0; e < NC * 8; ++e) { sacc += xf[i][e]; qacc = fmaf(xf[i][e], xf[i][e], qacc); }
      sacc = warp_sum(sacc);
      qacc = warp_sum(qacc);
      if (rg == 0 && lane == 0) { red[0][ks][i] = sacc; red[1][ks][i] = qacc; }
    }
    // the norm parameters sit in shared memory (copied there a layer ahead: a global load issued here would queue
    // behind the saturated weight stream for microseconds); this thread's copies are complete after the wait, all
    // threads' after the barrier
    DBG(1);
    if (norm_pending == 0) asm volatile("cp.async.wait_group 0;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
    DBG(2);
    asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
    DBG(3);
    // every warp sums the KS per-warp partials with a fixed shuffle tree (lane q holds slice q): far fewer issue
    // slots than a serial loop in each of the 16 warps, and the same bits in every warp
    const float inv_k = 1.0f / (float)K;                        // K is a power of two: multiplying is exact
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float tot = warp_sum(lane < KS ? red[0][lane][i] : 0.f);
      const float tsq = warp_sum(lane < KS ? red[1][lane][i] : 0.f);
      const float mu = tot * inv_k;
      const bool ln = PRO == PRO_NORM && a.norm_kind == ZB_NORM_LAYERNORM;     // (the gated norm is an RMS norm with its own eps)
      mean[i] = ln ? mu : 0.f;
      const float var = ln ? fmaxf(tsq * inv_k - mu * mu, 0.f) : tsq * inv_k;
      rstd[i] = rsqrtf(var + (PRO == PRO_GATED ? 1e-5f : a.eps));
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const size_t kk = koff + c * 256 + lane * 8;
      const uint4 nwv = *reinterpret_cast<const uint4*>(a.nw + kk);
      const uint4 nbv = a.nb ? *reinterpret_cast<const uint4*>(a.nb + kk) : make_uint4(0, 0, 0, 0);
      const uint32_t gv[4] = {nwv.x, nwv.y, nwv.z, nwv.w}, bv[4] = {nbv.x, nbv.y, nbv.z, nbv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = bf16lo(gv[j]), g1 = bf16hi(gv[j]), b0 = bf16lo(bv[j]), b1 = bf16hi(bv[j]);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          xf[i][c * 8 + 2 * j] = rbf((xf[i][c * 8 + 2 * j] - mean[i]) * rstd[i] * g0 + b0);
          xf[i][c * 8 + 2 * j + 1] = rbf((xf[i][c * 8 + 2 * j + 1] - mean[i]) * rstd[i] * g1 + b1);
        }
      }
    }
  }
  DBG(4);
  unsigned long long x2[F8 ? 1 : R][F8 ? 1 : NC * 4];
  uint32_t xh[F8 ? R : 1][F8 ? NC * 4 : 1];                   // FP8 mode: the activation pairs as f16x2
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int e = 0; e < NC * 4; ++e) {
      if (F8) xh[F8 ? i : 0][F8 ? e : 0] = pack_f16x2_sat(xf[i][2 * e], xf[i][2 * e + 1]);
      else x2[F8 ? 0 : i][F8 ? 0 : e] = pack_f32x2(xf[i][2 * e], xf[i][2 * e + 1]);
    }

  // ---- operands of this thread's epilogue (residual value, RoPE cos/sin, KV page): fetched NOW so their L2 round
  // trips overlap the weight streaming instead of trailing it ----
  const bool cfg = (EPI == EPI_HEADS && a.cfg_scale != 1.0f);
  const int rows_out = cfg ? a.B : R;                          // a.M == R in the persistent kernel
  const int nu = kPairs ? nrows / 2 : nrows;
  const int et = threadIdx.x;                                 // one epilogue item per thread (host guarantees nu*rows_out <= 512)
  const bool e_on = et < nu * rows_out;
  const int ej = e_on ? et / rows_out : 0, ei = e_on ? et % rows_out : 0;
  int en0 = 0, en1 = 1;
  if (kPairs) unit_rows<EPI>(a, u_begin + ej, en0, en1); else { en0 = u_begin + ej; en1 = en0 + 1; }
  float pre_resid = 0.f;
  float2 pre_cs = make_float2(1.f, 0.f);
  int pre_pos = 0, pre_page = 0;
  float wsc0 = 1.f, wsc1 = 1.f;                                // FP8 mode: scales of this item's weight rows
  if (F8 && e_on) { wsc0 = __ldg(a.wscale + en0); if (kPairs) wsc1 = __ldg(a.wscale + en1); }
  if (e_on) {
    if (EPI == EPI_RESID) pre_resid = untag(ld_relaxed_u32(rt + (size_t)ei * a.ldr + en0));   // validated by this CTA in an earlier phase
    if (EPI == EPI_QKV && qkv_pre && qkv_pre->valid) {
      pre_pos = qkv_pre->pos; pre_page = qkv_pre->page; pre_cs = qkv_pre->cs;
    } else if (EPI == EPI_QKV) {
      pre_pos = a.lengths[ei];
      const int qn_ = a.Hq * a.hd, kn_ = a.Hkv * a.hd;
      if (en0 < qn_ + kn_) {
        const int ri = a.rope_interleaved ? (en0 % a.hd) / 2 : (en0 % a.hd);
        pre_cs = *reinterpret_cast<const float2*>(a.rope + ((size_t)min(pre_pos, a.rope_len - 1) * (a.hd / 2) + ri) * 2);
      }
      if (en0 >= qn_) pre_page = a.page_table[(size_t)ei * a.max_pages + pre_pos / ZB_PAGE_TOKENS];
    }
  }

  constexpr int V = RW * R;
  const int my_idx = multi_reduce_index<V>(lane);
  const bool writer = (lane & ((32 / V) - 1)) == 0;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(rg * RW) * row_bytes + (uint32_t)(koff + lane * 8) * kWsz;
  const uint32_t part_lane = smem_u32(part) + (uint32_t)((((rg * RW + my_idx / R) * KS + ks) * R + my_idx % R) * 4);
  const uint32_t part_stage = (uint32_t)(RPS * KS * R * 4);
  uint32_t slot = (uint32_t)(gst % S), phase = (uint32_t)((gst / S) & 1);   // advanced incrementally: no division per stage
  uint32_t src = lane_base + slot * kStage, fb = full0 + slot * 8, eb = empty0 + slot * 8;
  uint32_t part_dst = part_lane;
  gst += nstage;
  for (int st = 0; st < nstage; ++st) {
    mbar_wait_u32(fb, phase);
    if (st == 0) DBG(5);
    float acc[V];
    if (F8) {
      uint32_t acch[V];
#pragma unroll
      for (int q = 0; q < V; ++q) acch[q] = 0u;
#pragma unroll
      for (int w = 0; w < RW; ++w) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint2 wv = lds64(src + w * row_bytes + c * 256);    // this lane's 8 weights of chunk c: k = koff + 256 c + 8 lane ...
          const uint32_t c0[2] = {wv.x, wv.y};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {                           // pair jj = weights (2 jj, 2 jj + 1) = one half of c0[jj / 2]
            const uint32_t w2 = e4m3x2_to_f16x2(c0[jj >> 1] >> ((jj & 1) * 16));
#pragma unroll
            for (int i = 0; i < R; ++i) hfma2_acc(acch[w * R + i], w2, xh[F8 ? i : 0][F8 ? c * 4 + jj : 0]);
          }
        }
      }
      __syncwarp();
      if (release && lane == 0) mbar_arrive_u32(eb);
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] = sum_f16x2(acch[q]);
    } else {
      unsigned long long acc2[V];
#pragma unroll
      for (int q = 0; q < V; ++q) acc2[q] = 0ull;
#pragma unroll
      for (int w = 0; w < RW; ++w) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint4 wv = lds128(src + w * row_bytes + c * 512);
          const uint32_t c0[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const unsigned long long w2 = pack_f32x2(bf16lo(c0[jj]), bf16hi(c0[jj]));
#pragma unroll
            for (int i = 0; i < R; ++i) ffma2(acc2[w * R + i], w2, x2[F8 ? 0 : i][F8 ? 0 : c * 4 + jj]);
          }
        }
      }
      __syncwarp();
      if (release && lane == 0) mbar_arrive_u32(eb);
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] = sum_f32x2(acc2[q]);
    }
    warp_reduce_multi<V>(acc);
    if (writer) sts32(part_dst, acc[0]);
    part_dst += part_stage;
    if (++slot == (uint32_t)S) { slot = 0; phase ^= 1u; src = lane_base; fb = full0; eb = empty0; }
    else { src += kStage; fb += 8; eb += 8; }
  }
  DBG(6);
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
  DBG(7);

  if (e_on) {
    float v0 = 0.f, v1 = 0.f, u0 = 0.f;
    if (kPairs) {
      const float* s0 = part + (size_t)(2 * ej) * KS * R;
      const float* s1 = s0 + (size_t)KS * R;
      for (int q = 0; q < KS; ++q) { v0 += s0[q * R + ei]; v1 += s1[q * R + ei]; }
    } else {
      const float* s0 = part + (size_t)ej * KS * R;
      for (int q = 0; q < KS; ++q) { v0 += s0[q * R + ei]; if (cfg) u0 += s0[q * R + a.B + ei]; }
    }
    if (F8) { v0 *= wsc0; v1 *= wsc1; u0 *= wsc0; }           // power-of-two row scales: exact
    if (EPI == EPI_RESID) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(pre_resid + rbf(v0), tag_out));
    } else if (EPI == EPI_STORE) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(v0, tag_out));
    } else if (EPI == EPI_SILU) {                            // same ops as gemv_epilogue<EPI_SILU>
      const float yv = rbf(v0), g = rbf(v1);
      const float sg = rbf(g / (1.0f + expf(-g)));
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(__fmul_rn(yv, sg), tag_out));
    } else if (EPI == EPI_QKV) {
      const int qn_ = a.Hq * a.hd, kn_ = a.Hkv * a.hd;
      float o0 = rbf(v0), o1 = rbf(v1);
      if (en0 < qn_ + kn_) {                                 // same un-contracted fp32 ops as gemv_epilogue / _torch.py:57-68
        const float r0 = __fsub_rn(__fmul_rn(o0, pre_cs.x), __fmul_rn(o1, pre_cs.y));
        const float r1 = __fadd_rn(__fmul_rn(o1, pre_cs.x), __fmul_rn(o0, pre_cs.y));
        o0 = r0; o1 = r1;
      }
      if (en0 < qn_) {
        st_relaxed_u32(qt + (size_t)ei * qn_ + en0, tag_word(o0, tag_out));
        st_relaxed_u32(qt + (size_t)ei * qn_ + en1, tag_word(o1, tag_out));
      } else {
        const int kvsel = en0 < qn_ + kn_ ? 0 : 1;
        const int c0i = en0 - qn_ - kvsel * kn_, c1i = en1 - qn_ - kvsel * kn_;
        // this step's attention reads the new token from the tagged side buffer; the cache copy is for later steps
        st_relaxed_u32(kvt + ((size_t)ei * 2 + kvsel) * kn_ + c0i, tag_word(o0, tag_out));
        st_relaxed_u32(kvt + ((size_t)ei * 2 + kvsel) * kn_ + c1i, tag_word(o1, tag_out));
        bf16* pb = a.kv_layer + ((size_t)pre_page * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;
        const int tk = pre_pos % ZB_PAGE_TOKENS;
        pb[((size_t)(c0i / a.hd) * ZB_PAGE_TOKENS + tk) * a.hd + (c0i % a.hd)] = f2bf(o0);
        pb[((size_t)(c1i / a.hd) * ZB_PAGE_TOKENS + tk) * a.hd + (c1i % a.hd)] = f2bf(o1);
      }
    } else {
      gemv_epilogue<EPI>(a, ei, en0, en1, kPairs, v0, v1, u0, 0.f);
    }
  }
  DBG(8);
#undef DBG
}

// ================================================================================================================
// tcgen05 consumer of the persistent kernel (decode_step_kernel<R, true>).
//
// The FFMA2 consumer above needs ~0.55 us per 32 KB stage whatever feeds it; the tensor pipe drains a stage in a few
// hundred cycles, so a phase is paced by HBM alone and what the ring prefetched while the previous phase exchanged its
// activations is consumed at once.  "Swapped" operand roles: the activation rows are the UMMA A operand (M = 64), the
// CTA's weight rows are the B operand (N = any multiple of 8 - so a matrix is cut by ROWS over all CTAs without a K
// split across CTAs), the fp32 accumulator sits in tensor memory.
//   all threads   poll the tagged activation words, normalise (PRO_NORM), write them as bf16 into the A image
//                 [tile][S * R rows][64 k] (128-byte swizzle)
//   warp 1        waits for ring stages and issues the MMAs; tcgen05.commit releases stages and signals the end
//   all warps     read their TMEM quarter (tcgen05.ld) into the fp32 `part` buffer, then every thread runs one epilogue item
// ================================================================================================================
__device__ __forceinline__ void mega_produce_tc(const bf16* Wt, const MegaTcGeo& g, int K, unsigned char* ring, uint64_t* full_bar, uint64_t* empty_bar,
                                                int S, int& gst, uint64_t pol, int lane) {
  if ((int)blockIdx.x >= g.nunits) return;                    // this CTA has no unit of the matrix (consumers skip it alike)
  const int nkb = K / (64 * g.S), nstage = (nkb + g.kps - 1) / g.kps;      // tiles of [S * RB rows][64 k]
  const size_t tile = (size_t)g.S * g.RB * 128;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(Wt) + (size_t)blockIdx.x * nkb * tile;
  for (int st = 0; st < nstage; ++st, ++gst) {
    const int slot = gst % S;
    if (gst >= S) mbar_wait(&empty_bar[slot], ((gst / S) - 1) & 1);
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(min(g.kps, nkb - st * g.kps) * tile);
      mbar_expect_tx(&full_bar[slot], bytes);
      unsigned char* dst = ring + (size_t)slot * kMegaStageBytes;
      const unsigned char* src = base + (size_t)st * g.kps * tile;
      if (pol) bulk_g2s(dst, src, bytes, &full_bar[slot], pol);
      else bulk_g2s_nohint(dst, src, bytes, &full_bar[slot]);
    }
    __syncwarp();
  }
}

struct MegaTcState { uint64_t *acc_bar, *afree_bar; int nacc, nafree; uint32_t tmem; unsigned char* abuf; };

template <int R, int NR, int PRO, int EPI>
__device__ __forceinline__ void mega_consume_tc(const GemvArgs& a, const MegaTcGeo& g, unsigned char* ring, float* part, uint64_t* full_bar,
                                                uint64_t* empty_bar, MegaTcState& ts, float (*red)[kMW][4], int S, int& gst, bool release, int warp,
                                                int lane, const uint32_t* xt, uint32_t tag_in, uint32_t* yt, uint32_t tag_out, const uint32_t* rt,
                                                uint32_t tag_resid, uint32_t* qt, uint32_t* kvt, unsigned long long* stamp, int norm_pending,
                                                const MegaQkvPre* qkv_pre, unsigned long long* dbg = nullptr) {
#define DBG(i) do { if (dbg && threadIdx.x == 0) dbg[i] = gtime(); } while (0)
  const int tid = threadIdx.x;                                // consumer threads 0 .. 255
  if ((int)blockIdx.x >= g.nunits) return;                    // no unit of this matrix (the producer skipped it alike)
  const int K = a.K, nchunk8 = K / 8;
  // activations: thread t holds elements [8 q, 8 q + 8) of every row for q = t + 256 c; spin on the operand loads
  // themselves until every word carries the producing phase's tag
  float xf[R][NR * 8];
  for (unsigned spins = 0;; ++spins) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        const int q = tid + 256 * c;
        if (q < nchunk8) {
          const uint32_t* src = xt + (size_t)i * a.ldx + (size_t)q * 8;
          const uint4 v0 = ld_relaxed_v4(src), v1 = ld_relaxed_v4(src + 4);
          ok = ok && tags_ok(v0, tag_in) && tags_ok(v1, tag_in);
          xf[i][c * 8 + 0] = untag(v0.x); xf[i][c * 8 + 1] = untag(v0.y); xf[i][c * 8 + 2] = untag(v0.z); xf[i][c * 8 + 3] = untag(v0.w);
          xf[i][c * 8 + 4] = untag(v1.x); xf[i][c * 8 + 5] = untag(v1.y); xf[i][c * 8 + 6] = untag(v1.z); xf[i][c * 8 + 7] = untag(v1.w);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) xf[i][c * 8 + e] = 0.f;
        }
      }
    if (ok) break;
    if (spins > kMegaSpinLimit) asm volatile("trap;");
  }
  if (stamp && tid == 0) *stamp = gtime();
  DBG(0);
  if (PRO == PRO_NORM) {
    float mean[R], rstd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float sacc = 0.f, qacc = 0.f;
#pragma unroll
      for (int e = 0; e < NR * 8; ++e) { sacc += xf[i][e]; qacc = fmaf(xf[i][e], xf[i][e], qacc); }
      sacc = warp_sum(sacc);
      qacc = warp_sum(qacc);
      if (lane == 0) { red[0][warp][i] = sacc; red[1][warp][i] = qacc; }
    }
    // norm parameters were copied to shared memory a layer ahead (see mega_consume)
    if (norm_pending == 0) asm volatile("cp.async.wait_group 0;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
    const float inv_k = 1.0f / (float)K;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float tot = warp_sum(lane < kMW ? red[0][lane][i] : 0.f);     // same shuffle tree, same bits in every warp
      const float tsq = warp_sum(lane < kMW ? red[1][lane][i] : 0.f);
      const float mu = tot * inv_k;
      mean[i] = (a.norm_kind == ZB_NORM_LAYERNORM) ? mu : 0.f;
      const float var = (a.norm_kind == ZB_NORM_LAYERNORM) ? fmaxf(tsq * inv_k - mu * mu, 0.f) : tsq * inv_k;
      rstd[i] = rsqrtf(var + a.eps);
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      const int q = tid + 256 * c;
      if (q < nchunk8) {
        const uint4 nwv = *reinterpret_cast<const uint4*>(a.nw + (size_t)q * 8);
        const uint4 nbv = a.nb ? *reinterpret_cast<const uint4*>(a.nb + (size_t)q * 8) : make_uint4(0, 0, 0, 0);
        const uint32_t gv[4] = {nwv.x, nwv.y, nwv.z, nwv.w}, bv[4] = {nbv.x, nbv.y, nbv.z, nbv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float g0 = bf16lo(gv[j]), g1 = bf16hi(gv[j]), b0 = bf16lo(bv[j]), b1 = bf16hi(bv[j]);
#pragma unroll
          for (int i = 0; i < R; ++i) {
            xf[i][c * 8 + 2 * j] = rbf((xf[i][c * 8 + 2 * j] - mean[i]) * rstd[i] * g0 + b0);
            xf[i][c * 8 + 2 * j + 1] = rbf((xf[i][c * 8 + 2 * j + 1] - mean[i]) * rstd[i] * g1 + b1);
          }
        }
      }
    }
  }

  // ---- this thread's epilogue item and its operands (residual value, RoPE cos/sin, KV page): fetched NOW so their L2
  // round trips overlap the weight streaming instead of trailing it ----
  const bool cfg = (EPI == EPI_HEADS && a.cfg_scale != 1.0f);
  const int rows_out = cfg ? a.B : R;
  const int nu = (EPI == EPI_QKV) ? g.RB / 2 : (EPI == EPI_SILU) ? g.RBv : g.RB;
  const int ej = tid / rows_out, ei = tid % rows_out;
  int en0, en1, lr0, lr1;
  if (EPI == EPI_QKV) { lr0 = 2 * ej; lr1 = lr0 + 1; en0 = (int)blockIdx.x * g.RB + lr0; en1 = en0 + 1; }
  else if (EPI == EPI_SILU) { lr0 = ej; lr1 = g.RBv + ej; en0 = (int)blockIdx.x * g.RBv + ej; en1 = en0 + a.F; }
  else { lr0 = lr1 = ej; en0 = en1 = (int)blockIdx.x * g.RB + ej; }
  const bool e_on = tid < nu * rows_out && en0 < (EPI == EPI_SILU ? a.F : a.N);
  float pre_resid = 0.f;
  float2 pre_cs = make_float2(1.f, 0.f);
  int pre_pos = 0, pre_page = 0;
  if (e_on) {
    if (EPI == EPI_RESID) {                                   // written by an earlier phase; its tag is checked (one word, normally there already)
      uint32_t w = ld_relaxed_u32(rt + (size_t)ei * a.ldr + en0);
      for (unsigned spins = 0; (w & 0xffffu) != tag_resid; ++spins) {
        if (spins > kMegaSpinLimit) asm volatile("trap;");
        w = ld_relaxed_u32(rt + (size_t)ei * a.ldr + en0);
      }
      pre_resid = untag(w);
    }
    if (EPI == EPI_QKV && qkv_pre && qkv_pre->valid) { pre_pos = qkv_pre->pos; pre_page = qkv_pre->page; pre_cs = qkv_pre->cs; }
  }

  DBG(1);
  // ---- A image + MMAs ----
  // Segment-diagonal GEMV: K is cut into S segments of Ks elements.  A row (s, i) = segment s of activation row i, B row
  // (s', r) = segment s' of weight row r, so D[(s, i)][(s', r)] is the partial dot product over segment s for s == s' (the
  // other blocks are never read) and y[i][r] = sum_s D[(s, i)][(s, r)].  One MMA then covers S * RB weight rows x 16 k
  // (up to 8 KB) instead of RB x 16 (512 B): the tensor pipe needs ~(M + N) * 32 B / 128 cycles per MMA whatever is in the
  // rows (measured, scripts/probes/mma_rate_probe.cu), far too slow for 16-row slices with nothing but real data in B.
  // Accumulator row (s, i) is placed in TMEM quarter s % NQ, slot (s / NQ) * R + i of its 16 rows (M = 64 keeps 16 rows per
  // 32-lane quarter): every quarter then needs the columns of only S / NQ segments on the way back, on two warps each.
  const int SG = g.S, Ks = K / SG, nt = Ks / 64;               // tiles ("k' blocks") of the unit
  const int NQ = min(4, SG), spq = (SG / NQ) * R;              // quarters in use, slots per quarter (<= 16)
  const bool wide = spq > 8;
  // image of one tile: quarter q = UMMA row groups 2q, 2q + 1.  With <= 8 slots only the even groups hold data: a stride
  // of 512 bytes between groups puts them 1 KB apart (the odd groups overlap them and land in lanes nobody reads)
  const uint32_t QS = wide ? 2048u : 1024u, TA = (uint32_t)NQ * QS, sbo = wide ? 1024u : 512u;
  const int TPC = min(nt, (int)(kMegaTcImageBytes / TA));      // tiles per image
  const int gst0 = gst;
  gst += (nt + g.kps - 1) / g.kps;
  const uint32_t a_base = smem_u32(ts.abuf);
  const uint32_t idesc = make_idesc(64, SG * g.RB);
  const uint32_t tile = (uint32_t)(SG * g.RB) * 128u;
  int slot = gst0 % S, w = 0;                                  // ring position of tile t (warp-uniform, advanced by every thread alike)
  uint32_t par = (uint32_t)((gst0 / S) & 1);
  for (int t0 = 0; t0 < nt; t0 += TPC) {
    if (t0 > 0) {                                              // the MMAs of the previous image have read it
      mbar_wait(ts.afree_bar, (uint32_t)(ts.nafree & 1));
      ++ts.nafree;
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      const int q = tid + 256 * c;
      if (q < nchunk8) {
        const int k0 = q * 8, sg = k0 / Ks, kk = k0 - sg * Ks, t = kk >> 6, j = (kk & 63) >> 3;
        if (t >= t0 && t < t0 + TPC) {
#pragma unroll
          for (int i = 0; i < R; ++i) {
            const int qd = sg & (NQ - 1), sl = (sg / NQ) * R + i;
            const uint32_t dst = a_base + (uint32_t)(t - t0) * TA + (uint32_t)qd * QS + (uint32_t)(sl >> 3) * 1024u + (uint32_t)(sl & 7) * 128u +
                                 (uint32_t)((j ^ (sl & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pack_bf16(xf[i][c * 8 + 0], xf[i][c * 8 + 1])),
                         "r"(pack_bf16(xf[i][c * 8 + 2], xf[i][c * 8 + 3])), "r"(pack_bf16(xf[i][c * 8 + 4], xf[i][c * 8 + 5])),
                         "r"(pack_bf16(xf[i][c * 8 + 6], xf[i][c * 8 + 7]))
                         : "memory");
          }
        }
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
    if (t0 == 0) DBG(2);
    const int t1 = min(nt, t0 + TPC);
    if (warp_id_uniform() == 1) {                             // the whole warp walks the ring on uniform values, one elected lane issues (tc.cuh: elect_one)
      tc_fence_after();
      uint32_t a_tile = a_base;
      for (int t = t0; t < t1; ++t, a_tile += TA) {
        if (w == 0 || t == t0) { mbar_wait(&full_bar[slot], par); tc_fence_after(); }
        const uint64_t da = make_smem_desc_sbo(a_tile, sbo);
        const uint64_t db = make_smem_desc(smem_u32(ring) + (uint32_t)slot * kMegaStageBytes + (uint32_t)w * tile);
        const bool last_of_stage = (w == g.kps - 1 || t == nt - 1);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(ts.tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (t | kk) != 0 ? 1u : 0u);
          if (release && last_of_stage) tc_commit(&empty_bar[slot]);
        }
        __syncwarp();
        if (last_of_stage) { w = 0; if (++slot == S) { slot = 0; par ^= 1u; } } else ++w;
      }
      if (elect_one()) tc_commit(t1 == nt ? ts.acc_bar : ts.afree_bar);
      __syncwarp();
    }
  }
  // ---- accumulator -> part[(segment * R + activation row) * RB + local weight row] ----
  // warps q and q + 4 may touch quarter q; they take its column pieces in turn.  A lane keeps the pieces of its own segment.
  {
    const int qd = warp & 3, half = warp >> 2;
    if (qd < NQ) {
      mbar_wait(ts.acc_bar, (uint32_t)(ts.nacc & 1));
      tc_fence_after();
      DBG(3);
      const int k_l = lane / R, i_l = lane - k_l * R;
      const bool valid = lane < spq;
      const uint32_t tq = ts.tmem + ((uint32_t)(32 * qd) << 16), part_s = smem_u32(part);
      int piece = 0;
      for (int k = 0; k < SG / NQ; ++k) {
        const int sg = qd + NQ * k;
        const bool keep = valid && k == k_l;
        const uint32_t pdst = part_s + (uint32_t)((sg * R + i_l) * g.RB) * 4u;
        for (int col = 0; col < g.RB; ++piece) {
          const int rem = g.RB - col;
          const bool my_piece = (piece & 1) == half;
          if (rem >= 32) {
            if (my_piece) {
              float v[32];
              tc_ld32(tq + (uint32_t)(sg * g.RB + col), v);
              if (keep) {
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(pdst + (uint32_t)(col + e) * 4u), "f"(v[e]), "f"(v[e + 1]), "f"(v[e + 2]), "f"(v[e + 3]) : "memory");
              }
            }
            col += 32;
          } else if (rem >= 16) {
            if (my_piece) {
              float v[16];
              tc_ld16(tq + (uint32_t)(sg * g.RB + col), v);
              if (keep) {
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(pdst + (uint32_t)(col + e) * 4u), "f"(v[e]), "f"(v[e + 1]), "f"(v[e + 2]), "f"(v[e + 3]) : "memory");
              }
            }
            col += 16;
          } else {
            if (my_piece) {
              float v[8];
              tc_ld8(tq + (uint32_t)(sg * g.RB + col), v);
              if (keep) {
#pragma unroll
                for (int e = 0; e < 8; e += 4)
                  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(pdst + (uint32_t)(col + e) * 4u), "f"(v[e]), "f"(v[e + 1]), "f"(v[e + 2]), "f"(v[e + 3]) : "memory");
              }
            }
            col += 8;
          }
        }
      }
      tc_fence_before();
    }
  }
  ++ts.nacc;
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
  DBG(4);

  if (e_on) {
    float v0 = 0.f, v1 = 0.f;                                  // segment partials summed in segment order
    for (int sg = 0; sg < SG; ++sg) { v0 += part[(sg * R + ei) * g.RB + lr0]; v1 += part[(sg * R + ei) * g.RB + lr1]; }
    if (EPI == EPI_RESID) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(pre_resid + rbf(v0), tag_out));
    } else if (EPI == EPI_STORE) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(v0, tag_out));
    } else if (EPI == EPI_SILU) {                            // same ops as gemv_epilogue<EPI_SILU>
      const float yv = rbf(v0), gt = rbf(v1);
      const float sg = rbf(gt / (1.0f + expf(-gt)));
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(__fmul_rn(yv, sg), tag_out));
    } else if (EPI == EPI_QKV) {
      const int qn_ = a.Hq * a.hd, kn_ = a.Hkv * a.hd;
      float o0 = rbf(v0), o1 = rbf(v1);
      if (en0 < qn_ + kn_) {                                 // same un-contracted fp32 ops as gemv_epilogue / _torch.py:57-68
        const float r0 = __fsub_rn(__fmul_rn(o0, pre_cs.x), __fmul_rn(o1, pre_cs.y));
        const float r1 = __fadd_rn(__fmul_rn(o1, pre_cs.x), __fmul_rn(o0, pre_cs.y));
        o0 = r0; o1 = r1;
      }
      if (en0 < qn_) {
        st_relaxed_u32(qt + (size_t)ei * qn_ + en0, tag_word(o0, tag_out));
        st_relaxed_u32(qt + (size_t)ei * qn_ + en1, tag_word(o1, tag_out));
      } else {
        const int kvsel = en0 < qn_ + kn_ ? 0 : 1;
        const int c0i = en0 - qn_ - kvsel * kn_, c1i = en1 - qn_ - kvsel * kn_;
        st_relaxed_u32(kvt + ((size_t)ei * 2 + kvsel) * kn_ + c0i, tag_word(o0, tag_out));
        st_relaxed_u32(kvt + ((size_t)ei * 2 + kvsel) * kn_ + c1i, tag_word(o1, tag_out));
        bf16* pb = a.kv_layer + ((size_t)pre_page * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;
        const int tk = pre_pos % ZB_PAGE_TOKENS;
        pb[((size_t)(c0i / a.hd) * ZB_PAGE_TOKENS + tk) * a.hd + (c0i % a.hd)] = f2bf(o0);
        pb[((size_t)(c1i / a.hd) * ZB_PAGE_TOKENS + tk) * a.hd + (c1i % a.hd)] = f2bf(o1);
      }
    } else {
      float u0 = 0.f;
      if (cfg) for (int sg = 0; sg < SG; ++sg) u0 += part[(sg * R + a.B + ei) * g.RB + lr0];
      gemv_epilogue<EPI>(a, ei, en0, en1, false, v0, 0.f, u0, 0.f);
    }
  }
  DBG(5);
#undef DBG
}

// K/V of the tokens cached by EARLIER steps for this CTA's first attention unit of the layer: issued before the
// in_proj phase so the tile is already in shared memory when the attention phase starts
struct MegaAttnMeta { int n_old, page, g, kv_len; };          // step constants of this CTA's first attention unit
__device__ __forceinline__ MegaAttnMeta mega_attention_meta(const MegaArgs& m, int unit, int nunits) {
  MegaAttnMeta t; t.n_old = 0; t.page = 0; t.g = 0; t.kv_len = 0;
  if (unit >= nunits) return t;
  const int split = unit % m.nsplit, r = unit / (m.nsplit * m.Hkv);
  t.g = (unit / m.nsplit) % m.Hkv;
  const int kv_len = m.lengths[r] + 1;
  t.kv_len = kv_len;
  t.n_old = max(0, min(kCH, kv_len - 1 - split * kCH));
  if (split * kCH < kv_len) t.page = m.page_table[(size_t)r * m.max_pages + split];
  return t;
}
__device__ __forceinline__ void mega_attention_prefetch(const MegaArgs& m, const bf16* kv_layer, const MegaAttnMeta& t, unsigned char* scratch) {
  if (t.n_old <= 0) { asm volatile("cp.async.commit_group;" ::: "memory"); return; }     // always one group: see cp.async.wait_group 1 in the in_proj norm
  bf16* ks = reinterpret_cast<bf16*>(scratch);
  bf16* vs = ks + kCH * kKStride;
  const bf16* kp = kv_layer + (((size_t)t.page * 2 + 0) * m.Hkv + t.g) * kCH * kHD;
  const bf16* vp = kv_layer + (((size_t)t.page * 2 + 1) * m.Hkv + t.g) * kCH * kHD;
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += kMW * 32) {
    const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
    if (tok < t.n_old) {
      cp_async16(ks + tok * kKStride + d8, kp + tok * kHD + d8);
      cp_async16(vs + tok * kHD + d8, vp + tok * kHD + d8);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void mega_attention_unit(const MegaArgs& m, const bf16* kv_layer, int unit, unsigned char* scratch, int warp, int lane,
                                                    const MegaAttnMeta* first, uint32_t tag_in, uint32_t tag_out, unsigned long long* stamp) {
  const bool prefetched = first != nullptr;                            // this CTA's first unit: its step constants are in *first
  bf16* ks = reinterpret_cast<bf16*>(scratch);                         // [64][136]
  bf16* vs = ks + kCH * kKStride;                                      // [64][128]
  float* qs = reinterpret_cast<float*>(vs + kCH * kHD);                // [8][128]
  float* ps = qs + 8 * kHD;                                            // [8][64]
  int* s_last = reinterpret_cast<int*>(ps + 8 * kCH);
  const int G = m.Hq / m.Hkv;
  const int split = unit % m.nsplit, g = (unit / m.nsplit) % m.Hkv, r = unit / (m.nsplit * m.Hkv);
  const int kv_len = first ? first->kv_len : m.lengths[r] + 1;
  const int nact = (kv_len + kCH - 1) / kCH;
  if (split >= nact) return;                                            // uniform for the CTA
  const int k0 = split * kCH, nk = min(kCH, kv_len - k0);
  const int page = first ? first->page : m.page_table[(size_t)r * m.max_pages + split];
  const bf16* kp = kv_layer + (((size_t)page * 2 + 0) * m.Hkv + g) * kCH * kHD;
  const bf16* vp = kv_layer + (((size_t)page * 2 + 1) * m.Hkv + g) * kCH * kHD;
  // tokens cached by earlier steps were prefetched (mega_attention_prefetch, before the in_proj phase) when
  // `prefetched`; this step's own token (index kv_len-1) is fetched now
  const int n_old = prefetched ? max(0, min(nk, kv_len - 1 - k0)) : 0;
  const int tok_new = kv_len - 1 - k0;                                  // this step's token, if it falls into this split
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += kMW * 32) {
    const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
    if (tok >= n_old && tok < nk && tok != tok_new) {
      cp_async16(ks + tok * kKStride + d8, kp + tok * kHD + d8);
      cp_async16(vs + tok * kHD + d8, vp + tok * kHD + d8);
    } else if (tok >= nk) {
      *reinterpret_cast<uint4*>(ks + tok * kKStride + d8) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vs + tok * kHD + d8) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int head = g * G + warp;
  if (warp < G) {                                                       // q of this step: tagged words from the in_proj phase
    const uint4 qv = poll_v4(m.qt + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4, tag_in);
    *reinterpret_cast<float4*>(&qs[warp * kHD + lane * 4]) = make_float4(untag(qv.x), untag(qv.y), untag(qv.z), untag(qv.w));
  }
  if (warp >= kMW - 2 && tok_new >= 0 && tok_new < nk) {                // K and V of this step's token: the last two warps
    const int kvsel = warp - (kMW - 2);
    const uint4 nv = poll_v4(m.kvt + ((size_t)r * 2 + kvsel) * m.Hkv * kHD + (size_t)g * kHD + lane * 4, tag_in);
    uint2 pk;
    pk.x = (nv.x >> 16) | (nv.y & 0xffff0000u);
    pk.y = (nv.z >> 16) | (nv.w & 0xffff0000u);
    bf16* dst = kvsel ? vs + tok_new * kHD : ks + tok_new * kKStride;
    *reinterpret_cast<uint2*>(dst + lane * 4) = pk;
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
  if (warp < G) {
    float sc[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tok = lane + 32 * j;
      float sacc = 0.f;
#pragma unroll
      for (int d8 = 0; d8 < kHD; d8 += 8) {
        const uint4 kv4 = *reinterpret_cast<const uint4*>(ks + tok * kKStride + d8);
        const float4 q0 = *reinterpret_cast<const float4*>(&qs[warp * kHD + d8]);
        const float4 q1 = *reinterpret_cast<const float4*>(&qs[warp * kHD + d8 + 4]);
        sacc = fmaf(bf16lo(kv4.x), q0.x, sacc); sacc = fmaf(bf16hi(kv4.x), q0.y, sacc);
        sacc = fmaf(bf16lo(kv4.y), q0.z, sacc); sacc = fmaf(bf16hi(kv4.y), q0.w, sacc);
        sacc = fmaf(bf16lo(kv4.z), q1.x, sacc); sacc = fmaf(bf16hi(kv4.z), q1.y, sacc);
        sacc = fmaf(bf16lo(kv4.w), q1.z, sacc); sacc = fmaf(bf16hi(kv4.w), q1.w, sacc);
      }
      sc[j] = (tok < nk) ? sacc * m.scale : -INFINITY;
    }
    const float mx = warp_max(fmaxf(sc[0], sc[1]));
    const float p0 = __expf(sc[0] - mx), p1 = __expf(sc[1] - mx);
    const float l = warp_sum(p0 + p1);
    ps[warp * kCH + lane] = p0; ps[warp * kCH + lane + 32] = p1;
    __syncwarp();
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tok = 0; tok < nk; ++tok) {
      const float pp = ps[warp * kCH + tok];
      const uint2 vv = *reinterpret_cast<const uint2*>(vs + tok * kHD + lane * 4);
      o[0] = fmaf(pp, bf16lo(vv.x), o[0]); o[1] = fmaf(pp, bf16hi(vv.x), o[1]);
      o[2] = fmaf(pp, bf16lo(vv.y), o[2]); o[3] = fmaf(pp, bf16hi(vv.y), o[3]);
    }
    float* part = m.attn_part + (((size_t)r * m.Hq + head) * m.nsplit + split) * kPart;
    *reinterpret_cast<float4*>(part + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    if (lane == 0) { part[kHD] = mx; part[kHD + 1] = l; }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
  if (threadIdx.x == 0) {
    // one acq_rel RMW publishes this CTA's partials (cumulative over the CTA barrier) and acquires the others'
    int32_t* cnt = m.attn_counters + (size_t)r * m.Hkv + g;
    int prev;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(prev) : "l"(cnt) : "memory");
    *s_last = (prev == nact - 1);
    if (*s_last) *cnt = 0;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
  if (*s_last && warp < G) {
    const float* base = m.attn_part + (((size_t)r * m.Hq + head) * m.nsplit) * kPart;
    // one pass with a running maximum (online softmax merge): the loads of all splits are independent, so the merge
    // costs one L2 round trip instead of two (maximum first, values second)
    float M = -INFINITY, L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int sp = 0; sp < nact; ++sp) {
      const float* ps_ = base + (size_t)sp * kPart;
      const float ms = __ldcg(ps_ + kHD), ls = __ldcg(ps_ + kHD + 1);
      const float4 ov = __ldcg(reinterpret_cast<const float4*>(ps_ + lane * 4));
      const float Mn = fmaxf(M, ms);
      const float so = __expf(M - Mn), sn = __expf(ms - Mn);   // M = -inf on the first split: so = 0
      L = fmaf(ls, sn, L * so);
      acc[0] = fmaf(ov.x, sn, acc[0] * so); acc[1] = fmaf(ov.y, sn, acc[1] * so);
      acc[2] = fmaf(ov.z, sn, acc[2] * so); acc[3] = fmaf(ov.w, sn, acc[3] * so);
      M = Mn;
    }
    const float inv = 1.0f / L;
    const uint4 outv = make_uint4(tag_word(acc[0] * inv, tag_out), tag_word(acc[1] * inv, tag_out), tag_word(acc[2] * inv, tag_out),
                                  tag_word(acc[3] * inv, tag_out));
    st_relaxed_v4(m.ayt + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4, outv);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");       // scratch free for the next unit
}

// One Mamba2 unit of the persistent kernel = (activation row r, head hh): causal-conv1d step over the head's 64 x channels and
// the shared 128 B + 128 C channels, selective state update of the head's [64 x 128] state (32 values per thread, in
// registers), y = C.h + D x.  Same arithmetic as mamba_scan_kernel with T = 1 (mamba_ssm causal_conv1d_update +
// selective_state_update, reached from zonos/backbone/_mamba_ssm.py:45-58).  Inputs: tagged words of the in_proj output
// row (z | xBC | dt); output: y as tagged bf16 words - the gate y * silu(z) is applied by the out_proj prologue (PRO_GATED),
// which keeps the product in fp32 like RMSNormGated does.
constexpr int kScanStageOff = 16 * 1024;                      // staging area of the prefetched state inside the phase scratch
// recurrent state of this CTA's first unit: cp.async into shared memory before the in_proj phase (it depends on earlier steps
// only), so the scan phase does not start with an HBM round trip.  Thread t stages exactly what thread t consumes.
__device__ __forceinline__ void mega_scan_prefetch(const MegaArgs& m, const MegaLayer& L, int unit, unsigned char* scratch) {
  constexpr int P = 64, N = 128, DC = 4, CH = P + 2 * N;
  if (unit < 0) { asm volatile("cp.async.commit_group;" ::: "memory"); return; }
  const int tid = threadIdx.x, r = unit / m.m_nheads, hh = unit % m.m_nheads;
  const int p = tid >> 2, quarter = tid & 3;
  unsigned char* st = scratch + kScanStageOff;
  const bf16* sp = L.ssm_state + (((size_t)r * m.m_nheads + hh) * P + p) * N + quarter * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) cp_async16(st + tid * 64 + j * 16, sp + j * 32);
  for (int c = tid; c < CH; c += kMW * 32) {
    const int gc = c < P ? hh * P + c : m.d_inner + (c - P);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(st + 16384 + c * 8)), "l"(L.conv_state + ((size_t)r * m.conv_dim + gc) * DC) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void mega_scan_unit(const MegaArgs& m, const MegaLayer& L, int unit, unsigned char* scratch, uint32_t tag_in, uint32_t tag_out,
                                               unsigned long long* stamp, bool prefetched) {
  constexpr int P = 64, N = 128, DC = 4, CH = P + 2 * N;
  float (*win)[CH] = reinterpret_cast<float (*)[CH]>(scratch);                 // [DC][CH] rolling window, oldest first (channel-contiguous)
  float (*cw)[CH] = reinterpret_cast<float (*)[CH]>(scratch + CH * DC * 4);
  float* cb = reinterpret_cast<float*>(scratch + 2 * CH * DC * 4);
  float* cout = cb + CH;
  const int tid = threadIdx.x, r = unit / m.m_nheads, hh = unit % m.m_nheads;
  const int p = tid >> 2, quarter = tid & 3;
  auto chan = [&](int c) { return c < P ? hh * P + c : m.d_inner + (c - P); };   // index into xBC / conv_state
  // state of earlier steps: staged by mega_scan_prefetch, or fetched now
  const unsigned char* st = scratch + kScanStageOff;
  if (prefetched) asm volatile("cp.async.wait_group 0;" ::: "memory");
  for (int c = tid; c < CH; c += kMW * 32) {
    const int gc = chan(c);
    const uint2 wv = prefetched ? *reinterpret_cast<const uint2*>(st + 16384 + c * 8)
                                : *reinterpret_cast<const uint2*>(L.conv_state + ((size_t)r * m.conv_dim + gc) * DC);
    const uint2 cv = __ldg(reinterpret_cast<const uint2*>(L.conv_w + (size_t)gc * DC));
    win[0][c] = bf16lo(wv.x); win[1][c] = bf16hi(wv.x); win[2][c] = bf16lo(wv.y); win[3][c] = bf16hi(wv.y);
    cw[0][c] = bf16lo(cv.x); cw[1][c] = bf16hi(cv.x); cw[2][c] = bf16lo(cv.y); cw[3][c] = bf16hi(cv.y);
    cb[c] = bf2f(L.conv_b[gc]);
  }
  float h[32];
  // state elements of this thread: n = 32 k + 8 quarter + e (see mamba_scan_kernel)
  bf16* sp = L.ssm_state + (((size_t)r * m.m_nheads + hh) * P + p) * N + quarter * 8;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const uint4 v = prefetched ? *reinterpret_cast<const uint4*>(st + tid * 64 + j * 2) : *reinterpret_cast<const uint4*>(sp + 4 * j);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { h[j + 2 * q] = bf16lo(w4[q]); h[j + 2 * q + 1] = bf16hi(w4[q]); }
  }
  const float A = -expf(bf2f(L.A_log[hh])), Dv = bf2f(L.Dp[hh]), dtb = bf2f(L.dt_bias[hh]);
  const uint32_t* row = m.zxt + (size_t)r * m.ipo;
  auto poll1 = [&](const uint32_t* q) {
    uint32_t w = ld_relaxed_u32(q);
    for (unsigned spins = 0; (w & 0xffffu) != tag_in; ++spins) {
      if (spins > kMegaSpinLimit) asm volatile("trap;");
      w = ld_relaxed_u32(q);
    }
    return untag(w);
  };
  for (int c = tid; c < CH; c += kMW * 32) {
    const float xin = poll1(row + m.d_inner + chan(c));
    float acc = cb[c];
#pragma unroll
    for (int j = 0; j < DC - 1; ++j) { win[j][c] = win[j + 1][c]; acc = fmaf(win[j][c], cw[j][c], acc); }
    win[DC - 1][c] = xin;
    acc = fmaf(xin, cw[DC - 1][c], acc);
    cout[c] = rbf(acc / (1.0f + expf(-acc)));                 // SiLU, output in the activation dtype
  }
  const float dtr = poll1(row + m.d_inner + m.conv_dim + hh) + dtb;
  const float zg = quarter == 0 ? poll1(row + hh * P + p) : 0.f;   // gate input of this thread's output element
  if (stamp && tid == 0) *stamp = gtime();
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
  const float dt = dtr > 20.0f ? dtr : log1pf(expf(dtr));      // softplus (F.softplus threshold 20)
  const float dA = expf(dt * A);
  const float xp = cout[p], dtx = dt * xp;
  float yacc = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int n = (j >> 3) * 32 + quarter * 8 + (j & 7);
    h[j] = fmaf(h[j], dA, dtx * cout[P + n]);
    yacc = fmaf(h[j], cout[P + N + n], yacc);
    h[j] = rbf(h[j]);                                          // decode: the stored state is in the cache dtype
  }
  yacc += __shfl_xor_sync(0xffffffffu, yacc, 1);
  yacc += __shfl_xor_sync(0xffffffffu, yacc, 2);
  if (quarter == 0) {
    // g = bf16(y) * silu(z) in fp32 (what RMSNormGated computes from the bf16 y and z), published as two tagged 16-bit halves
    const float y = rbf(yacc + Dv * xp);
    const uint32_t gb = __float_as_uint(y * (zg / (1.0f + expf(-zg))));
    uint32_t* dst = m.ygt + (size_t)r * 2 * m.d_inner + hh * P + p;
    st_relaxed_u32(dst, (gb & 0xffff0000u) | tag_out);
    st_relaxed_u32(dst + m.d_inner, (gb << 16) | tag_out);
  }
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 v;
    v.x = pack_bf16(h[j], h[j + 1]); v.y = pack_bf16(h[j + 2], h[j + 3]); v.z = pack_bf16(h[j + 4], h[j + 5]); v.w = pack_bf16(h[j + 6], h[j + 7]);
    *reinterpret_cast<uint4*>(sp + 4 * j) = v;
  }
  for (int c = tid; c < CH; c += kMW * 32) {
    if (c >= P && hh != 0) continue;                            // the shared B/C channels are written once (head 0)
    uint2 o;
    o.x = pack_bf16(win[0][c], win[1][c]); o.y = pack_bf16(win[2][c], win[3][c]);
    *reinterpret_cast<uint2*>(L.conv_state + ((size_t)r * m.conv_dim + chan(c)) * DC) = o;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");   // scratch free for the next unit / phase
}

__device__ __forceinline__ void mega_fill(GemvArgs& a, const MegaArgs& m, int R) {
  memset(&a, 0, sizeof(a));
  a.M = R; a.eps = m.eps; a.norm_kind = m.norm_kind; a.T = 1; a.Hq = m.Hq; a.Hkv = m.Hkv; a.hd = m.hd;
  a.rope_interleaved = m.rope_interleaved; a.rope = m.rope; a.rope_len = m.rope_len; a.lengths = m.lengths;
  a.page_table = m.page_table; a.max_pages = m.max_pages; a.F = m.F; a.B = m.B; a.cfg_scale = m.cfg_scale; a.logits = m.logits; a.QV = m.QV;
}

// MODE 0: FFMA2 consumer on the caller's bf16 weights (default); 1: tcgen05 consumer on the tile-ordered copy (ZB_MEGA_TC=1);
// 2: FP8 mode - the FFMA consumer on the e4m3 copy (ZB_FP8=1, transformer stacks only)
template <int R, int MODE>
__global__ void __launch_bounds__((kMW + 1) * 32, 1) decode_step_kernel(const __grid_constant__ MegaArgs m) {
  constexpr bool TC = MODE == 1, F8 = MODE == 2;
  constexpr int RM = (F8 && R <= 2) ? 2 : 1;                   // FP8 mode: rows of a ring stage relative to bf16 (see mega_produce)
  constexpr int kStage = kMegaStageBytes / (F8 ? 2 : 1) * RM, kWsz = F8 ? 1 : 2;
  extern __shared__ __align__(128) unsigned char smem_m[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tc_bar[2];
  __shared__ float red[2][kMW][4];
  __shared__ uint32_t tmem_slot;
  if (loop_idle(m.loop, m.T_delayed)) return;                 // same answer in every CTA: the loop state only changes in the sampler
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = m.ring_stages;
  // TC: the ring holds UMMA tiles (1024-byte aligned swizzle atoms), followed by the 32 KB activation image
  unsigned char* ring = TC ? reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_m) + 1023) & ~(uintptr_t)1023) : smem_m;
  unsigned char* abuf = ring + (size_t)S * kStage;
  float* part = reinterpret_cast<float*>(abuf + (TC ? kMegaTcImageBytes : 0));
  unsigned char* attn_scratch = reinterpret_cast<unsigned char*>(part) + m.part_bytes;
  bf16* nbuf = reinterpret_cast<bf16*>(attn_scratch + kMegaAttnBytes);   // [2 buffers][weight | bias][D]: norm parameters, copied a layer ahead
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TC ? 1 : kMW); }
    mbar_init(&tc_bar[0], 1); mbar_init(&tc_bar[1], 1);
    mbar_fence_init();
  }
  if (TC && warp == 0) {                                      // 256 accumulator columns (N <= 256 weight rows per CTA and matrix)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  __syncthreads();
  MegaTcState ts;
  ts.acc_bar = &tc_bar[0]; ts.afree_bar = &tc_bar[1]; ts.nacc = 0; ts.nafree = 0; ts.tmem = 0; ts.abuf = abuf;
  if (TC) { tc_fence_after(); ts.tmem = tmem_slot; }
  const int qn = m.Hq * m.hd, nqkv = (m.Hq + 2 * m.Hkv) * m.hd;
  GemvArgs a;

  if (warp == kMW) {
    // ===== producer: the whole step's weights, in consumption order =====
    const uint64_t pol = m.evict_first ? l2_evict_first_policy() : 0ull;
    int gst = 0;
    MegaLayer L = m.layers[0], Lnext = L;
    for (int li = 0; li < m.n_layer; ++li, L = Lnext) {
      if (li + 1 < m.n_layer) Lnext = m.layers[li + 1];       // the next layer's pointers are on their way while this one streams
      if (L.kind == 1) {                                      // Mamba2 layer: in_proj [ipo, D], out_proj [D, d_inner]
        mega_fill(a, m, R);
        a.W = L.in_proj; a.N = m.ipo; a.K = m.D;
        mega_produce<EPI_STORE>(a, ring, full_bar, empty_bar, S, gst, pol, lane);
        a.W = L.out_proj; a.N = m.D; a.K = m.d_inner;
        mega_produce<EPI_RESID>(a, ring, full_bar, empty_bar, S, gst, pol, lane);
        continue;
      }
      if (TC) {
        mega_produce_tc(L.in_t, m.tg[TG_QKV], m.D, ring, full_bar, empty_bar, S, gst, pol, lane);
        mega_produce_tc(L.out_t, m.tg[TG_OUT], qn, ring, full_bar, empty_bar, S, gst, pol, lane);
        mega_produce_tc(L.fc1_t, m.tg[TG_FC1], m.D, ring, full_bar, empty_bar, S, gst, pol, lane);
        mega_produce_tc(L.fc2_t, m.tg[TG_FC2], m.F, ring, full_bar, empty_bar, S, gst, pol, lane);
        continue;
      }
      mega_fill(a, m, R);
      a.W = L.in_proj; a.N = nqkv; a.K = m.D;
      mega_produce<EPI_QKV>(a, ring, full_bar, empty_bar, S, gst, pol, lane, kWsz, RM);
      a.W = L.out_proj; a.N = m.D; a.K = qn;
      mega_produce<EPI_STORE>(a, ring, full_bar, empty_bar, S, gst, pol, lane, kWsz, RM);
      a.W = L.fc1; a.N = 2 * m.F; a.K = m.D;
      mega_produce<EPI_SILU>(a, ring, full_bar, empty_bar, S, gst, pol, lane, kWsz, RM);
      a.W = L.fc2; a.N = m.D; a.K = m.F;
      mega_produce<EPI_RESID>(a, ring, full_bar, empty_bar, S, gst, pol, lane, kWsz, RM);
    }
    if (TC) {
      mega_produce_tc(m.heads_t, m.tg[TG_HEADS], m.D, ring, full_bar, empty_bar, S, gst, pol, lane);
      return;
    }
    mega_fill(a, m, R);
    a.W = m.heads; a.N = m.QV; a.K = m.D;
    mega_produce<EPI_HEADS>(a, ring, full_bar, empty_bar, S, gst, pol, lane, kWsz, RM);
    return;
  }

  // ===== consumers =====
  const unsigned epoch = m.sync[1];                         // written by this session's previous live step
  const int nph = m.nph;                                     // 2 + (4 + out_proj_repeats) per attention layer + 3 per Mamba2 layer
  int ph = 0, gst = 0, stamp_i = 0;
  const MegaAttnMeta ameta = mega_attention_meta(m, blockIdx.x, R * m.Hkv * m.nsplit);
  const bool stamping = m.timeline && blockIdx.x == 0;
  // norm parameters -> shared memory: one 16-byte cp.async per consumer thread and buffer half
  auto norm_prefetch = [&](int buf, const bf16* w, const bf16* b) {
    // TC: a CTA without a unit of the in_proj / fc1 matrix skips the phase that waits for this buffer, and two cp.async
    // to one address in flight at once may land in either order.  Free for the others: nothing is pending here.
    if (TC) asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int chunks = m.D / 8;
    bf16* dstw = nbuf + (size_t)buf * 2 * m.D;
    for (int q = threadIdx.x; q < 2 * chunks; q += kMW * 32) {
      const bf16* src = q < chunks ? w + q * 8 : (b ? b + (q - chunks) * 8 : nullptr);
      if (src) cp_async16(dstw + q * 8, src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  norm_prefetch(0, m.layers[0].norm_w, m.layers[0].norm_b);
  if (m.layers[0].kind == 1) norm_prefetch(1, m.layers[0].mnorm_w, m.layers[0].mnorm_w + m.D);
  else norm_prefetch(1, m.layers[0].norm2_w, m.layers[0].norm2_b);
  MegaQkvPre qkv_pre; qkv_pre.valid = false; qkv_pre.pos = 0; qkv_pre.page = 0; qkv_pre.cs = make_float2(1.f, 0.f);
  {
    mega_fill(a, m, R);
    a.N = nqkv; a.K = m.D;
    int u_begin, nrows;
    mega_slice<EPI_QKV>(a, u_begin, nrows);
    if (TC) { nrows = (int)blockIdx.x < m.tg[TG_QKV].nunits ? m.tg[TG_QKV].RB : 0; u_begin = (int)blockIdx.x * (m.tg[TG_QKV].RB / 2); }
    const int et = threadIdx.x;
    if (et < (nrows / 2) * R) {
      const int ej = et / R, ei = et % R;
      int en0, en1;
      unit_rows<EPI_QKV>(a, u_begin + ej, en0, en1);            // (TC: interleaved pairs only, en0 = 2 * unit)
      qkv_pre.pos = m.lengths[ei];
      if (en0 < qn + m.Hkv * m.hd) {
        const int ri = m.rope_interleaved ? (en0 % m.hd) / 2 : (en0 % m.hd);
        qkv_pre.cs = *reinterpret_cast<const float2*>(m.rope + ((size_t)min(qkv_pre.pos, m.rope_len - 1) * (m.hd / 2) + ri) * 2);
      }
      if (en0 >= qn) qkv_pre.page = m.page_table[(size_t)ei * m.max_pages + qkv_pre.pos / ZB_PAGE_TOKENS];
      qkv_pre.valid = true;
    }
  }
#define MEGA_STAMP() do { if (stamping && threadIdx.x == 0 && stamp_i < 126) m.timeline[stamp_i] = gtime(); ++stamp_i; } while (0)
#define MEGA_STAMP_SLOT() ((stamping && stamp_i < 126) ? &m.timeline[stamp_i++] : (++stamp_i, (unsigned long long*)nullptr))
#define TAG(p) mega_tag(epoch, nph, (p))
  if (m.steplog && blockIdx.x == 0 && threadIdx.x == 0 && m.loop) m.steplog[2 * min(m.loop->steps, 4000)] = gtime();
  MEGA_STAMP();
  // phase 0: codebook embedding sum (sequential bf16 adds, codec_utils.py:37) for this CTA's columns, both CFG rows
  {
    const int d_begin = (int)((long long)blockIdx.x * m.D / gridDim.x), d_end = (int)((long long)(blockIdx.x + 1) * m.D / gridDim.x);
    const long long col = m.loop ? (long long)m.loop->offset : 0;
    for (int t = threadIdx.x; t < (d_end - d_begin) * m.B; t += kMW * 32) {
      const int b = t / (d_end - d_begin), dd = d_begin + t % (d_end - d_begin);
      float acc = 0.f;
      for (int k = 0; k < m.Q; ++k) {
        long long id = m.delayed[((size_t)b * m.Q + k) * m.T_delayed + col];
        id = id < 0 ? 0 : (id >= m.vocab ? m.vocab - 1 : id);
        acc = rbf(acc + bf2f(m.emb[k][(size_t)id * m.D + dd]));
      }
      st_relaxed_u32(m.xt + (size_t)b * m.D + dd, tag_word(acc, TAG(0)));
      st_relaxed_u32(m.xt + (size_t)(m.B + b) * m.D + dd, tag_word(acc, TAG(0)));
    }
  }
  MEGA_STAMP();
  ph = 1;

  MegaLayer L = m.layers[0], Lnext = L;
  for (int li = 0; li < m.n_layer; ++li, L = Lnext) {
    if (li + 1 < m.n_layer) Lnext = m.layers[li + 1];         // pointers of the next layer: loaded a layer before they are needed
    // parameters of the next layer's two norm buffers: [0] its first norm (or the final norm), [1] norm2 of an attention
    // layer or the 2 D = d_inner gated-norm weights of a Mamba2 layer
    const bool lastl = li + 1 == m.n_layer;
    const bf16* nx0w = lastl ? m.normf_w : Lnext.norm_w;
    const bf16* nx0b = lastl ? m.normf_b : Lnext.norm_b;
    const bf16* nx1w = Lnext.kind == 1 ? Lnext.mnorm_w : Lnext.norm2_w;
    const bf16* nx1b = Lnext.kind == 1 ? Lnext.mnorm_w + m.D : Lnext.norm2_b;
    if (!TC && L.kind == 1) {
      // ---- Mamba2 layer: three phases instead of 4 + out_proj_repeats ----
      // M1: norm -> in_proj -> z | xBC | dt as tagged words
      mega_fill(a, m, R);
      a.W = L.in_proj; a.N = m.ipo; a.K = m.D; a.ldx = m.D; a.nw = nbuf; a.nb = L.norm_b ? nbuf + m.D : nullptr; a.ldy = m.ipo;
      const bool has_unit = (int)blockIdx.x < R * m.m_nheads;
      mega_scan_prefetch(m, L, has_unit ? (int)blockIdx.x : -1, attn_scratch);
      {
        unsigned long long* slot = MEGA_STAMP_SLOT();
        mega_consume<R, 2, 4, PRO_NORM, EPI_STORE>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), m.zxt, TAG(ph),
                                                   nullptr, nullptr, nullptr, slot, 1);
      }
      MEGA_STAMP(); ++ph;
      // M2: conv1d step + selective state update, one (row, head) unit per CTA
      {
        unsigned long long* slot = MEGA_STAMP_SLOT();
        for (int unit = blockIdx.x; unit < R * m.m_nheads; unit += gridDim.x)
          mega_scan_unit(m, L, unit, attn_scratch, TAG(ph - 1), TAG(ph), unit == (int)blockIdx.x ? slot : nullptr, unit == (int)blockIdx.x);
      }
      MEGA_STAMP(); ++ph;
      norm_prefetch(0, nx0w, nx0b);                            // buffer 0 is free since M1
      // M3: y * silu(z) -> gated RMSNorm -> out_proj + residual
      mega_fill(a, m, R);
      a.W = L.out_proj; a.N = m.D; a.K = m.d_inner; a.ldx = 2 * m.d_inner; a.ldy = m.D; a.ldr = m.D; a.nw = nbuf + 2 * m.D; a.nb = nullptr;
      {
        unsigned long long* slot = MEGA_STAMP_SLOT();
        mega_consume<R, 2, 4, PRO_GATED, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ygt, TAG(ph - 1), m.xt, TAG(ph),
                                                    m.xt, nullptr, nullptr, slot, 1, nullptr, nullptr, m.ygt + m.d_inner, 2 * m.d_inner, TAG(ph - 1));
      }
      MEGA_STAMP(); ++ph;
      if (!lastl) norm_prefetch(1, nx1w, nx1b);                // buffer 1 is free since the gated norm
      continue;
    }
    // A: norm -> in_proj -> RoPE -> KV append (+ q)
    mega_fill(a, m, R);
    a.W = L.in_proj; a.N = nqkv; a.K = m.D; a.ldx = m.D; a.nw = nbuf; a.nb = L.norm_b ? nbuf + m.D : nullptr; a.kv_layer = L.kv_layer;
    a.wscale = L.s_in;
    mega_attention_prefetch(m, L.kv_layer, ameta, attn_scratch);
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      if (TC)
        mega_consume_tc<R, 1, PRO_NORM, EPI_QKV>(a, m.tg[TG_QKV], ring, part, full_bar, empty_bar, ts, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1),
                                                 nullptr, TAG(ph), nullptr, 0u, m.qt, m.kvt, slot, 1, &qkv_pre, (stamping && li == 1) ? m.timeline + 200 : nullptr);
      else
      mega_consume<R, 2, 4 * RM, PRO_NORM, EPI_QKV, F8>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), nullptr, TAG(ph),
                                               nullptr, m.qt, m.kvt, slot, 1, &qkv_pre, (stamping && li == 1) ? m.timeline + 200 : nullptr);
    }
    const int ph_in = ph;                                      // x was last written by phase ph_in - 1 (embedding or the previous fc2)
    MEGA_STAMP(); ++ph;
    // B: attention over the paged cache
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      for (int unit = blockIdx.x; unit < R * m.Hkv * m.nsplit; unit += gridDim.x)
        mega_attention_unit(m, L.kv_layer, unit, attn_scratch, warp, lane, unit == (int)blockIdx.x ? &ameta : nullptr, TAG(ph - 1), TAG(ph),
                            unit == (int)blockIdx.x ? slot : nullptr);
    }
    MEGA_STAMP(); ++ph;
    // buffer 0 is free since the in_proj phase: the next layer's first norm (or the final norm) starts its way to shared memory
    norm_prefetch(0, nx0w, nx0b);
    // C/D: out_proj (twice in the reference); the slice stays in the ring between the passes
    {
      const int gst0 = gst;
      const uint32_t* src = m.ayt;
      for (int rep = 0; rep < m.out_proj_repeats; ++rep) {
        const bool last = rep == m.out_proj_repeats - 1;
        mega_fill(a, m, R);
        a.W = L.out_proj; a.N = m.D; a.K = qn; a.ldx = qn; a.ldy = m.D; a.ldr = m.D; a.wscale = L.s_out;
        int g2 = gst0;
        unsigned long long* slot = MEGA_STAMP_SLOT();
        if (last) {
          if (TC)
            mega_consume_tc<R, 1, PRO_NONE, EPI_RESID>(a, m.tg[TG_OUT], ring, part, full_bar, empty_bar, ts, red, S, g2, true, warp, lane, src,
                                                       TAG(ph - 1), m.xt, TAG(ph), m.xt, TAG(ph_in - 1), nullptr, nullptr, slot, 0, nullptr,
                                                       (stamping && li == 1) ? m.timeline + 208 : nullptr);
          else
          mega_consume<R, 2, 4 * RM, PRO_NONE, EPI_RESID, F8>(a, ring, part, full_bar, empty_bar, red, S, g2, true, warp, lane, src, TAG(ph - 1), m.xt, TAG(ph),
                                                     m.xt, nullptr, nullptr, slot);
        } else {
          uint32_t* dst = (src == m.y1t) ? m.ayt : m.y1t;
          if (TC)
            mega_consume_tc<R, 1, PRO_NONE, EPI_STORE>(a, m.tg[TG_OUT], ring, part, full_bar, empty_bar, ts, red, S, g2, false, warp, lane, src,
                                                       TAG(ph - 1), dst, TAG(ph), nullptr, 0u, nullptr, nullptr, slot, 0, nullptr);
          else
          mega_consume<R, 2, 4 * RM, PRO_NONE, EPI_STORE, F8>(a, ring, part, full_bar, empty_bar, red, S, g2, false, warp, lane, src, TAG(ph - 1), dst, TAG(ph),
                                                     nullptr, nullptr, nullptr, slot);
          src = dst;
        }
        gst = g2;
        MEGA_STAMP(); ++ph;
      }
    }
    // E: norm2 -> fc1 -> value * silu(gate)
    mega_fill(a, m, R);
    a.W = L.fc1; a.N = 2 * m.F; a.K = m.D; a.ldx = m.D; a.nw = nbuf + 2 * m.D; a.nb = L.norm2_b ? nbuf + 3 * m.D : nullptr; a.ldy = m.F;
    a.wscale = L.s_fc1;
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      if (TC)
        mega_consume_tc<R, 1, PRO_NORM, EPI_SILU>(a, m.tg[TG_FC1], ring, part, full_bar, empty_bar, ts, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1),
                                                  m.ht, TAG(ph), nullptr, 0u, nullptr, nullptr, slot, 0, nullptr, (stamping && li == 1) ? m.timeline + 216 : nullptr);
      else
      mega_consume<R, 2, 4 * RM, PRO_NORM, EPI_SILU, F8>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), m.ht, TAG(ph),
                                                nullptr, nullptr, nullptr, slot, 0, nullptr, (stamping && li == 1) ? m.timeline + 220 : nullptr);
    }
    MEGA_STAMP(); ++ph;
    if (!lastl) norm_prefetch(1, nx1w, nx1b);
    // F: fc2 + residual
    mega_fill(a, m, R);
    a.W = L.fc2; a.N = m.D; a.K = m.F; a.ldx = m.F; a.ldy = m.D; a.ldr = m.D; a.wscale = L.s_fc2;
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      if (TC) {
        // x was last written by the out_proj pass two phases back (phase ph - 2)
        if (m.F > 2048)
          mega_consume_tc<R, 4, PRO_NONE, EPI_RESID>(a, m.tg[TG_FC2], ring, part, full_bar, empty_bar, ts, red, S, gst, true, warp, lane, m.ht,
                                                     TAG(ph - 1), m.xt, TAG(ph), m.xt, TAG(ph - 2), nullptr, nullptr, slot, 0, nullptr,
                                                     (stamping && li == 1) ? m.timeline + 224 : nullptr);
        else
          mega_consume_tc<R, 1, PRO_NONE, EPI_RESID>(a, m.tg[TG_FC2], ring, part, full_bar, empty_bar, ts, red, S, gst, true, warp, lane, m.ht,
                                                     TAG(ph - 1), m.xt, TAG(ph), m.xt, TAG(ph - 2), nullptr, nullptr, slot, 0, nullptr);
      } else if (m.F == 8192)
        mega_consume<R, 4, 2 * RM, PRO_NONE, EPI_RESID, F8>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);
      else
        mega_consume<R, 2, 4 * RM, PRO_NONE, EPI_RESID, F8>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);
    }
    MEGA_STAMP(); ++ph;
  }
  // heads: final norm -> fused heads -> fp32 -> CFG mix
  mega_fill(a, m, R);
  a.W = m.heads; a.N = m.QV; a.K = m.D; a.ldx = m.D; a.nw = nbuf; a.nb = m.normf_b ? nbuf + m.D : nullptr; a.wscale = m.s_heads;
  {
    unsigned long long* slot = MEGA_STAMP_SLOT();
    if (TC)
      mega_consume_tc<R, 1, PRO_NORM, EPI_HEADS>(a, m.tg[TG_HEADS], ring, part, full_bar, empty_bar, ts, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1),
                                                 nullptr, 0u, nullptr, 0u, nullptr, nullptr, slot, 0, nullptr);
    else
    mega_consume<R, 2, 4 * RM, PRO_NORM, EPI_HEADS, F8>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), nullptr, 0u,
                                               nullptr, nullptr, nullptr, slot);
  }
  MEGA_STAMP();
  if (TC) {                                                   // every accumulator read is behind the last phase's CTA barrier
    asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
    if (warp == 0) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(ts.tmem), "r"(256u) : "memory");
    }
  }
  // CTA 0 can only get here after it consumed outputs of every CTA, i.e. after every CTA read the epoch
  if (blockIdx.x == 0 && threadIdx.x == 0) m.sync[1] = epoch + 1;
  if (m.steplog && blockIdx.x == 0 && threadIdx.x == 0 && m.loop) m.steplog[2 * min(m.loop->steps, 4000) + 1] = gtime();
#undef MEGA_STAMP
#undef MEGA_STAMP_SLOT
#undef TAG
}

// ------------------------------------------------------------------ plain norm ---------------
struct NormArgs { const bf16* x; int64_t ldx; bf16* y; int64_t ldy; const bf16* w; const bf16* b; int D; float eps; int kind; };
__global__ void __launch_bounds__(256) norm_kernel(NormArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* src = a.x + (size_t)blockIdx.x * a.ldx;
  bf16* dst = a.y + (size_t)blockIdx.x * a.ldy;
  float s = 0.f;
  for (int k = threadIdx.x; k < a.D; k += 256) s += bf2f(src[k]);
  s = warp_sum(s);
  if (lane == 0) red[warp][0] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w][0];
  const float mean = a.kind == ZB_NORM_LAYERNORM ? tot / (float)a.D : 0.f;
  float sq = 0.f;
  for (int k = threadIdx.x; k < a.D; k += 256) { float d = bf2f(src[k]) - mean; sq += d * d; }
  sq = warp_sum(sq);
  if (lane == 0) red[warp][1] = sq;
  __syncthreads();
  float tsq = 0.f;
  for (int w = 0; w < 8; ++w) tsq += red[w][1];
  const float rstd = rsqrtf(tsq / (float)a.D + a.eps);
  for (int k = threadIdx.x; k < a.D; k += 256)
    dst[k] = f2bf((bf2f(src[k]) - mean) * rstd * bf2f(a.w[k]) + (a.b ? bf2f(a.b[k]) : 0.f));
}


// ------------------------------------------------------------------ Mamba2: conv1d step + selective state update ---
// Replaces mamba_ssm's causal_conv1d_update / causal_conv1d_fn + selective_state_update / mamba_chunk_scan_combined
// (reached from zonos/backbone/_mamba_ssm.py:45-58; recurrence in SURVEY.md Appendix C) for T >= 1 tokens per row.
// grid (rows, nheads), 256 threads: thread (p = tid/4, quarter = tid%4) owns 32 of the 128 states of head dimension p
// in fp32 registers for the whole token loop, so the 16 KB-per-head SSM state is read and written exactly once per
// call; the depthwise conv window of the head's 64 x-channels and of the shared B/C channels lives in shared memory.
// Decode (T == 1) stores the state in the cache dtype every step like selective_state_update; prefill rounds only the
// final state like the chunked scan.  Output: g = bf16(y) * silu(z) in fp32 (input of the gated RMSNorm).
struct ScanArgs {
  const bf16* zx;           // [M, in_proj_out] rows z | xBC | dt
  int T, d_inner, d_state, d_conv, headdim, nheads, ngroups, in_proj_out, conv_dim;
  const bf16 *conv_w, *conv_b, *dt_bias, *A_log, *Dp;
  bf16* conv_state;         // [rows, conv_dim, d_conv]
  bf16* ssm_state;          // [rows, nheads, headdim, d_state]
  float* g;                 // [M, d_inner]
  const zb_loop_state* loop; int T_delayed;
};

__global__ void __launch_bounds__(256, 4) mamba_scan_kernel(ScanArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (loop_idle(a.loop, a.T_delayed)) return;
  constexpr int P = 64, N = 128, DC = 4;                       // headdim, d_state, d_conv (checked on the host)
  constexpr int CH = P + 2 * N;                                // channels this CTA convolves: its 64 x + B + C
  __shared__ float win[DC][CH];                                // rolling window, oldest first; channel-contiguous: no bank conflicts
  __shared__ float cw[DC][CH];
  __shared__ float cb[CH];
  __shared__ float cout[CH];
  const int r = blockIdx.x, hh = blockIdx.y, tid = threadIdx.x;
  const int p = tid >> 2, quarter = tid & 3;
  auto chan = [&](int c) { return c < P ? hh * P + c : a.d_inner + (c - P); };   // index into xBC / conv_state
  // A CTA is one short chain load -> conv -> state update -> store, and only 4 fit an SM (h[32] per thread): issue EVERY
  // independent load before the first use, 8 / 16 bytes at a time (decode at 64 rows was 78 us per launch against 21 us of bytes)
  const int c0 = tid, c1 = tid + 256;                            // this thread's conv channels (c1 only for tid < CH - 256)
  const bool has1 = c1 < CH;
  const int g0 = chan(c0), g1 = has1 ? chan(c1) : g0;
  // state elements of this thread: n = 32 k + 8 quarter + e (k < 4, e < 8): the four threads of a state row read 64 contiguous
  // bytes per load, and their shared-memory reads of B / C below fall into different banks
  bf16* sp = a.ssm_state + (((size_t)r * a.nheads + hh) * P + p) * N + quarter * 8;
  uint4 sv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) sv[j] = *reinterpret_cast<const uint4*>(sp + 32 * j);
  const uint2 ws0 = *reinterpret_cast<const uint2*>(a.conv_state + ((size_t)r * a.conv_dim + g0) * DC);
  const uint2 ws1 = has1 ? *reinterpret_cast<const uint2*>(a.conv_state + ((size_t)r * a.conv_dim + g1) * DC) : make_uint2(0, 0);
  const uint2 wc0 = __ldg(reinterpret_cast<const uint2*>(a.conv_w + (size_t)g0 * DC));
  const uint2 wc1 = __ldg(reinterpret_cast<const uint2*>(a.conv_w + (size_t)g1 * DC));
  const float b0 = bf2f(a.conv_b[g0]), b1 = bf2f(a.conv_b[g1]);
  const bf16* row0 = a.zx + (size_t)r * a.T * a.in_proj_out;
  float nx0 = bf2f(row0[a.d_inner + g0]), nx1 = has1 ? bf2f(row0[a.d_inner + g1]) : 0.f;
  float ndt = bf2f(row0[a.d_inner + a.conv_dim + hh]), nz = bf2f(row0[hh * P + p]);
  const float A = -expf(bf2f(a.A_log[hh])), Dv = bf2f(a.Dp[hh]), dtb = bf2f(a.dt_bias[hh]);
  win[0][c0] = bf16lo(ws0.x); win[1][c0] = bf16hi(ws0.x); win[2][c0] = bf16lo(ws0.y); win[3][c0] = bf16hi(ws0.y);
  cw[0][c0] = bf16lo(wc0.x); cw[1][c0] = bf16hi(wc0.x); cw[2][c0] = bf16lo(wc0.y); cw[3][c0] = bf16hi(wc0.y);
  cb[c0] = b0;
  if (has1) {
    win[0][c1] = bf16lo(ws1.x); win[1][c1] = bf16hi(ws1.x); win[2][c1] = bf16lo(ws1.y); win[3][c1] = bf16hi(ws1.y);
    cw[0][c1] = bf16lo(wc1.x); cw[1][c1] = bf16hi(wc1.x); cw[2][c1] = bf16lo(wc1.y); cw[3][c1] = bf16hi(wc1.y);
    cb[c1] = b1;
  }
  float h[32];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w4[4] = {sv[j].x, sv[j].y, sv[j].z, sv[j].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { h[8 * j + 2 * q] = bf16lo(w4[q]); h[8 * j + 2 * q + 1] = bf16hi(w4[q]); }
  }
  // (each thread reads back only the window / weights of its own channels: no barrier needed before the token loop)
  for (int t = 0; t < a.T; ++t) {
    const float x0 = nx0, x1 = nx1, dt_in = ndt, z_in = nz;
    if (t + 1 < a.T) {
      const bf16* nrow = a.zx + ((size_t)r * a.T + t + 1) * a.in_proj_out;
      nx0 = bf2f(nrow[a.d_inner + chan(c0)]);
      if (c1 < CH) nx1 = bf2f(nrow[a.d_inner + chan(c1)]);
      ndt = bf2f(nrow[a.d_inner + a.conv_dim + hh]); nz = bf2f(nrow[hh * P + p]);
    }
    for (int c = tid; c < CH; c += 256) {
      const float xin = c == c0 ? x0 : x1;
      float acc = cb[c];
#pragma unroll
      for (int j = 0; j < DC - 1; ++j) { win[j][c] = win[j + 1][c]; acc = fmaf(win[j][c], cw[j][c], acc); }
      win[DC - 1][c] = xin;
      acc = fmaf(xin, cw[DC - 1][c], acc);
      cout[c] = rbf(acc / (1.0f + expf(-acc)));               // SiLU, output in the activation dtype
    }
    __syncthreads();
    const float dtr = dt_in + dtb;
    const float dt = dtr > 20.0f ? dtr : log1pf(expf(dtr));      // softplus (F.softplus threshold 20)
    const float dA = expf(dt * A);
    const float xp = cout[p], dtx = dt * xp;
    float yacc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = (j >> 3) * 32 + quarter * 8 + (j & 7);
      h[j] = fmaf(h[j], dA, dtx * cout[P + n]);
      yacc = fmaf(h[j], cout[P + N + n], yacc);
      if (a.T == 1) h[j] = rbf(h[j]);                            // decode: the stored state is in the cache dtype
    }
    yacc += __shfl_xor_sync(0xffffffffu, yacc, 1);
    yacc += __shfl_xor_sync(0xffffffffu, yacc, 2);
    if (quarter == 0) {
      const float y = rbf(yacc + Dv * xp);
      const float z = z_in;
      a.g[((size_t)r * a.T + t) * a.d_inner + hh * P + p] = y * (z / (1.0f + expf(-z)));
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 v;
      v.x = pack_bf16(h[j], h[j + 1]); v.y = pack_bf16(h[j + 2], h[j + 3]); v.z = pack_bf16(h[j + 4], h[j + 5]); v.w = pack_bf16(h[j + 6], h[j + 7]);
    *reinterpret_cast<uint4*>(sp + 4 * j) = v;                  // elements 32 (j / 8) + 8 quarter ...
  }
  for (int c = tid; c < CH; c += 256) {
    if (c >= P && hh != 0) continue;                              // the shared B/C channels are written once (head 0)
    uint2 o;
    o.x = pack_bf16(win[0][c], win[1][c]); o.y = pack_bf16(win[2][c], win[3][c]);
    *reinterpret_cast<uint2*>(a.conv_state + ((size_t)r * a.conv_dim + chan(c)) * DC) = o;
  }
}

// gated RMSNorm of mamba_ssm (RMSNormGated, norm_before_gate=False): xn = bf16(g * rsqrt(mean(g^2) + eps) * w)
struct GNormArgs { const float* g; bf16* y; const bf16* w; int D; float eps; const zb_loop_state* loop; int T_delayed; };
__global__ void __launch_bounds__(256) gated_norm_kernel(GNormArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (loop_idle(a.loop, a.T_delayed)) return;
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = a.g + (size_t)blockIdx.x * a.D;
  float sq = 0.f;
  for (int k = threadIdx.x; k < a.D; k += 256) { const float v = src[k]; sq = fmaf(v, v, sq); }
  sq = warp_sum(sq);
  if (lane == 0) red[warp] = sq;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / (float)a.D + a.eps);
  for (int k = threadIdx.x; k < a.D; k += 256) a.y[(size_t)blockIdx.x * a.D + k] = f2bf(src[k] * rstd * bf2f(a.w[k]));
}

// ------------------------------------------------------------------ host side ----------------
// ---- tile-ordered weight copy for the tcgen05 consumer of the persistent kernel (MegaTcGeo) -------------------------
// FP8 mode: one weight row per CTA -> e4m3 bytes + a power-of-two scale with |w| / scale in [1, 2) for the row's largest
// element (so q * scale is exactly a bf16 number, and e4m3's 3 mantissa bits are all the rounding there is)
__global__ void __launch_bounds__(256) quant_e4m3_kernel(const bf16* W, unsigned char* q, float* scale, int K) {
  __shared__ float red[8];
  const bf16* w = W + (size_t)blockIdx.x * K;
  float amax = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) amax = fmaxf(amax, fabsf(bf2f(w[k])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) amax = fmaxf(amax, red[i]);
  const float sc = (amax > 0.f && amax < 3.0e38f) ? ldexpf(1.0f, ilogbf(amax)) : 1.0f;
  unsigned char* dst = q + (size_t)blockIdx.x * K;
  // |w| / sc < 2; a value that would round UP to 2.0 is held at 1.875 (the largest e4m3 number below 2), so that quantising the
  // dequantised row finds the same scale and reproduces the same bytes (the quantiser is idempotent)
  for (int k = threadIdx.x; k < K; k += 256)
    dst[k] = (unsigned char)__nv_cvt_float_to_fp8(fminf(fmaxf(bf2f(w[k]) / sc, -1.875f), 1.875f), __NV_SATFINITE, __NV_E4M3);
  if (threadIdx.x == 0) scale[blockIdx.x] = sc;
}

struct PretileArgs { const bf16* W; uint4* out; int N, K, RB, RBv, F, nunits, S; };
// one thread per 16-byte chunk of the output; physical chunk c of tile row n holds logical chunk c ^ (n % 8) (128-byte swizzle)
__global__ void __launch_bounds__(256) pretile_kernel(PretileArgs a) {
  const int NB = a.S * a.RB, Ks = a.K / a.S;                 // tile rows, k per segment
  const size_t nt = (size_t)Ks / 64, total = (size_t)a.nunits * nt * NB * 8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % 8), n = (int)((idx / 8) % NB);
    const size_t t = (idx / (8 * (size_t)NB)) % nt, u = idx / (8 * (size_t)NB * nt);
    const int sg = n / a.RB, r = n % a.RB;
    const size_t kb = ((size_t)sg * Ks) / 64 + t;
    long long row; bool valid;
    if (a.RBv) {                                              // fc1: value rows, then the gate rows of the same features
      const long long f = (long long)u * a.RBv + (r % a.RBv);
      valid = f < a.F; row = (r < a.RBv ? 0 : a.F) + f;
    } else {
      row = (long long)u * a.RB + r; valid = row < a.N;
    }
    const int j = c ^ (n & 7);
    a.out[idx] = valid ? __ldg(reinterpret_cast<const uint4*>(a.W + (size_t)row * a.K + kb * 64 + (size_t)j * 8)) : make_uint4(0, 0, 0, 0);
  }
}

inline int env_int(const char* name, int dflt);
struct MegaTcPlan { MegaTcGeo g[TG_COUNT]; int K[TG_COUNT]; size_t off[TG_COUNT], layer_bytes, total; };
bool mega_tc_plan(const zb_model_desc& d, int grid, int R, MegaTcPlan* p) {
  memset(p, 0, sizeof(*p));
  const int qn = d.n_heads * d.head_dim, nqkv = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim, QV = d.n_codebooks * d.head_vocab;
  auto unit = [&](int N) { return ((N + grid - 1) / grid + 7) / 8 * 8; };
  auto geo = [&](int kind, int N, int K, bool gated) -> bool {
    MegaTcGeo& g = p->g[kind];
    if (gated) { g.RBv = unit(N / 2); g.RB = 2 * g.RBv; g.nunits = (N / 2 + g.RBv - 1) / g.RBv; }
    else { g.RBv = 0; g.RB = unit(N); g.nunits = (N + g.RB - 1) / g.RB; }
    p->K[kind] = K;
    if (K % 64 || K > 8192 || g.RB > 256 || g.nunits > grid) return false;
    // segments: as many as fit the MMA (S * RB <= 256 columns, 4 rows x S <= 64 accumulator rows) and divide K into 64-k tiles
    g.S = 1;
    const int smax = env_int("ZB_MEGA_TC_SMAX", 16);          // (debug: 1 = plain row slices)
    while (g.S < 16 && 2 * g.S <= smax && 2 * g.S * g.RB <= 256 && K % (2 * g.S * 64) == 0) g.S *= 2;
    g.kps = kMegaStageBytes / (g.S * g.RB * 128);
    const int items = (kind == TG_QKV ? g.RB / 2 : gated ? g.RBv : g.RB) * R;      // one epilogue item per consumer thread
    return g.kps >= 1 && items <= kMW * 32;
  };
  if (!geo(TG_QKV, nqkv, d.d_model, false) || !geo(TG_OUT, d.d_model, qn, false) || !geo(TG_FC1, 2 * d.d_ff, d.d_model, true) ||
      !geo(TG_FC2, d.d_model, d.d_ff, false) || !geo(TG_HEADS, QV, d.d_model, false))
    return false;
  if (d.d_model > 2048 || qn > 2048 || !d.rope_interleaved) return false;       // one 2048-k image per phase except fc2; pairs (2i, 2i+1)
  size_t off = 0;
  for (int k = 0; k < TG_HEADS; ++k) { p->off[k] = off; off += (size_t)p->g[k].nunits * p->g[k].RB * p->K[k] * 2; }
  p->layer_bytes = off;
  p->off[TG_HEADS] = (size_t)d.n_layer * off;
  p->total = p->off[TG_HEADS] + (size_t)p->g[TG_HEADS].nunits * p->g[TG_HEADS].RB * p->K[TG_HEADS] * 2;
  return true;
}

struct Scratch {
  bf16 *q, *attn_y, *y1, *h, *xn, *zx, *gn; float *part, *g; int32_t* counters; int nsplit;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Scratch carve(const zb_model* mdl, void* base, int M, int nsplit, size_t* total) {
  const zb_model_desc& d = mdl->d;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t qn = (size_t)d.n_heads * d.head_dim;
  size_t o_q = take((size_t)M * qn * 2), o_ay = take((size_t)M * qn * 2), o_y1 = take((size_t)M * d.d_model * 2);
  size_t o_h = take((size_t)M * d.d_ff * 2);
  size_t o_xn = take((size_t)M * d.d_model * 2);
  const size_t ipo = (size_t)2 * d.d_inner + 2 * d.m_ngroups * d.d_state + (d.m_headdim ? d.d_inner / d.m_headdim : 0);
  size_t o_zx = take((size_t)M * ipo * 2), o_g = take((size_t)M * d.d_inner * 4), o_gn = take((size_t)M * d.d_inner * 2);
  size_t o_part = take((size_t)M * d.n_heads * nsplit * kPart * 4);
  if (total) *total = off;
  Scratch s{};
  if (base) {
    char* b = (char*)base;
    s.q = (bf16*)(b + o_q); s.attn_y = (bf16*)(b + o_ay); s.y1 = (bf16*)(b + o_y1); s.h = (bf16*)(b + o_h); s.xn = (bf16*)(b + o_xn); s.zx = (bf16*)(b + o_zx); s.g = (float*)(b + o_g); s.gn = (bf16*)(b + o_gn);
    s.part = (float*)(b + o_part);
  }
  s.nsplit = nsplit;
  return s;
}

template <int MT, int PRO, int EPI>
zb_status launch_gemv_t(zb_ctx* ctx, GemvArgs& a, int mtiles, cudaStream_t stream) {
  const size_t smem = (size_t)MT * a.K * 2;
  if (smem > 48 * 1024) ZB_CUDA(ctx, zb_ensure_smem(ctx, gemv_kernel<MT, PRO, EPI>, smem));
  const int npairs = (EPI == EPI_SILU) ? a.F : (a.N + 1) / 2;
  int gx = (npairs + kWarps - 1) / kWarps;
  const int cap = ctx->num_sms * 2;                 // two CTAs of 8 warps per SM keep >32 KB of loads in flight
  if (gx > cap) gx = cap;
  dim3 grid(gx, mtiles);
  ZB_CUDA(ctx, zb_launch_pdl(gemv_kernel<MT, PRO, EPI>, grid, dim3(kThreads), smem, stream, a));
  ctx->launches++;
  return ZB_OK;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int R, int NC, int RW, int PRO, int EPI>
zb_status launch_gemv3_t(zb_ctx* ctx, GemvArgs& a, cudaStream_t stream) {
  const int nunits = (EPI == EPI_SILU) ? a.F : (EPI == EPI_QKV) ? a.N / 2 : a.N;
  const int grid = nunits < ctx->num_sms ? nunits : ctx->num_sms;
  const int KS = a.K / (NC * 256);
  ZB_REQUIRE(ctx, KS >= 1 && KS <= kW3 && kW3 % KS == 0, "gemv3: K=%d does not fit the warp layout", a.K);
  const int RPS = (kW3 / KS) * RW, stage_bytes = RPS * a.K * 2;
  static const int ring_kb = env_int("ZB_GEMV_RING_KB", 96);      // two CTAs per SM: this kernel + its PDL successor
  static const int ahead = env_int("ZB_GEMV_L2_AHEAD", 0);
  int stages = ring_kb * 1024 / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > kMaxStages) stages = kMaxStages;
  a.ring_stages = stages; a.prefetch_ahead = ahead;
  const int nrows_max = ((nunits + grid - 1) / grid) * ((EPI == EPI_SILU || EPI == EPI_QKV) ? 2 : 1);
  const int nrows_pad = (nrows_max + RPS - 1) / RPS * RPS;
  const size_t smem = (size_t)stages * stage_bytes + (size_t)nrows_pad * KS * R * sizeof(float);
  ZB_REQUIRE(ctx, smem <= 220 * 1024, "gemv3: %zu bytes of shared memory", smem);
  ZB_CUDA(ctx, zb_ensure_smem(ctx, gemv3_kernel<R, NC, RW, PRO, EPI>, smem));
  ZB_CUDA(ctx, zb_launch_pdl(gemv3_kernel<R, NC, RW, PRO, EPI>, dim3(grid), dim3((kW3 + 1) * 32), smem, stream, a));
  ctx->launches++;
  return ZB_OK;
}

// warp layout per K (16 consumer warps = KS k-slices x RG row groups; a stage holds RPS = RG*RW rows):
//   K=8192: NC=2 KS=16 RG=1 | K=4096: NC=1 KS=16 RG=1 | K=2048: NC=1 KS=8 RG=2
//   K=1024: NC=1 KS=4 RG=4  | K=512:  NC=1 KS=2  RG=8 | K=256:  NC=1 KS=1 RG=16
// RW (rows per warp per stage) sets the stage size: K=2048: RW 4/2/1 -> 32/16/8 KB; K=8192: RW 2/1 -> 32/16 KB.
// Measured on B200 (scripts/sweep_gemv.sh): 32 KB stages are fastest (per-stage barrier + reduction cost dominates).
template <int R, int PRO, int EPI>
zb_status launch_gemv3_r(zb_ctx* ctx, GemvArgs& a, cudaStream_t stream) {
  static const int stage_kb = env_int("ZB_GEMV_STAGE_KB", 32);
  if (a.K == 8192) {
    if (stage_kb <= 16) return launch_gemv3_t<R, 2, 1, PRO, EPI>(ctx, a, stream);
    return launch_gemv3_t<R, 2, 2, PRO, EPI>(ctx, a, stream);
  }
  if (stage_kb <= 8) return launch_gemv3_t<R, 1, 1, PRO, EPI>(ctx, a, stream);
  if (stage_kb <= 16) return launch_gemv3_t<R, 1, 2, PRO, EPI>(ctx, a, stream);
  return launch_gemv3_t<R, 1, 4, PRO, EPI>(ctx, a, stream);
}

inline bool gemv3_ok(int K, int N, int epi) {
  if (K != 256 && K != 512 && K != 1024 && K != 2048 && K != 4096 && K != 8192) return false;
  if (epi == EPI_QKV && (N & 1)) return false;
  return true;
}

template <int PRO, int EPI>
zb_status launch_gemv(zb_ctx* ctx, GemvArgs& a, cudaStream_t stream) {
  // tile rows: HEADS pairs cond/uncond rows inside one tile, so tiles hold MT/2 utterances
  const bool pairs = (EPI == EPI_HEADS && a.cfg_scale != 1.0f);
  const int rows = pairs ? 2 * a.B : a.M;
  ZB_REQUIRE(ctx, a.K % 8 == 0, "gemv: K=%d must be a multiple of 8", a.K);
  if (rows <= 4 && gemv3_ok(a.K, a.N, EPI) && (!pairs || rows == a.M)) {       // decode: whole batch in registers
    if (rows <= 2) return launch_gemv3_r<2, PRO, EPI>(ctx, a, stream);
    return launch_gemv3_r<4, PRO, EPI>(ctx, a, stream);
  }
  const int mt = kMaxMT;
  const int tiles = pairs ? (a.B + mt / 2 - 1) / (mt / 2) : (a.M + mt - 1) / mt;
  return launch_gemv_t<kMaxMT, PRO, EPI>(ctx, a, tiles, stream);
}

}  // namespace

size_t zb_backbone_scratch_bytes(const zb_model* model, int R, int T, int max_kv_len) {
  size_t total = 0;
  const int nsplit = (max_kv_len + kCH - 1) / kCH;
  carve(model, nullptr, R * T, nsplit, &total);
  return total;
}

zb_status zb_launch_embed(zb_ctx* ctx, const zb_embed_launch& L, cudaStream_t stream) {
  const zb_model_desc& d = L.model->d;
  ZB_REQUIRE(ctx, d.n_codebooks <= 16 && d.d_model % 8 == 0, "embed: unsupported dims");
  EmbedArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < d.n_codebooks; ++k) a.tab[k] = (const bf16*)L.model->emb[k];
  a.codes = L.codes; a.sb = L.sb; a.sq = L.sq; a.st = L.st; a.B = L.B; a.T = L.T; a.Q = d.n_codebooks; a.D = d.d_model;
  a.vocab = d.emb_vocab; a.repeat = L.repeat; a.out = L.out; a.out_rs = L.out_rs; a.loop = L.loop; a.T_delayed = L.T_delayed;
  ZB_CUDA(ctx, zb_launch_pdl(embed_kernel, dim3(L.B * L.T), dim3(256), 0, stream, a));
  ctx->launches++;
  return ZB_OK;
}

zb_status zb_run_layers(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, bf16* x, int R, int T,
                        int max_kv_len, const zb_loop_state* loop, int T_delayed, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  const int M = R * T;
  ZB_REQUIRE(ctx, d.head_dim == kHD, "head_dim %d unsupported (128 only)", d.head_dim);
  ZB_REQUIRE(ctx, d.n_heads_kv > 0 && d.n_heads % d.n_heads_kv == 0 && d.n_heads / d.n_heads_kv <= 8, "GQA group size unsupported");
  ZB_REQUIRE(ctx, cache && cache->rows >= R, "cache has %d rows, need %d", cache ? cache->rows : 0, R);
  const int nsplit = (max_kv_len + kCH - 1) / kCH;
  ZB_REQUIRE(ctx, nsplit <= cache->max_pages_per_row, "sequence of %d tokens exceeds the page table (%d pages)", max_kv_len,
             cache->max_pages_per_row);
  size_t need = 0;
  carve(model, nullptr, M, nsplit, &need);
  ZB_REQUIRE(ctx, ctx->scratch_bytes >= need, "internal: scratch not reserved (%zu < %zu)", ctx->scratch_bytes, need);
  Scratch s = carve(model, ctx->scratch, M, nsplit, nullptr);
  ZB_REQUIRE(ctx, (size_t)M * d.n_heads_kv <= ZB_NUM_COUNTERS, "too many rows (%d) for the attention merge counters", M);
  s.counters = ctx->counters;     // always zero between launches (the merging CTA resets its word)
  const size_t page_elems = (size_t)2 * d.n_heads_kv * ZB_PAGE_TOKENS * d.head_dim;
  const int qn = d.n_heads * d.head_dim;
  const int G = d.n_heads / d.n_heads_kv;

  const bool tc = M > 4 && d.d_model % 64 == 0 && d.d_ff % 64 == 0;   // dense path: tcgen05 GEMMs
  const bool tc_qkv = tc && (d.rope_interleaved || d.head_dim == 128);   // (rotate-half pairs need one head per 128-row tile)
  auto norm_rows = [&](const bf16* w, const bf16* b) -> zb_status {
    NormArgs na;
    na.x = x; na.ldx = d.d_model; na.y = s.xn; na.ldy = d.d_model; na.w = w; na.b = b; na.D = d.d_model; na.eps = d.norm_eps; na.kind = d.norm_kind;
    ZB_CUDA(ctx, zb_launch_pdl(norm_kernel, dim3(M), dim3(256), 0, stream, na));
    ctx->launches++;
    return ZB_OK;
  };
  for (int li = 0; li < d.n_layer; ++li) {
    const zb_layer& L = model->layers[li];
    GemvArgs a;
    if (L.kind == ZB_LAYER_MAMBA2) {
      // ---- Mamba2 layer: norm -> in_proj -> conv1d step + selective state update -> gated RMSNorm -> out_proj + residual
      const int nheads = d.d_inner / d.m_headdim, conv_dim = d.d_inner + 2 * d.m_ngroups * d.d_state;
      const int ipo = 2 * d.d_inner + 2 * d.m_ngroups * d.d_state + nheads;
      ZB_REQUIRE(ctx, d.m_headdim == 64 && d.d_state == 128 && d.d_conv == 4 && d.m_ngroups == 1, "Mamba2: only headdim 64, d_state 128, d_conv 4, ngroups 1");
      ZB_REQUIRE(ctx, cache->conv_state && cache->ssm_state, "Mamba2 layer %d: the cache has no conv/ssm state", li);
      const int mi = model->mamba_index[li];
      if (tc) {
        if (zb_status st = norm_rows((const bf16*)L.norm_w, (const bf16*)L.norm_b)) return st;
        zb_gemm_tc g;
        g.decode = (T == 1);
        g.W = (const bf16*)L.in_proj; g.x = s.xn; g.ldx = d.d_model; g.M = M; g.N = ipo; g.K = d.d_model; g.epi = 0; g.y = s.zx; g.ldy = ipo;
        if (zb_status st = zb_launch_gemm_tc(ctx, g, stream)) return st;
      } else {
        memset(&a, 0, sizeof(a));
        a.W = (const bf16*)L.in_proj; a.x = x; a.ldx = d.d_model; a.M = M; a.N = ipo; a.K = d.d_model;
        a.nw = (const bf16*)L.norm_w; a.nb = (const bf16*)L.norm_b; a.eps = d.norm_eps; a.norm_kind = d.norm_kind;
        a.y = s.zx; a.ldy = ipo; a.loop = loop; a.T_delayed = T_delayed;
        if (zb_status st = launch_gemv<PRO_NORM, EPI_STORE>(ctx, a, stream)) return st;
      }
      ScanArgs sa;
      memset(&sa, 0, sizeof(sa));
      sa.zx = s.zx; sa.T = T; sa.d_inner = d.d_inner; sa.d_state = d.d_state; sa.d_conv = d.d_conv; sa.headdim = d.m_headdim; sa.nheads = nheads;
      sa.ngroups = d.m_ngroups; sa.in_proj_out = ipo; sa.conv_dim = conv_dim;
      sa.conv_w = (const bf16*)L.conv_w; sa.conv_b = (const bf16*)L.conv_b; sa.dt_bias = (const bf16*)L.dt_bias; sa.A_log = (const bf16*)L.A_log; sa.Dp = (const bf16*)L.D;
      sa.conv_state = (bf16*)cache->conv_state + (size_t)mi * cache->rows * conv_dim * d.d_conv;
      sa.ssm_state = (bf16*)cache->ssm_state + (size_t)mi * cache->rows * nheads * d.m_headdim * d.d_state;
      sa.g = s.g; sa.loop = loop; sa.T_delayed = T_delayed;
      ZB_CUDA(ctx, zb_launch_pdl(mamba_scan_kernel, dim3(R, nheads), dim3(256), 0, stream, sa));
      ctx->launches++;
      GNormArgs ga;
      ga.g = s.g; ga.y = s.gn; ga.w = (const bf16*)L.mnorm_w; ga.D = d.d_inner; ga.eps = 1e-5f; ga.loop = loop; ga.T_delayed = T_delayed;
      ZB_CUDA(ctx, zb_launch_pdl(gated_norm_kernel, dim3(M), dim3(256), 0, stream, ga));
      ctx->launches++;
      if (tc && d.d_inner % 64 == 0) {
        zb_gemm_tc g;
        g.decode = (T == 1);
        g.W = (const bf16*)L.out_proj; g.x = s.gn; g.ldx = d.d_inner; g.M = M; g.N = d.d_model; g.K = d.d_inner;
        g.epi = 1; g.y = x; g.ldy = d.d_model; g.resid = x; g.ldr = d.d_model;
        if (zb_status st = zb_launch_gemm_tc(ctx, g, stream)) return st;
      } else {
        memset(&a, 0, sizeof(a));
        a.W = (const bf16*)L.out_proj; a.x = s.gn; a.ldx = d.d_inner; a.M = M; a.N = d.d_model; a.K = d.d_inner;
        a.y = x; a.ldy = d.d_model; a.resid = x; a.ldr = d.d_model; a.loop = loop; a.T_delayed = T_delayed;
        if (zb_status st = launch_gemv<PRO_NONE, EPI_RESID>(ctx, a, stream)) return st;
      }
      continue;
    }
    bf16* kv_layer = (bf16*)cache->kv_pages + (size_t)model->attn_index[li] * cache->num_pages * page_elems;
    // 1. norm -> in_proj -> RoPE -> KV append (+ q)
    if (tc_qkv) {
      if (zb_status st = norm_rows((const bf16*)L.norm_w, (const bf16*)L.norm_b)) return st;
      zb_gemm_tc g;
      g.decode = (T == 1);
      g.W = (const bf16*)L.in_proj; g.x = s.xn; g.ldx = d.d_model; g.M = M; g.N = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim; g.K = d.d_model;
      g.epi = 2; g.T = T; g.Hq = d.n_heads; g.Hkv = d.n_heads_kv; g.hd = d.head_dim; g.rope_interleaved = d.rope_interleaved;
      g.rope = d.rope_table; g.rope_len = d.rope_len; g.lengths = cache->lengths; g.page_table = cache->page_table;
      g.max_pages = cache->max_pages_per_row; g.kv_layer = kv_layer; g.q_out = s.q;
      if (zb_status st = zb_launch_gemm_tc(ctx, g, stream)) return st;
    } else {
      memset(&a, 0, sizeof(a));
      a.W = (const bf16*)L.in_proj; a.x = x; a.ldx = d.d_model; a.M = M; a.N = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim; a.K = d.d_model;
      a.nw = (const bf16*)L.norm_w; a.nb = (const bf16*)L.norm_b; a.eps = d.norm_eps; a.norm_kind = d.norm_kind;
      a.T = T; a.Hq = d.n_heads; a.Hkv = d.n_heads_kv; a.hd = d.head_dim; a.rope_interleaved = d.rope_interleaved;
      a.rope = d.rope_table; a.rope_len = d.rope_len; a.lengths = cache->lengths; a.page_table = cache->page_table;
      a.max_pages = cache->max_pages_per_row; a.num_pages = cache->num_pages; a.kv_layer = kv_layer; a.q_out = s.q;
      a.loop = loop; a.T_delayed = T_delayed;
      if (zb_status st = launch_gemv<PRO_NORM, EPI_QKV>(ctx, a, stream)) return st;
    }
    // 2. attention over the paged cache
    static const int pf_mma = env_int("ZB_PREFILL_MMA", 1);
    if (T > 1 && pf_mma && zb_attn_prefill_supported(d)) {
      if (zb_status st = zb_launch_attn_prefill(ctx, d, cache, s.q, kv_layer, s.attn_y, R, T, stream)) return st;
    } else {
      AttnArgs at;
      memset(&at, 0, sizeof(at));
      at.q = s.q; at.kv_layer = kv_layer; at.lengths = cache->lengths; at.page_table = cache->page_table;
      at.max_pages = cache->max_pages_per_row; at.T = T; at.Hq = d.n_heads; at.Hkv = d.n_heads_kv; at.nsplit = nsplit;
      at.scale = 1.0f / sqrtf((float)d.head_dim); at.part = s.part; at.counters = s.counters; at.y = s.attn_y;
      at.loop = loop; at.T_delayed = T_delayed;
      dim3 grid(M, d.n_heads_kv, nsplit);
      ZB_CUDA(ctx, zb_launch_pdl(attn_kernel, grid, dim3(32 * G), 0, stream, at));
      ctx->launches++;
    }
    // 3. out_proj (x repeats), the last one adds the residual
    const bf16* src = s.attn_y;
    for (int rep = 0; rep < d.out_proj_repeats; ++rep) {
      const bool last = rep == d.out_proj_repeats - 1;
      bf16* dst = last ? x : ((src == s.y1) ? s.attn_y : s.y1);
      if (!last) ZB_REQUIRE(ctx, qn == d.d_model, "out_proj_repeats > 1 needs H*hd == D");
      if (tc) {
        zb_gemm_tc g;
        g.decode = (T == 1);
        g.W = (const bf16*)L.out_proj; g.x = src; g.ldx = qn; g.M = M; g.N = d.d_model; g.K = qn;
        g.epi = last ? 1 : 0; g.y = dst; g.ldy = d.d_model; g.resid = x; g.ldr = d.d_model;
        if (zb_status st = zb_launch_gemm_tc(ctx, g, stream)) return st;
      } else {
        memset(&a, 0, sizeof(a));
        a.W = (const bf16*)L.out_proj; a.x = src; a.ldx = qn; a.M = M; a.N = d.d_model; a.K = qn;
        a.loop = loop; a.T_delayed = T_delayed; a.y = dst; a.ldy = d.d_model;
        if (last) {
          a.resid = x; a.ldr = d.d_model;
          if (zb_status st = launch_gemv<PRO_NONE, EPI_RESID>(ctx, a, stream)) return st;
        } else {
          if (zb_status st = launch_gemv<PRO_NONE, EPI_STORE>(ctx, a, stream)) return st;
        }
      }
      src = dst;
    }
    // 4. norm2 -> fc1 -> value * silu(gate)      5. fc2 + residual
    if (tc) {
      if (zb_status st = norm_rows((const bf16*)L.norm2_w, (const bf16*)L.norm2_b)) return st;
      zb_gemm_tc g;
      g.decode = (T == 1);
      g.W = (const bf16*)L.fc1; g.x = s.xn; g.ldx = d.d_model; g.M = M; g.N = 2 * d.d_ff; g.K = d.d_model; g.F = d.d_ff;
      g.epi = 3; g.y = s.h; g.ldy = d.d_ff;
      if (zb_status st = zb_launch_gemm_tc(ctx, g, stream)) return st;
      zb_gemm_tc g2;
      g2.decode = (T == 1);
      g2.W = (const bf16*)L.fc2; g2.x = s.h; g2.ldx = d.d_ff; g2.M = M; g2.N = d.d_model; g2.K = d.d_ff;
      g2.epi = 1; g2.y = x; g2.ldy = d.d_model; g2.resid = x; g2.ldr = d.d_model;
      if (zb_status st = zb_launch_gemm_tc(ctx, g2, stream)) return st;
    } else {
      memset(&a, 0, sizeof(a));
      a.W = (const bf16*)L.fc1; a.x = x; a.ldx = d.d_model; a.M = M; a.N = 2 * d.d_ff; a.K = d.d_model; a.F = d.d_ff;
      a.nw = (const bf16*)L.norm2_w; a.nb = (const bf16*)L.norm2_b; a.eps = d.norm_eps; a.norm_kind = d.norm_kind;
      a.y = s.h; a.ldy = d.d_ff; a.loop = loop; a.T_delayed = T_delayed;
      if (zb_status st = launch_gemv<PRO_NORM, EPI_SILU>(ctx, a, stream)) return st;
      memset(&a, 0, sizeof(a));
      a.W = (const bf16*)L.fc2; a.x = s.h; a.ldx = d.d_ff; a.M = M; a.N = d.d_model; a.K = d.d_ff;
      a.y = x; a.ldy = d.d_model; a.resid = x; a.ldr = d.d_model; a.loop = loop; a.T_delayed = T_delayed;
      if (zb_status st = launch_gemv<PRO_NONE, EPI_RESID>(ctx, a, stream)) return st;
    }
  }
  return ZB_OK;
}

zb_status zb_launch_final_norm(zb_ctx* ctx, const zb_model* model, const bf16* x, int R, int T, int last_only, bf16* y,
                               cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  NormArgs a;
  a.w = (const bf16*)d.norm_f_w; a.b = (const bf16*)d.norm_f_b; a.D = d.d_model; a.eps = d.norm_eps; a.kind = d.norm_kind;
  if (last_only) { a.x = x + (size_t)(T - 1) * d.d_model; a.ldx = (int64_t)T * d.d_model; a.y = y; a.ldy = d.d_model; }
  else { a.x = x; a.ldx = d.d_model; a.y = y; a.ldy = d.d_model; }
  norm_kernel<<<last_only ? R : R * T, 256, 0, stream>>>(a);
  ZB_CHECK_LAUNCH(ctx);
  return ZB_OK;
}

zb_status zb_launch_heads(zb_ctx* ctx, const zb_model* model, const bf16* hidden, int64_t row_stride, int R, int apply_norm,
                          float cfg_scale, float* logits, const zb_loop_state* loop, int T_delayed, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  GemvArgs a;
  memset(&a, 0, sizeof(a));
  a.W = (const bf16*)d.heads; a.x = hidden; a.ldx = row_stride; a.M = R; a.N = d.n_codebooks * d.head_vocab; a.K = d.d_model;
  a.nw = (const bf16*)d.norm_f_w; a.nb = (const bf16*)d.norm_f_b; a.eps = d.norm_eps; a.norm_kind = d.norm_kind;
  a.cfg_scale = cfg_scale; a.logits = logits; a.QV = a.N; a.loop = loop; a.T_delayed = T_delayed;
  if (cfg_scale != 1.0f) {
    ZB_REQUIRE(ctx, R % 2 == 0, "CFG needs an even number of rows");
    a.B = R / 2;
  } else a.B = R;
  if (R > 4 && R <= 256 && d.d_model % 64 == 0) {   // (runs even when the device loop is done: outputs are then unused)
    // dense path: (final norm ->) tcgen05 GEMM with the fp32 CFG-mix epilogue
    if (zb_status st = zb_tc_workspace_reserve(ctx, 0)) return st;
    const bf16* src = hidden;
    long long ld = row_stride;
    if (apply_norm) {
      bf16* xn = (bf16*)((char*)ctx->tc_ws + ((size_t)48 << 20));
      NormArgs na;
      na.x = hidden; na.ldx = row_stride; na.y = xn; na.ldy = d.d_model; na.w = (const bf16*)d.norm_f_w; na.b = (const bf16*)d.norm_f_b;
      na.D = d.d_model; na.eps = d.norm_eps; na.kind = d.norm_kind;
      ZB_CUDA(ctx, zb_launch_pdl(norm_kernel, dim3(R), dim3(256), 0, stream, na));
      ctx->launches++;
      src = xn; ld = d.d_model;
    }
    zb_gemm_tc g;
    g.W = (const bf16*)d.heads; g.x = src; g.ldx = ld; g.M = R; g.N = a.N; g.K = d.d_model; g.epi = 4;
    g.B = a.B; g.cfg_scale = cfg_scale; g.logits = logits; g.QV = a.N;
    return zb_launch_gemm_tc(ctx, g, stream);
  }
  if (apply_norm) return launch_gemv<PRO_NORM, EPI_HEADS>(ctx, a, stream);
  return launch_gemv<PRO_NONE, EPI_HEADS>(ctx, a, stream);
}


static unsigned long long* g_timeline = nullptr;   // debug: set by zb_debug_timeline
static unsigned long long* g_steplog = nullptr;    // debug: set by zb_debug_steplog

// ---- persistent decode step (host side) ----
size_t zb_mega_layers_bytes(const zb_model* model) { return (size_t)model->d.n_layer * sizeof(MegaLayer); }
static int mega_ipo(const zb_model_desc& d) { return 2 * d.d_inner + 2 * d.m_ngroups * d.d_state + (d.m_headdim ? d.d_inner / d.m_headdim : 0); }
size_t zb_mega_arena_bytes(const zb_model* model, int R) {
  const zb_model_desc& d = model->d;
  const size_t qn = (size_t)d.n_heads * d.head_dim, kn = (size_t)d.n_heads_kv * d.head_dim;
  const size_t mamba = model->n_mamba > 0 ? (size_t)mega_ipo(d) + 2 * d.d_inner : 0;  // in_proj output row, g (two 16-bit halves)
  return ((size_t)R * (2 * d.d_model + 2 * qn + d.d_ff + 2 * kn + mamba)) * sizeof(uint32_t);
}

bool zb_mega_tc_enabled(const zb_model* model, int R) {
  // Opt-in (read per session: tests compare both consumers in one process).  Measured on B200, full size, 200 frames:
  // 1.21 ms per step against 1.06 for the FFMA2 consumer at batch 1, 1.55 against 1.43 at batch 2.  The streaming phases
  // are HBM-bound with either consumer; what the tensor pipe saves there (fc1 9.2 -> 8.7 us) it loses in the fixed cost
  // of every phase (activation image + CTA barrier 0.8 us, MMA tail 0.5, accumulator read-back 0.5 against registers only).
  const int enabled = env_int("ZB_MEGA_TC", 0);
  MegaTcPlan p;
  return enabled && mega_tc_plan(model->d, model->ctx->num_sms, R, &p);
}

// (re)build the tile-ordered weight copy: once per model, and after zb_model_weights_changed()
static zb_status mega_tc_pretile(zb_ctx* ctx, const zb_model* model, const MegaTcPlan& p, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  if (model->tcw && (model->tcw_bytes < p.total || model->tcw_grid != ctx->num_sms)) {
    ZB_CUDA(ctx, cudaDeviceSynchronize());
    ZB_CUDA(ctx, cudaFree(model->tcw));
    model->tcw = nullptr; model->tcw_valid = false;
  }
  if (!model->tcw) {
    ZB_CUDA(ctx, cudaMalloc(&model->tcw, p.total));
    model->tcw_bytes = p.total; model->tcw_grid = ctx->num_sms; model->tcw_valid = false;
  }
  if (model->tcw_valid) return ZB_OK;
  auto run = [&](const void* W, size_t off, int kind, int N, int F) -> zb_status {
    PretileArgs a;
    a.W = (const bf16*)W; a.out = (uint4*)((char*)model->tcw + off); a.N = N; a.K = p.K[kind]; a.RB = p.g[kind].RB; a.RBv = p.g[kind].RBv; a.F = F;
    a.nunits = p.g[kind].nunits; a.S = p.g[kind].S;
    pretile_kernel<<<ctx->num_sms * 8, 256, 0, stream>>>(a);
    ZB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return ZB_OK;
  };
  const int qn = d.n_heads * d.head_dim, nqkv = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim;
  for (int li = 0; li < d.n_layer; ++li) {
    const zb_layer& L = model->layers[li];
    const size_t base = (size_t)li * p.layer_bytes;
    if (zb_status st = run(L.in_proj, base + p.off[TG_QKV], TG_QKV, nqkv, 0)) return st;
    if (zb_status st = run(L.out_proj, base + p.off[TG_OUT], TG_OUT, d.d_model, 0)) return st;
    if (zb_status st = run(L.fc1, base + p.off[TG_FC1], TG_FC1, 2 * d.d_ff, d.d_ff)) return st;
    if (zb_status st = run(L.fc2, base + p.off[TG_FC2], TG_FC2, d.d_model, 0)) return st;
  }
  if (zb_status st = run(d.heads, p.off[TG_HEADS], TG_HEADS, d.n_codebooks * d.head_vocab, 0)) return st;
  model->tcw_valid = true;
  (void)qn;
  return ZB_OK;
}

// ---- FP8 mode (opt-in, ZB_FP8=1): e4m3 copy of the decode matrices, [layer][in_proj | out_proj | fc1 | fc2][heads][scales] ----
struct MegaF8Plan {
  size_t off[4], layer_bytes, heads_off, scales_off;          // bytes
  size_t soff[4], layer_scales, heads_soff;                   // floats, relative to scales_off
  size_t total;
  int N[4], K[4];
};
static MegaF8Plan mega_f8_plan(const zb_model_desc& d) {
  MegaF8Plan p;
  const int qn = d.n_heads * d.head_dim, nqkv = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim;
  const int N[4] = {nqkv, d.d_model, 2 * d.d_ff, d.d_model}, K[4] = {d.d_model, qn, d.d_model, d.d_ff};
  size_t b = 0, f = 0;
  for (int i = 0; i < 4; ++i) { p.N[i] = N[i]; p.K[i] = K[i]; p.off[i] = b; p.soff[i] = f; b += (size_t)N[i] * K[i]; f += (size_t)N[i]; }
  p.layer_bytes = b; p.layer_scales = f;
  p.heads_off = (size_t)d.n_layer * b;
  const size_t hb = (size_t)d.n_codebooks * d.head_vocab * d.d_model;
  p.scales_off = (p.heads_off + hb + 255) / 256 * 256;
  p.heads_soff = (size_t)d.n_layer * f;
  p.total = p.scales_off + (p.heads_soff + (size_t)d.n_codebooks * d.head_vocab) * sizeof(float);
  return p;
}
// Read per session (tests compare both modes in one process).  Transformer stacks only: the Mamba2 phases keep bf16 weights.
bool zb_mega_fp8_enabled(const zb_model* model, int R) {
  return env_int("ZB_FP8", 0) && model->n_mamba == 0 && model->d.heads && !zb_mega_tc_enabled(model, R);
}
static zb_status mega_f8_build(zb_ctx* ctx, const zb_model* model, const MegaF8Plan& p, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  if (model->f8w && model->f8w_bytes < p.total) {
    ZB_CUDA(ctx, cudaDeviceSynchronize());
    ZB_CUDA(ctx, cudaFree(model->f8w));
    model->f8w = nullptr; model->f8w_valid = false;
  }
  if (!model->f8w) {
    ZB_CUDA(ctx, cudaMalloc(&model->f8w, p.total));
    model->f8w_bytes = p.total; model->f8w_valid = false;
  }
  if (model->f8w_valid) return ZB_OK;
  unsigned char* base = (unsigned char*)model->f8w;
  float* scales = (float*)(base + p.scales_off);
  auto run = [&](const void* W, size_t off, size_t soff, int N, int K) -> zb_status {
    quant_e4m3_kernel<<<N, 256, 0, stream>>>((const bf16*)W, base + off, scales + soff, K);
    ZB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return ZB_OK;
  };
  for (int li = 0; li < d.n_layer; ++li) {
    const zb_layer& L = model->layers[li];
    const void* W[4] = {L.in_proj, L.out_proj, L.fc1, L.fc2};
    for (int i = 0; i < 4; ++i)
      if (zb_status st = run(W[i], (size_t)li * p.layer_bytes + p.off[i], (size_t)li * p.layer_scales + p.soff[i], p.N[i], p.K[i])) return st;
  }
  if (zb_status st = run(d.heads, p.heads_off, p.heads_soff, d.n_codebooks * d.head_vocab, d.d_model)) return st;
  model->f8w_valid = true;
  return ZB_OK;
}

zb_status zb_mega_layers_build(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, void* host_buf, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  MegaTcPlan tp;
  const bool tc = model->n_mamba == 0 && mega_tc_plan(d, ctx->num_sms, 2, &tp) && zb_mega_tc_enabled(model, 2) && d.heads;
  if (tc) { if (zb_status st = mega_tc_pretile(ctx, model, tp, stream)) return st; }
  const bool f8 = !tc && zb_mega_fp8_enabled(model, 2);
  const MegaF8Plan fp = mega_f8_plan(d);
  if (f8) { if (zb_status st = mega_f8_build(ctx, model, fp, stream)) return st; }
  const size_t page_elems = (size_t)2 * d.n_heads_kv * ZB_PAGE_TOKENS * d.head_dim;
  MegaLayer* out = (MegaLayer*)host_buf;
  static const int share = env_int("ZB_DEBUG_SHARE_LAYERS", 0);   // debug: every layer streams layer 0's weights (L2-resident experiment)
  for (int li = 0; li < d.n_layer; ++li) {
    const zb_layer& L = model->layers[share ? 0 : li];
    memset(&out[li], 0, sizeof(MegaLayer));
    out[li].norm_w = (const bf16*)L.norm_w; out[li].norm_b = (const bf16*)L.norm_b; out[li].in_proj = (const bf16*)L.in_proj;
    out[li].out_proj = (const bf16*)L.out_proj;
    if (L.kind == ZB_LAYER_MAMBA2) {
      ZB_REQUIRE(ctx, cache->conv_state && cache->ssm_state, "persistent decode: Mamba2 layer %d but the cache has no conv/ssm state", li);
      const int nheads = d.d_inner / d.m_headdim, conv_dim = d.d_inner + 2 * d.m_ngroups * d.d_state, mi = model->mamba_index[li];
      out[li].kind = 1;
      out[li].conv_w = (const bf16*)L.conv_w; out[li].conv_b = (const bf16*)L.conv_b; out[li].dt_bias = (const bf16*)L.dt_bias;
      out[li].A_log = (const bf16*)L.A_log; out[li].Dp = (const bf16*)L.D; out[li].mnorm_w = (const bf16*)L.mnorm_w;
      out[li].conv_state = (bf16*)cache->conv_state + (size_t)mi * cache->rows * conv_dim * d.d_conv;
      out[li].ssm_state = (bf16*)cache->ssm_state + (size_t)mi * cache->rows * nheads * d.m_headdim * d.d_state;
      continue;
    }
    ZB_REQUIRE(ctx, L.kind == ZB_LAYER_ATTENTION, "persistent decode: layer %d is neither an attention nor a Mamba2 layer", li);
    out[li].norm2_w = (const bf16*)L.norm2_w; out[li].norm2_b = (const bf16*)L.norm2_b;
    out[li].fc1 = (const bf16*)L.fc1; out[li].fc2 = (const bf16*)L.fc2;
    out[li].kv_layer = (bf16*)cache->kv_pages + (size_t)model->attn_index[li] * cache->num_pages * page_elems;
    out[li].in_t = out[li].out_t = out[li].fc1_t = out[li].fc2_t = nullptr;
    if (tc) {
      const char* base = (const char*)model->tcw + (size_t)li * tp.layer_bytes;
      out[li].in_t = (const bf16*)(base + tp.off[TG_QKV]); out[li].out_t = (const bf16*)(base + tp.off[TG_OUT]);
      out[li].fc1_t = (const bf16*)(base + tp.off[TG_FC1]); out[li].fc2_t = (const bf16*)(base + tp.off[TG_FC2]);
    }
    if (f8) {                                                 // the weight pointers of this session's table are the e4m3 copy
      const char* base = (const char*)model->f8w + (size_t)(share ? 0 : li) * fp.layer_bytes;
      const float* sc = (const float*)((const char*)model->f8w + fp.scales_off) + (size_t)(share ? 0 : li) * fp.layer_scales;
      out[li].in_proj = (const bf16*)(base + fp.off[0]); out[li].out_proj = (const bf16*)(base + fp.off[1]);
      out[li].fc1 = (const bf16*)(base + fp.off[2]); out[li].fc2 = (const bf16*)(base + fp.off[3]);
      out[li].s_in = sc + fp.soff[0]; out[li].s_out = sc + fp.soff[1]; out[li].s_fc1 = sc + fp.soff[2]; out[li].s_fc2 = sc + fp.soff[3];
    }
  }
  return ZB_OK;
}

bool zb_mega_supported(const zb_model* model, int R) {
  const zb_model_desc& d = model->d;
  static const int enabled = env_int("ZB_DECODE_MEGA", 1);
  auto k_ok = [](int K) { return K == 512 || K == 1024 || K == 2048 || K == 4096; };   // 512 k per warp, 8 warps
  const int qn = d.n_heads * d.head_dim;
  if (model->n_mamba > 0) {
    // hybrid stacks: Mamba2 layers as three phases of the same kernel (ZB_MEGA_HYBRID=0: the multi-kernel graph path)
    const int grid = model->ctx->num_sms, ipo = mega_ipo(d);
    if (!env_int("ZB_MEGA_HYBRID", 1) || d.m_headdim != 64 || d.d_state != 128 || d.d_conv != 4 || d.m_ngroups != 1 || d.d_inner != 2 * d.d_model ||
        !k_ok(d.d_inner) || ((ipo + grid - 1) / grid) * R > kMW * 32)
      return false;
  }
  return enabled && R >= 2 && R <= 4 && d.head_dim == kHD && k_ok(d.d_model) && k_ok(qn) && (k_ok(d.d_ff) || d.d_ff == 8192) &&
         d.n_codebooks <= 16 && d.n_heads / d.n_heads_kv <= 8 && (d.out_proj_repeats == 1 || qn == d.d_model) &&
         d.out_proj_repeats >= 1 && d.out_proj_repeats <= 2 && (((d.n_heads + 2 * d.n_heads_kv) * d.head_dim) % 2 == 0);
}

zb_status zb_launch_decode_step(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* mega_layers_dev, unsigned* sync,
                                uint32_t* arena, int R, int max_kv_len, float cfg_scale, float* logits, const int64_t* delayed, int T_delayed,
                                const zb_loop_state* loop, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  const int nsplit = (max_kv_len + kCH - 1) / kCH;
  size_t need = 0;
  carve(model, nullptr, R, nsplit, &need);
  ZB_REQUIRE(ctx, ctx->scratch_bytes >= need, "internal: scratch not reserved (%zu < %zu)", ctx->scratch_bytes, need);
  Scratch s = carve(model, ctx->scratch, R, nsplit, nullptr);
  MegaArgs m;
  memset(&m, 0, sizeof(m));
  m.layers = (const MegaLayer*)mega_layers_dev; m.n_layer = d.n_layer;
  m.D = d.d_model; m.F = d.d_ff; m.Hq = d.n_heads; m.Hkv = d.n_heads_kv; m.hd = d.head_dim; m.eps = d.norm_eps; m.norm_kind = d.norm_kind;
  m.rope_interleaved = d.rope_interleaved; m.out_proj_repeats = d.out_proj_repeats;
  m.normf_w = (const bf16*)d.norm_f_w; m.normf_b = (const bf16*)d.norm_f_b; m.heads = (const bf16*)d.heads; m.QV = d.n_codebooks * d.head_vocab;
  m.B = R / 2; m.cfg_scale = cfg_scale; m.logits = logits; m.rope = d.rope_table; m.rope_len = d.rope_len;
  m.lengths = cache->lengths; m.page_table = cache->page_table; m.max_pages = cache->max_pages_per_row;
  for (int k = 0; k < d.n_codebooks; ++k) m.emb[k] = (const bf16*)model->emb[k];
  m.Q = d.n_codebooks; m.vocab = d.emb_vocab; m.delayed = delayed; m.T_delayed = T_delayed;
  {  // tagged activation words: the session's own zero-initialised arena (zb_mega_arena_bytes)
    const size_t qn_ = (size_t)d.n_heads * d.head_dim, kn_ = (size_t)d.n_heads_kv * d.head_dim;
    uint32_t* p = arena;
    m.xt = p; p += (size_t)R * d.d_model;
    m.qt = p; p += (size_t)R * qn_;
    m.ayt = p; p += (size_t)R * qn_;
    m.y1t = p; p += (size_t)R * d.d_model;
    m.ht = p; p += (size_t)R * d.d_ff;
    m.kvt = p; p += (size_t)R * 2 * kn_;
  }
  m.nph = 2;
  for (int li = 0; li < d.n_layer; ++li) m.nph += model->layers[li].kind == ZB_LAYER_MAMBA2 ? 3 : 4 + d.out_proj_repeats;
  if (model->n_mamba > 0) {
    m.d_inner = d.d_inner; m.m_nheads = d.d_inner / d.m_headdim; m.ipo = mega_ipo(d); m.conv_dim = d.d_inner + 2 * d.m_ngroups * d.d_state;
    const size_t qn_ = (size_t)d.n_heads * d.head_dim, kn_ = (size_t)d.n_heads_kv * d.head_dim;
    uint32_t* p = arena + (size_t)R * (2 * d.d_model + 2 * qn_ + d.d_ff + 2 * kn_);
    m.zxt = p; p += (size_t)R * m.ipo;
    m.ygt = p;
  }
  m.attn_part = s.part; m.attn_counters = ctx->counters; m.nsplit = nsplit;
  m.scale = 1.0f / sqrtf((float)d.head_dim); m.loop = loop; m.sync = sync; m.timeline = g_timeline; m.steplog = g_steplog;
  ZB_REQUIRE(ctx, cfg_scale != 1.0f, "persistent decode expects CFG rows");
  const int grid = ctx->num_sms;
  // partial-sum buffer: the largest padded row count x k-slices over all matrices of the step
  // FP8 mode (decided below from the same inputs): R <= 2 doubles the rows of a stage (RW x 2 in the kernel's instantiations)
  const bool f8_early = model->n_mamba == 0 && zb_mega_fp8_enabled(model, R) && model->f8w && model->f8w_valid;
  const int rm = (f8_early && R <= 2) ? 2 : 1;
  auto part_need = [&](int nunits, bool pairs, int K, int NC, int RW) {
    RW *= rm;
    const int KS = K / (NC * 256), RPS = (kMW / KS) * RW;
    const int rows = ((nunits + grid - 1) / grid) * (pairs ? 2 : 1);
    return (size_t)((rows + RPS - 1) / RPS * RPS) * KS * R * sizeof(float);
  };
  const int qn = d.n_heads * d.head_dim;
  size_t pb = part_need((d.n_heads + 2 * d.n_heads_kv) * d.head_dim / 2, true, d.d_model, 2, 4);
  pb = std::max(pb, part_need(d.d_model, false, qn, 2, 4));
  pb = std::max(pb, part_need(d.d_ff, true, d.d_model, 2, 4));
  pb = std::max(pb, d.d_ff == 8192 ? part_need(d.d_model, false, d.d_ff, 4, 2) : part_need(d.d_model, false, d.d_ff, 2, 4));
  pb = std::max(pb, part_need(m.QV, false, d.d_model, 2, 4));
  if (model->n_mamba > 0) {
    pb = std::max(pb, part_need(m.ipo, false, d.d_model, 2, 4));
    pb = std::max(pb, part_need(d.d_model, false, d.d_inner, 2, 4));
  }
  pb = (pb + 1023) / 1024 * 1024;
  {  // one epilogue item per consumer thread
    const int max_items = std::max(((d.d_ff + grid - 1) / grid) * R, ((m.QV + grid - 1) / grid) * R);
    ZB_REQUIRE(ctx, max_items <= kMW * 32 && ((d.d_model + grid - 1) / grid) * R <= kMW * 32, "persistent decode: %d epilogue items per CTA", max_items);
  }
  MegaTcPlan tp;
  const bool tc = model->n_mamba == 0 && zb_mega_tc_enabled(model, R) && mega_tc_plan(d, grid, R, &tp) && model->tcw && model->tcw_valid;
  if (tc) {
    for (int k = 0; k < TG_COUNT; ++k) m.tg[k] = tp.g[k];
    m.heads_t = (const bf16*)((const char*)model->tcw + tp.off[TG_HEADS]);
    pb = kMegaTcPartBytes;
  }
  // FP8 mode: the session's layer table (zb_mega_layers_build) points at the e4m3 copy; a stage holds the same rows in 16 KB
  const bool f8 = !tc && zb_mega_fp8_enabled(model, R) && model->f8w && model->f8w_valid;
  if (f8) {
    const MegaF8Plan fp = mega_f8_plan(d);
    m.heads = (const bf16*)((const char*)model->f8w + fp.heads_off);
    m.s_heads = (const float*)((const char*)model->f8w + fp.scales_off) + fp.heads_soff;
  }
  ZB_REQUIRE(ctx, f8 == f8_early, "internal: FP8 mode decided two ways");
  const size_t stage_bytes = f8 ? (size_t)(kMegaStageBytes / 2) * rm : (size_t)kMegaStageBytes;
  const size_t attn_bytes = kMegaAttnBytes + (size_t)4 * d.d_model * sizeof(bf16);   // attention tiles + two norm-parameter buffers
  const size_t avail = 227 * 1024 - 2048 - (tc ? kMegaTcImageBytes + 1024 : 0);      // (tc: activation image, 1024-byte alignment of the ring)
  int stages = (int)((avail - pb - attn_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  ZB_REQUIRE(ctx, stages >= 3, "persistent decode: not enough shared memory for the ring");
  m.ring_stages = stages; m.part_bytes = (int)pb;
  static const int evict_first = env_int("ZB_MEGA_EVICT_FIRST", 1);
  m.evict_first = evict_first;
  const size_t smem = (size_t)stages * stage_bytes + pb + attn_bytes + (tc ? kMegaTcImageBytes + 1024 : 0);
  auto launch = [&](auto kernel) -> zb_status {
    ZB_CUDA(ctx, zb_ensure_smem(ctx, kernel, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3((kMW + 1) * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;      // all CTAs must be co-resident: every phase spins on words the other CTAs write
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ZB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, m));
    ctx->launches++;
    return ZB_OK;
  };
  if (env_int("ZB_MEGA_VERBOSE", 0) && loop && ctx->launches % 64 == 0)
    fprintf(stderr, "[zb] persistent decode step: %s consumer, %d stages, smem %zu\n", tc ? "tcgen05" : f8 ? "FP8 (e4m3 -> HFMA2)" : "FFMA2", stages, smem);
  if (tc) return R <= 2 ? launch(decode_step_kernel<2, 1>) : launch(decode_step_kernel<4, 1>);
  if (f8) return R <= 2 ? launch(decode_step_kernel<2, 2>) : launch(decode_step_kernel<4, 2>);
  if (R <= 2) return launch(decode_step_kernel<2, 0>);
  return launch(decode_step_kernel<4, 0>);
}

// ---- diagnostics: launch ONE production kernel on scratch activations (bench.py roofline leg) ----
extern "C" ZB_API zb_status zb_debug_timeline(unsigned long long* dev_buf) { g_timeline = dev_buf; return ZB_OK; }
extern "C" ZB_API zb_status zb_debug_steplog(unsigned long long* dev_buf) { g_steplog = dev_buf; return ZB_OK; }
unsigned long long* zb_debug_steplog_ptr() { return g_steplog; }

extern "C" zb_status zb_bench_kernel(zb_ctx* ctx, const zb_model* model, int32_t layer, int32_t which, int32_t rows, int32_t iters,
                                     zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, model && layer >= 0 && layer < model->d.n_layer && rows >= 1 && rows <= 8 && iters >= 1, "zb_bench_kernel: bad arguments");
  const zb_model_desc& d = model->d;
  const size_t need = zb_backbone_scratch_bytes(model, rows, 1, ZB_PAGE_TOKENS) + (size_t)rows * d.d_model * 2 + 256;
  if (zb_status st = zb_scratch_reserve(ctx, need)) return st;
  Scratch s = carve(model, ctx->scratch, rows, 1, nullptr);
  bf16* x = (bf16*)((char*)ctx->scratch + (ctx->scratch_bytes - (((size_t)rows * d.d_model * 2 + 255) / 256 * 256)));
  for (int it = 0; it < iters; ++it) {
    const zb_layer& L = model->layers[(layer + it) % d.n_layer];      // cycle layers: every launch streams from HBM
    GemvArgs a;
    memset(&a, 0, sizeof(a));
    a.M = rows;
    a.timeline = g_timeline ? g_timeline + (size_t)it * 8 : nullptr;
    zb_status st;
    if (which == 2) {          // norm2 -> fc1 -> value * silu(gate)
      a.W = (const bf16*)L.fc1; a.x = x; a.ldx = d.d_model; a.N = 2 * d.d_ff; a.K = d.d_model; a.F = d.d_ff;
      a.nw = (const bf16*)L.norm2_w; a.nb = (const bf16*)L.norm2_b; a.eps = d.norm_eps; a.norm_kind = d.norm_kind;
      a.y = s.h; a.ldy = d.d_ff;
      st = launch_gemv<PRO_NORM, EPI_SILU>(ctx, a, (cudaStream_t)stream);
    } else if (which == 3) {   // fc2 + residual
      a.W = (const bf16*)L.fc2; a.x = s.h; a.ldx = d.d_ff; a.N = d.d_model; a.K = d.d_ff;
      a.y = x; a.ldy = d.d_model; a.resid = x; a.ldr = d.d_model;
      st = launch_gemv<PRO_NONE, EPI_RESID>(ctx, a, (cudaStream_t)stream);
    } else if (which == 1) {   // out_proj
      a.W = (const bf16*)L.out_proj; a.x = s.attn_y; a.ldx = d.n_heads * d.head_dim; a.N = d.d_model; a.K = d.n_heads * d.head_dim;
      a.y = s.y1; a.ldy = d.d_model;
      st = launch_gemv<PRO_NONE, EPI_STORE>(ctx, a, (cudaStream_t)stream);
    } else {
      return zb_fail(ctx, ZB_ERR_INVALID, "zb_bench_kernel: unknown kernel id %d", which);
    }
    if (st) return st;
  }
  return ZB_OK;
}
