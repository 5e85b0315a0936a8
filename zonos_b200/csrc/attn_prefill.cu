// Causal attention over the paged KV cache for calls with more than one token per row (prefill, chunked prefill).
//
// Replaces F.scaled_dot_product_attention(q, k, v, is_causal=seqlen > 1) of zonos/backbone/_torch.py:415 for T > 1.
// grid (rows, kv heads, query tiles); a CTA of 8 warps owns 8 x (8 / G) query tokens of one (row, kv head) pair - every
// warp 8 query columns = (8 / G tokens) x (G heads of the GQA group) - and walks the pair's keys in 64-token chunks
// (= one KV page: a contiguous 16 KB run for K and for V), double-buffered in shared memory with cp.async in the
// 128-byte-swizzle layout mma.cuh expects.  S^T = K q^T and O^T = V^T P^T run on mma.sync tensor-core tiles with an
// online softmax in the accumulator fragments; the causal limit is per column.  Prefill attention is < 1 % of the
// prefill FLOPs (the Linears dominate: 3.4 GFLOP per token against 0.02 here at 160 tokens), so the legacy mma.sync
// shape is enough; what mattered was to stop re-reading K/V once per query token with scalar FMAs.
// Chunked prefill (tokens already cached) uses bottom-right alignment: query t sees keys 0 .. lengths[r] + t.
#include "internal.h"
#include "mma.cuh"

namespace {

constexpr int kPfWarps = 8;
constexpr int kPfThreads = kPfWarps * 32;
constexpr int kPfHd = 128;
constexpr int kPfChunk = ZB_PAGE_TOKENS;                     // 64 keys per chunk
constexpr int kPfTile = 2 * kPfChunk * kPfHd * 2;            // K + V of one chunk: 32 KB

struct AttnPfArgs {
  const bf16* q;            // [rows * T, Hq * 128]
  const bf16* kv_layer;     // pages of this layer
  const int32_t* lengths; const int32_t* page_table; int max_pages;
  int T, Hq, Hkv, G;
  float scale;
  bf16* y;                  // [rows * T, Hq * 128]
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;                              // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(kPfThreads) attn_prefill_kernel(AttnPfArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_pf[];
  pdl_launch_dependents();
  pdl_wait();                                                // q and this call's K/V come from the in_proj kernel
  const int r = blockIdx.x, g = blockIdx.y;
  const int G = a.G, tpw = 8 / G, qt = kPfWarps * tpw;       // query tokens per warp / per CTA
  const int t0 = blockIdx.z * qt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = lane >> 2, lr = lane & 3, n0 = 2 * lr;
  const int old_len = a.lengths[r];
  const int t_hi = min(a.T, t0 + qt);                        // this CTA's query tokens: [t0, t_hi)
  const int kv_end = old_len + t_hi;                         // keys any of them may see: [0, kv_end)
  const int nchunk = (kv_end + kPfChunk - 1) / kPfChunk;
  const uint32_t sbase = smem_u32(smem_pf);

  auto load_chunk = [&](int c, int buf) {
    const int page = a.page_table[(size_t)r * a.max_pages + c];
    const bf16* kp = a.kv_layer + (((size_t)page * 2 + 0) * a.Hkv + g) * kPfChunk * kPfHd;
    const bf16* vp = kp + (size_t)a.Hkv * kPfChunk * kPfHd;
    const uint32_t kb = sbase + (uint32_t)buf * kPfTile, vb = kb + 16384;
    for (int i = threadIdx.x; i < kPfChunk * 16; i += kPfThreads) {
      const int tok = i >> 4, ch = i & 15;
      const bool valid = c * kPfChunk + tok < kv_end;        // rows past the last visible key: zeros (0 * NaN = NaN otherwise)
      cp_async16_zfill(kv_chunk_addr(kb, tok, ch), kp + (size_t)tok * kPfHd + ch * 8, valid);
      cp_async16_zfill(kv_chunk_addr(vb, tok, ch), vp + (size_t)tok * kPfHd + ch * 8, valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // this thread's fragment columns: column n = lane / 4 for the q fragments, n0 + {0, 1} for the accumulators
  const int tq_b = t0 + warp * tpw + lq / G;                 // query token of B-fragment column lq
  uint32_t qf[8][2];
  {
    const bool ok = tq_b < a.T;
    const bf16* qp = a.q + ((size_t)r * a.T + (ok ? tq_b : 0)) * a.Hq * kPfHd + (size_t)(g * G + lq % G) * kPfHd + 2 * lr;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qf[ks][0] = ok ? *reinterpret_cast<const unsigned*>(qp + 16 * ks) : 0u;
      qf[ks][1] = ok ? *reinterpret_cast<const unsigned*>(qp + 16 * ks + 8) : 0u;
    }
  }
  int kvlen[2];                                              // keys visible to accumulator column j: 0 .. kvlen - 1
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int tq = t0 + warp * tpw + (n0 + j) / G;
    kvlen[j] = tq < a.T ? old_len + tq + 1 : 0;
  }
  const int warp_kv = min(kv_end, old_len + min(a.T, t0 + (warp + 1) * tpw));      // keys this warp needs at all

  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
  float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};

  if (nchunk > 0) load_chunk(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) {
      load_chunk(c + 1, (c + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (c * kPfChunk < warp_kv) {
      const uint32_t kb = sbase + (uint32_t)(c & 1) * kPfTile, vb = kb + 16384;
      const int lim[2] = {kvlen[0] - c * kPfChunk, kvlen[1] - c * kPfChunk};
      attn_tile64(kb, vb, qf, a.scale, lim, o, mrun, lrun, lane);
    }
    __syncthreads();                                         // the buffer is refilled two iterations later
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    lrun[j] += __shfl_xor_sync(0xffffffffu, lrun[j], 4);
    lrun[j] += __shfl_xor_sync(0xffffffffu, lrun[j], 8);
    lrun[j] += __shfl_xor_sync(0xffffffffu, lrun[j], 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int n = n0 + j, tq = t0 + warp * tpw + n / G;
    if (tq < a.T) {
      const float inv = 1.0f / lrun[j];
      bf16* on = a.y + ((size_t)r * a.T + tq) * a.Hq * kPfHd + (size_t)(g * G + n % G) * kPfHd;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { on[16 * dt + lq] = f2bf(o[dt][j] * inv); on[16 * dt + lq + 8] = f2bf(o[dt][2 + j] * inv); }
    }
  }
}

}  // namespace

bool zb_attn_prefill_supported(const zb_model_desc& d) {
  const int G = d.n_heads_kv > 0 ? d.n_heads / d.n_heads_kv : 0;
  return d.head_dim == kPfHd && d.n_heads_kv > 0 && d.n_heads % d.n_heads_kv == 0 && (G == 1 || G == 2 || G == 4 || G == 8);
}

zb_status zb_launch_attn_prefill(zb_ctx* ctx, const zb_model_desc& d, const zb_cache* cache, const bf16* q, const bf16* kv_layer, bf16* y, int R, int T,
                                 cudaStream_t stream) {
  AttnPfArgs a;
  a.q = q; a.kv_layer = kv_layer; a.lengths = cache->lengths; a.page_table = cache->page_table; a.max_pages = cache->max_pages_per_row;
  a.T = T; a.Hq = d.n_heads; a.Hkv = d.n_heads_kv; a.G = d.n_heads / d.n_heads_kv; a.scale = 1.0f / sqrtf((float)d.head_dim); a.y = y;
  const int qt = kPfWarps * (8 / a.G);
  const size_t smem = 2 * (size_t)kPfTile;
  ZB_CUDA(ctx, zb_ensure_smem(ctx, attn_prefill_kernel, smem));
  ZB_CUDA(ctx, zb_launch_pdl(attn_prefill_kernel, dim3(R, d.n_heads_kv, (T + qt - 1) / qt), dim3(kPfThreads), smem, stream, a));
  ctx->launches++;
  return ZB_OK;
}
