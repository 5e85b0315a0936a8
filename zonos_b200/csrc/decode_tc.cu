// Persistent decode step with a tcgen05 / TMEM consumer, for R = 2 ... 128 activation rows (batch 1 ... 64 with CFG).
//
// ONE cooperative launch per codec frame runs  embed -> n_layer x (in_proj, attention, out_proj x repeats, fc1, fc2)
// -> final norm + heads + CFG mix.  148 CTAs (one per SM) stay resident; each owns ONE work unit of every weight
// matrix of the step:  unit = (block of <= 128 weight rows, slice of the K dimension).
//
//   warp 0      TMA producer of everything that does NOT depend on this step's activations: the CTA's weight tiles of
//               ALL matrices and the K/V tiles (tokens cached by earlier steps) of ALL attention phases, in consumption
//               order, through ONE shared-memory ring of 16 KB slots (cp.async.bulk.tensor.2d, 128-byte swizzle,
//               mbarrier complete_tx).  HBM keeps streaming across phase boundaries and grid barriers.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer.  "Swap-AB": the weight rows are the A operand
//               (UMMA M = 128 lanes), the activation rows are the B operand (UMMA N = 16 ... 128), the fp32 accumulator
//               D[weight row, activation row] lives in tensor memory.
//   warp 2      TMA producer of the activation operand (a second, shallower ring: the tiles come from L2), released
//               per GEMM by the compute warps once the grid barrier in front of the phase has been passed.
//   warps 3-10  drain the accumulator (tcgen05.ld) into fp32 partials, run the element-wise sub-phases between the
//               GEMMs, and run attention: every warp owns a contiguous range of 32-token K/V tiles and computes
//               S^T = K q^T and O^T = V^T P^T with mma.sync (tokens on the M dimension, the <= 8 query heads of the
//               kv head on N), online softmax in the accumulator fragments.
//
// Every matrix phase is: GEMM partials (K slices of one row block meet in L2) -> grid barrier -> reduce + fused
// epilogue (residual add + the NEXT norm, RoPE + paged KV append, SiLU gate, CFG mix) -> grid barrier.  The residual
// epilogues own whole rows (one CTA per activation row), so the normalised operand of the next GEMM is written right
// there and every B operand is a plain bf16 matrix that TMA can tile.  All reductions run in a fixed order: results do
// not depend on timing.
//
// Reference call sites replaced: the same as decode.cu (zonos/backbone/_torch.py:326-328,401,57-68,105-106,415,
// 419-420,473-474,238; zonos/utilities/codec_utils.py:37,68-79; zonos/model.py:229-233).  Rounding points are the
// reference's: every Linear / norm / SiLU / residual add rounds to bf16; RoPE and the CFG mix are un-contracted fp32.
#include <stdlib.h>

#include <algorithm>

#include "mma.cuh"
#include "tc.cuh"

namespace {

constexpr int kCW = 8;                       // compute warps
constexpr int kCThreads = kCW * 32;          // 256
constexpr int kTcThreads = 96 + kCThreads;   // + weight/KV producer warp + MMA warp + activation producer warp
constexpr int kMaxParts = 4;                 // attention parts ((row, kv head) pair x split) per warp
constexpr int kMaxSplit = 8;                 // splits of a pair's tokens when there are fewer pairs than attention warps
constexpr int kMaxSlots = 7;
constexpr int kMaxBSlots = 6;
constexpr int kHd = 128;
constexpr int kSegN = 256;                   // features per element-wise warp task
constexpr int kPartStride = 132;             // floats per attention partial: o[128], max, sum, pad
constexpr int kSlot = 32 * 1024;             // ring slot: two k-blocks of 128 weight rows, or 64 tokens (one page) of K and V.  Measured
                                             // (scripts/probes/tma_tensor_probe.cu): a ring moves ~0.3 us per STAGE whatever its size up to 32 KB
constexpr int kKbSlot = 2;                   // k-blocks per slot
constexpr int kTileTok = 64;                 // tokens per attention tile
constexpr unsigned kSpinLimit = 1u << 22;

enum { G_QKV = 0, G_OUT = 1, G_FC1 = 2, G_FC2 = 3, G_HEADS = 4, G_COUNT = 5 };
enum { E_STORE = 0, E_RESID = 1, E_QKV = 2, E_SILU = 3, E_HEADS = 4 };
enum { M_IN = 0, M_OUT = 1, M_FC1 = 2, M_FC2 = 3, M_KV = 4, M_COUNT = 5 };          // tensor maps of a layer
enum { B_XN = 1, B_AY = 2, B_Y1 = 3, B_H = 4 };                                      // activation maps (table entry n_layer; [0] = heads)

struct TcGemm { int N, K, Nw, RB, RBv, nrb, nks, _pad; };
struct alignas(64) TcLayer {
  CUtensorMap map[M_COUNT];
  const bf16 *norm_w, *norm_b, *norm2_w, *norm2_b;
  bf16* kv_layer;
  char pad[24];
};
static_assert(sizeof(TcLayer) % 64 == 0, "tensor maps must stay 64-byte aligned");

struct TcArgs {
  const TcLayer* layers; int n_layer;
  TcGemm g[G_COUNT];
  int R, B, Rp;
  int D, F, Hq, Hkv, G; float eps; int norm_kind, rope_interleaved, out_proj_repeats;
  const bf16 *normf_w, *normf_b; int QV; float cfg_scale; float* logits;
  const float* rope; int rope_len;
  const int32_t* lengths; const int32_t* page_table; int max_pages;
  const bf16* emb[16]; int Q, vocab; const int64_t* delayed; int T_delayed;
  bf16 *x, *xn, *q, *ay, *y1, *h; float* ws; float* attn_part; unsigned* pair_cnt;
  unsigned* bar; const zb_loop_state* loop;
  int stages, bstages, bslot_bytes, nbar, na, fc1_direct, l2pf; float scale;
  unsigned long long* timeline;
};

struct Unit { int active, rb, ks, kb0, kb1; };
__device__ __forceinline__ Unit unit_of(const TcGemm& g) {
  Unit u;
  const int b = blockIdx.x;
  u.active = b < g.nrb * g.nks;
  u.rb = b / g.nks; u.ks = b % g.nks;
  const int nkb = g.K / 64;
  u.kb0 = (int)((long long)u.ks * nkb / g.nks);
  u.kb1 = (int)((long long)(u.ks + 1) * nkb / g.nks);
  if (!u.active) u.kb0 = u.kb1 = 0;
  return u;
}
// j-th k-block of a unit: the row blocks that share a K slice start at different offsets, so that at any moment the
// CTAs pull DIFFERENT lines of the (shared, L2-resident) activation operand instead of all hammering the same ones
__device__ __forceinline__ int unit_kb(const Unit& u, int j) {
  const int n = u.kb1 - u.kb0;
  return u.kb0 + (j + u.rb * 5) % n;
}

__device__ __forceinline__ unsigned long long gtime_tc() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kCThreads) : "memory"); }   // compute warps only
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- grid barrier among the compute warps of all CTAs (cooperative launch: every CTA is resident) ----------------
__device__ __forceinline__ void grid_barrier(const TcArgs& a, unsigned& epoch, int ctid) {
  fence_proxy_async_all();                                  // this thread's global writes may be read by TMA (async proxy) in other CTAs
  cbar();
  ++epoch;
  if (ctid == 0) {
    // release: cumulative over the writes of the whole CTA (ordered before this thread by the CTA barrier above)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
    const unsigned target = epoch * gridDim.x;              // the counter is never reset inside a session: epoch continues across steps
    unsigned v;
    for (unsigned spins = 0;; ++spins) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
      if (v >= target) break;
      if (spins > kSpinLimit) asm volatile("trap;");      // a lost CTA must end in an error, not a hung GPU
    }
  }
  cbar();
}

// ---- attention schedule -------------------------------------------------------------------------------------------
// Work item ("part") = one (row, kv head) pair, or 1/nsplit of its 64-token tiles when there are fewer pairs than
// attention warps.  Part i runs on CTA i % grid, warp (i / grid) % na; a warp owns at most kMaxParts parts.  Only the
// tokens cached by EARLIER steps come through the ring; this step's own token is folded in by the pair's last split.
struct AttnSched {
  int nsplit;
  int np[kCW];                                   // parts of warp w
  int pair[kCW][kMaxParts], sidx[kCW][kMaxParts], c0[kCW][kMaxParts], c1[kCW][kMaxParts];   // tiles [c0, c1) of the pair
  int ntile[kCW];                                // tiles of warp w in total
};
// called by one full warp: lane = w * kMaxParts + q
__device__ __forceinline__ void attn_schedule(const TcArgs& a, AttnSched& s, int lane) {
  static_assert(kCW * kMaxParts == 32, "one lane per (warp, part)");
  const int P = a.R * a.Hkv, NWA = gridDim.x * a.na;
  const int nsplit = max(1, min(kMaxSplit, NWA / P));
  const int w = lane / kMaxParts, q = lane % kMaxParts;
  const int i = blockIdx.x + gridDim.x * (w + a.na * q);
  const bool valid = w < a.na && i < P * nsplit;
  int pair = 0, sidx = 0, c0 = 0, c1 = 0;
  if (valid) {
    pair = i / nsplit; sidx = i % nsplit;
    const int nt = max(1, (a.lengths[pair / a.Hkv] + kTileTok - 1) / kTileTok);
    c0 = sidx * nt / nsplit; c1 = (sidx + 1) * nt / nsplit;
  }
  s.pair[w][q] = pair; s.sidx[w][q] = sidx; s.c0[w][q] = c0; s.c1[w][q] = c1;
  int n = c1 - c0;
  n += __shfl_xor_sync(0xffffffffu, n, 1);
  n += __shfl_xor_sync(0xffffffffu, n, 2);
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if (q == 0) { s.ntile[w] = n; s.np[w] = __popc((bal >> (w * kMaxParts)) & ((1u << kMaxParts) - 1)); }
  if (lane == 0) s.nsplit = nsplit;
}
// ring order of a CTA's attention tiles: round robin over its warps (i-th tile of warp 0, of warp 1, ...)
__device__ __forceinline__ int attn_seq(const AttnSched& s, int w, int i) {
  int seq = 0;
#pragma unroll
  for (int v = 0; v < kCW; ++v) {
    const int nv = s.ntile[v];
    seq += min(nv, i) + ((v < w && nv > i) ? 1 : 0);
  }
  return seq;
}
__device__ __forceinline__ int attn_tiles_cta(const AttnSched& s) {
  int n = 0;
#pragma unroll
  for (int v = 0; v < kCW; ++v) n += s.ntile[v];
  return n;
}

// ---- producer (warp 0, one thread): the CTA's unit of one weight matrix through the ring ----------------------------
__device__ __forceinline__ void ring_acquire(uint64_t* empty_bar, int S, int gst) {
  const int slot = gst % S, use = gst / S;
  if (use > 0) mbar_wait(&empty_bar[slot], (uint32_t)((use - 1) & 1));
}
// Consumer side.  The ring has SEVERAL consumers (the MMA thread and the attention warps), so a consumer may start to
// wait for use u of a slot while the slot's use u-1 (another consumer's tile) is still in flight; a parity wait would
// then return at once.  slot_use[slot] = last use whose data has been seen to land: wait for u-1 there first.
__device__ __forceinline__ void ring_wait_full(uint64_t* full_bar, volatile int* slot_use, int S, int seq, bool writer) {
  const int slot = seq % S, use = seq / S;
  if (use > 0) {
    for (unsigned spins = 0; slot_use[slot] < use - 1; ++spins)
      if (spins > (1u << 26)) asm volatile("trap;");
  }
  mbar_wait(&full_bar[slot], (uint32_t)(use & 1));
  if (writer) slot_use[slot] = use;
}
__device__ __forceinline__ void tc_produce(const TcArgs& a, const TcGemm& g, const CUtensorMap* map, int kind, unsigned char* ring,
                                           uint64_t* full_bar, uint64_t* empty_bar, int& gst) {
  const Unit u = unit_of(g);
  const int S = a.stages, n = u.kb1 - u.kb0;
  for (int j = 0; j < n; j += kKbSlot, ++gst) {
    const int cnt = min(kKbSlot, n - j);
    ring_acquire(empty_bar, S, gst);
    const int slot = gst % S;
    mbar_expect_tx(&full_bar[slot], (uint32_t)(cnt * g.RB * 128));    // out-of-bounds rows are zero-filled and still counted
    for (int e = 0; e < cnt; ++e) {
      const int kb = unit_kb(u, j + e);
      unsigned char* dst = ring + (size_t)slot * kSlot + (size_t)e * (kSlot / kKbSlot);
      if (kind == G_FC1) {                                            // value rows | gate rows of the same features
        tma_load_2d(dst, map, kb * 64, u.rb * g.RBv, &full_bar[slot]);
        tma_load_2d(dst + g.RBv * 128, map, kb * 64, a.F + u.rb * g.RBv, &full_bar[slot]);
      } else {
        tma_load_2d(dst, map, kb * 64, u.rb * g.RB, &full_bar[slot]);
      }
    }
  }
}
// L2 prefetch of the CTA's whole unit of a weight matrix, issued one GEMM ahead of the ring: HBM then keeps streaming
// while the ring (4-6 slots) is full, i.e. during the dump / barrier / epilogue sub-phases, and the ring fills from L2
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_prefetch(const TcArgs& a, const TcGemm& g, const CUtensorMap* map, int kind) {
  if (!a.l2pf) return;
  const Unit u = unit_of(g);
  for (int j = 0; j < u.kb1 - u.kb0; ++j) {
    const int kb = unit_kb(u, j);
    if (kind == G_FC1) {
      tma_prefetch_2d(map, kb * 64, u.rb * g.RBv);
      tma_prefetch_2d(map, kb * 64, a.F + u.rb * g.RBv);
    } else {
      tma_prefetch_2d(map, kb * 64, u.rb * g.RB);
    }
  }
}
// the CTA's K/V tiles of one attention phase: slot = K[64 tok][d 0..63] | K[..][d 64..127] | V likewise (8 KB boxes)
__device__ __forceinline__ void tc_produce_kv(const TcArgs& a, const AttnSched& s, const CUtensorMap* map, unsigned char* ring, uint64_t* full_bar,
                                              uint64_t* empty_bar, int& gst) {
  const int S = a.stages;
  int nmax = 0;
  int cq[kCW], cc[kCW];                                               // cursor of every warp: part, tile
#pragma unroll
  for (int w = 0; w < kCW; ++w) { nmax = max(nmax, s.ntile[w]); cq[w] = 0; cc[w] = s.c0[w][0]; }
  for (int i = 0; i < nmax; ++i) {
#pragma unroll
    for (int w = 0; w < kCW; ++w) {
      if (s.ntile[w] <= i) continue;
      while (cc[w] >= s.c1[w][cq[w]]) { ++cq[w]; cc[w] = s.c0[w][cq[w]]; }    // (empty parts are skipped)
      const int pair = s.pair[w][cq[w]], r = pair / a.Hkv, g = pair % a.Hkv;
      const int tok0 = cc[w] * kTileTok;
      ++cc[w];
      const int page = a.page_table[(size_t)r * a.max_pages + tok0 / ZB_PAGE_TOKENS];
      const int rowk = ((page * 2 + 0) * a.Hkv + g) * ZB_PAGE_TOKENS + tok0 % ZB_PAGE_TOKENS;
      const int rowv = rowk + a.Hkv * ZB_PAGE_TOKENS;
      ring_acquire(empty_bar, S, gst);
      const int slot = gst % S;
      unsigned char* dst = ring + (size_t)slot * kSlot;
      mbar_expect_tx(&full_bar[slot], (uint32_t)kSlot);
      tma_load_2d(dst, map, 0, rowk, &full_bar[slot]);
      tma_load_2d(dst + 8192, map, 64, rowk, &full_bar[slot]);
      tma_load_2d(dst + 16384, map, 0, rowv, &full_bar[slot]);
      tma_load_2d(dst + 24576, map, 64, rowv, &full_bar[slot]);
      ++gst;
    }
  }
}

// ---- activation producer (warp 2, one thread): the B tiles of one unit ----------------------------------------------
__device__ __forceinline__ void tc_produce_b(const TcArgs& a, const TcGemm& g, const CUtensorMap* map, unsigned char* bring, uint64_t* bfull,
                                             uint64_t* bempty, uint64_t* b_go, int& bst, int& ngo) {
  mbar_wait(b_go, (uint32_t)(ngo & 1));                    // the phase's input is complete (grid barrier passed)
  ++ngo;
  fence_proxy_async_all();
  const Unit u = unit_of(g);
  const int S = a.bstages, n = u.kb1 - u.kb0;
  for (int j = 0; j < n; j += kKbSlot, ++bst) {
    const int cnt = min(kKbSlot, n - j);
    const int slot = bst % S, use = bst / S;
    if (use > 0) mbar_wait(&bempty[slot], (uint32_t)((use - 1) & 1));
    mbar_expect_tx(&bfull[slot], (uint32_t)(cnt * a.Rp * 128));
    for (int e = 0; e < cnt; ++e)
      tma_load_2d(bring + (size_t)slot * a.bslot_bytes + (size_t)e * a.Rp * 128, map, unit_kb(u, j + e) * 64, 0, &bfull[slot]);
  }
}

// ---- MMA issuer: one unit --------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_issue(const TcArgs& a, const TcGemm& g, unsigned char* ring, unsigned char* bring, uint64_t* full_bar,
                                         uint64_t* empty_bar, uint64_t* bfull, uint64_t* bempty, uint64_t* acc_full, volatile int* slot_use, uint32_t tmem,
                                         int& gst, int& bst) {
  const Unit u = unit_of(g);
  if (!u.active) return;
  const int S = a.stages, SB = a.bstages, n = u.kb1 - u.kb0;
  const uint32_t idesc = make_idesc(128, a.Rp);
  const bool lane0 = (threadIdx.x & 31) == 0;
  // the whole warp walks the rings on warp-uniform values, one elected lane issues (tc.cuh: elect_one).  Ring positions
  // advance by increments: a division per slot on this single warp's dependent instruction chain costs as much as the MMAs.
  int slot = gst % S, use = gst / S, bslot = bst % SB, buse = bst / SB;
  const uint32_t ring0 = smem_u32(ring), bring0 = smem_u32(bring);
  for (int j = 0; j < n; j += kKbSlot) {
    const int cnt = min(kKbSlot, n - j);
    mbar_wait(&bfull[bslot], (uint32_t)(buse & 1));
    if (use > 0) {
      for (unsigned spins = 0; slot_use[slot] < use - 1; ++spins)
        if (spins > (1u << 26)) asm volatile("trap;");
    }
    mbar_wait(&full_bar[slot], (uint32_t)(use & 1));
    if (lane0) slot_use[slot] = use;
    tc_fence_after();
    const uint32_t ra = ring0 + (uint32_t)slot * kSlot, rb = bring0 + (uint32_t)bslot * (uint32_t)a.bslot_bytes;
    if (elect_one()) {
      for (int e = 0; e < cnt; ++e) {
        const uint64_t da = make_smem_desc(ra + (uint32_t)e * (kSlot / kKbSlot));
        const uint64_t db = make_smem_desc(rb + (uint32_t)e * (uint32_t)a.Rp * 128u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)                    // UMMA K = 16 bf16 = 32 bytes
          tc_mma(tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (j | e | kk) ? 1u : 0u);
      }
      tc_commit(&empty_bar[slot]);
      tc_commit(&bempty[bslot]);
    }
    __syncwarp();
    if (++slot == S) { slot = 0; ++use; }
    if (++bslot == SB) { bslot = 0; ++buse; }
  }
  gst += (n + kKbSlot - 1) / kKbSlot; bst += (n + kKbSlot - 1) / kKbSlot;
  if (elect_one()) tc_commit(acc_full);
  __syncwarp();
}

// ---- accumulator -> fp32 partials ws[k slice][activation row][feature] -----------------------------------------
__device__ __forceinline__ void tc_dump(const TcArgs& a, const TcGemm& g, int kind, const Unit& u, uint32_t tmem, int cw, int lane) {
  const int quarter = (cw + 3) & 3, half = cw >> 2;           // warp id = cw + 3; a warp may only touch TMEM lanes 32*(warp id % 4) ...
  const int ncol = a.Rp / 2;
  const int lr = quarter * 32 + lane;
  int n; bool ok;
  if (kind == G_FC1) {
    if (lr < g.RBv) { n = u.rb * g.RBv + lr; ok = n < a.F; }
    else { const int i = u.rb * g.RBv + lr - g.RBv; ok = lr < 2 * g.RBv && i < a.F; n = a.F + i; }
  } else {
    n = u.rb * g.RB + lr; ok = lr < g.RB && n < g.N;
  }
  float* wsb = a.ws + (size_t)u.ks * a.R * g.Nw;
  int c0 = 0;
  for (; c0 < ncol; c0 += 8) {
    if (half * ncol + c0 >= a.R) break;
    float v[8];
    tc_ld8(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * ncol + c0), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = half * ncol + c0 + j;
      if (ok && m < a.R) wsb[(size_t)m * g.Nw + n] = v[j];
    }
  }
}

// ---- fc1 without a K split: value * silu(gate) straight from tensor memory ------------------------------------------
// The CTA's accumulator holds value rows (TMEM lanes 0 .. RBv-1) and the gate rows of the SAME features (lanes RBv ..
// 2 RBv-1) for all activation rows (columns).  Gate values cross from their lanes to the value lanes through shared
// memory (the activation ring is idle between two GEMMs); no partials, no grid barrier, no second pass over L2.
__device__ __forceinline__ void tc_epi_silu_direct(const TcArgs& a, const TcGemm& g, const Unit& u, uint32_t tmem, float* sbuf, int cw, int lane) {
  const int quarter = (cw + 3) & 3, half = cw >> 2;
  const int ncol = a.Rp / 2, ld = a.Rp + 1;                   // padded rows: consecutive weight rows hit consecutive banks
  const int lr = quarter * 32 + lane;
  for (int c0 = 0; c0 < ncol; c0 += 8) {
    if (half * ncol + c0 >= a.R) break;                       // warp-uniform
    float v[8];
    tc_ld8(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * ncol + c0), v);
    if (lr < 2 * g.RBv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sbuf[lr * ld + half * ncol + c0 + j] = v[j];
    }
  }
  cbar();
  // (activation row m, feature pair) tasks over all 256 threads: consecutive threads write consecutive features
  const int hp2 = g.RBv / 2, ctid = cw * 32 + lane;
  const int nvalid = min(g.RBv, a.F - u.rb * g.RBv);          // features of this block that exist (last block)
  for (int t = ctid; t < a.R * hp2; t += kCThreads) {
    const int m = t / hp2, i = (t % hp2) * 2;
    if (i >= nvalid) continue;
    float r2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float yv = rbf(sbuf[(i + e) * ld + m]), gg = rbf(sbuf[(g.RBv + i + e) * ld + m]);
      const float sg = rbf(gg / (1.0f + expf(-gg)));                  // F.silu on bf16: fp32 math, bf16 result (fast ex2/rcp: 1 % of the step, not taken)
      r2[e] = __fmul_rn(yv, sg);
    }
    *reinterpret_cast<uint32_t*>(a.h + (size_t)m * a.F + (size_t)u.rb * g.RBv + i) = pack_bf16(r2[0], r2[1]);
  }
}

// ---- compute-warp side of one GEMM phase ------------------------------------------------------------------------
struct CState { int gst, nacc; unsigned epoch; unsigned long long* tl; int ti; };
// debug timeline: stamp slot cs.ti of this CTA's row (only while cs.tl is set: one layer of a step), compute thread 0
#define TL_STAMP(cs, ctid) do { if ((cs).tl && (ctid) == 0 && (cs).ti < 128) (cs).tl[(cs).ti] = gtime_tc(); ++(cs).ti; } while (0)
__device__ __forceinline__ void tc_gemm_compute(const TcArgs& a, int kind, uint64_t* b_go, uint64_t* acc_full, uint32_t tmem, float* gbuf, CState& cs,
                                                int ctid) {
  const TcGemm& g = a.g[kind];
  const Unit u = unit_of(g);
  if (ctid == 0) mbar_arrive(b_go);                           // (after a grid barrier) release the activation producer
  cs.gst += (u.kb1 - u.kb0 + kKbSlot - 1) / kKbSlot;
  TL_STAMP(cs, ctid);
  if (!u.active) { ++cs.ti; return; }
  mbar_wait(acc_full, (uint32_t)(cs.nacc & 1));
  ++cs.nacc;
  tc_fence_after();
  TL_STAMP(cs, ctid);                                         // accumulator complete
  if (kind == G_FC1 && a.fc1_direct) tc_epi_silu_direct(a, g, u, tmem, gbuf, ctid >> 5, ctid & 31);
  else tc_dump(a, g, kind, u, tmem, ctid >> 5, ctid & 31);
  tc_fence_before();
}

// ---- element-wise sub-phase: sum the K-slice partials in slice order, fused epilogue -------------------------------
template <int NB>
__device__ __forceinline__ void sum_partials_batch(const float*& p, size_t stride, float (&v)[8]) {
  float4 lo[NB], hi[NB];
#pragma unroll
  for (int e = 0; e < NB; ++e) {
    lo[e] = __ldcg(reinterpret_cast<const float4*>(p + e * stride));
    hi[e] = __ldcg(reinterpret_cast<const float4*>(p + e * stride + 4));
  }
#pragma unroll
  for (int e = 0; e < NB; ++e) {
    v[0] += lo[e].x; v[1] += lo[e].y; v[2] += lo[e].z; v[3] += lo[e].w; v[4] += hi[e].x; v[5] += hi[e].y; v[6] += hi[e].z; v[7] += hi[e].w;
  }
  p += NB * stride;
}
__device__ __forceinline__ void sum_partials(const TcArgs& a, const TcGemm& g, int m, int n0, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const size_t stride = (size_t)a.R * g.Nw;
  const float* p = a.ws + (size_t)m * g.Nw + n0;
  // batches of 8 / 4 / 2 / 1 slices whose loads are in flight together: the reduce + epilogue sub-phases are chains of L2 round
  // trips on the critical path of every GEMM (9 slices: 2 round trips, 3 with batches of 4); the additions keep the slice
  // order, so the result does not depend on how the loads are batched
  int left = g.nks;
  while (left >= 8) { sum_partials_batch<8>(p, stride, v); left -= 8; }
  if (left >= 4) { sum_partials_batch<4>(p, stride, v); left -= 4; }
  if (left >= 2) { sum_partials_batch<2>(p, stride, v); left -= 2; }
  if (left >= 1) sum_partials_batch<1>(p, stride, v);
}

__device__ __forceinline__ uint4 norm8(const uint4& v, float mean, float rstd, const uint4& gw, const uint4& gb) {
  const uint32_t xv[4] = {v.x, v.y, v.z, v.w}, wv[4] = {gw.x, gw.y, gw.z, gw.w}, bv[4] = {gb.x, gb.y, gb.z, gb.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float lo = (bf16lo(xv[j]) - mean) * rstd * bf16lo(wv[j]) + bf16lo(bv[j]);
    const float hi = (bf16hi(xv[j]) - mean) * rstd * bf16hi(wv[j]) + bf16hi(bv[j]);
    o[j] = pack_bf16(lo, hi);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// norm of the rows this CTA has just written into a.x (rows m0, m0 + dup, ... < R share the statistics) -> a.xn.
// (s, q): this thread's sums over its own elements of the row.  Uniform per CTA (contains CTA barriers).
__device__ __forceinline__ void row_norm_store(const TcArgs& a, int m, int copies, int stride, float s, float q, const bf16* nw, const bf16* nb,
                                               float* red, int cw, int lane, int ctid) {
  // norm parameters of this thread's first chunk: fetched before the barriers (they do not depend on the statistics)
  const int n0p = ctid * 8;
  const bool pre = n0p < a.D;
  const uint4 gw0 = pre ? __ldg(reinterpret_cast<const uint4*>(nw + n0p)) : make_uint4(0, 0, 0, 0);
  const uint4 gb0 = (pre && nb) ? __ldg(reinterpret_cast<const uint4*>(nb + n0p)) : make_uint4(0, 0, 0, 0);
  s = warp_sum(s); q = warp_sum(q);
  if (lane == 0) { red[2 * cw] = s; red[2 * cw + 1] = q; }
  cbar();
  float S = 0.f, Q = 0.f;
#pragma unroll
  for (int w = 0; w < kCW; ++w) { S += red[2 * w]; Q += red[2 * w + 1]; }
  cbar();                                                     // red is reused by the next row
  const float inv = 1.0f / (float)a.D;
  const bool ln = a.norm_kind == ZB_NORM_LAYERNORM;
  const float mu = ln ? S * inv : 0.f;
  const float var = ln ? fmaxf(Q * inv - mu * mu, 0.f) : Q * inv;
  const float rstd = rsqrtf(var + a.eps);
  for (int n0 = ctid * 8; n0 < a.D; n0 += kCThreads * 8) {
    const uint4 xv = *reinterpret_cast<const uint4*>(a.x + (size_t)m * a.D + n0);      // written by this very thread
    const uint4 gw = n0 == n0p ? gw0 : __ldg(reinterpret_cast<const uint4*>(nw + n0));
    const uint4 gb = n0 == n0p ? gb0 : (nb ? __ldg(reinterpret_cast<const uint4*>(nb + n0)) : make_uint4(0, 0, 0, 0));
    const uint4 o = norm8(xv, mu, rstd, gw, gb);
    for (int c = 0; c < copies; ++c) *reinterpret_cast<uint4*>(a.xn + (size_t)(m + c * stride) * a.D + n0) = o;
  }
}

// residual epilogue, one CTA per activation row: x = bf16(x + bf16(linear)), xn = next norm(x)
__device__ __forceinline__ void tc_epi_resid(const TcArgs& a, int kind, const bf16* nw, const bf16* nb, float* red, int cw, int lane, int ctid) {
  const TcGemm& g = a.g[kind];
  for (int m = blockIdx.x; m < a.R; m += gridDim.x) {
    float s = 0.f, q = 0.f;
    for (int n0 = ctid * 8; n0 < a.D; n0 += kCThreads * 8) {
      float v[8];
      bf16* xp = a.x + (size_t)m * a.D + n0;
      const uint4 old = ldcg16(xp);                              // (issued with the partial sums' loads: one round trip, not two)
      sum_partials(a, g, m, n0, v);
      const uint32_t ov[4] = {old.x, old.y, old.z, old.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = pack_bf16(bf16lo(ov[j]) + rbf(v[2 * j]), bf16hi(ov[j]) + rbf(v[2 * j + 1]));
        const float lo = bf16lo(o[j]), hi = bf16hi(o[j]);
        s += lo + hi; q = fmaf(lo, lo, q); q = fmaf(hi, hi, q);
      }
      *reinterpret_cast<uint4*>(xp) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    row_norm_store(a, m, 1, 0, s, q, nw, nb, red, cw, lane, ctid);
  }
}

// embedding sum (sequential bf16 adds, codec_utils.py:37), one CTA per utterance; the CFG halves see the same codes
// (generation_utils.py:191-192); xn = first norm(x)
__device__ __forceinline__ void tc_embed(const TcArgs& a, const bf16* nw, const bf16* nb, float* red, int cw, int lane, int ctid) {
  const long long col = a.loop ? (long long)a.loop->offset : 0;
  const int copies = a.R / a.B;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float s = 0.f, q = 0.f;
    for (int n0 = ctid * 8; n0 < a.D; n0 += kCThreads * 8) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int k = 0; k < a.Q; ++k) {
        long long id = a.delayed[((size_t)b * a.Q + k) * a.T_delayed + col];
        id = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
        const uint4 rw = __ldg(reinterpret_cast<const uint4*>(a.emb[k] + (size_t)id * a.D + n0));
        const uint32_t w[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * i] = rbf(acc[2 * i] + bf16lo(w[i])); acc[2 * i + 1] = rbf(acc[2 * i + 1] + bf16hi(w[i])); }
      }
      uint4 o;
      o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]); o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s += acc[i]; q = fmaf(acc[i], acc[i], q); }
      for (int c = 0; c < copies; ++c) *reinterpret_cast<uint4*>(a.x + (size_t)(b + c * a.B) * a.D + n0) = o;
    }
    row_norm_store(a, b, copies, a.B, s, q, nw, nb, red, cw, lane, ctid);
  }
}

__device__ __forceinline__ void tc_epi(const TcArgs& a, int kind, int epi, bf16* dst, bf16* kv_layer, int cw, int lane) {
  const TcGemm& g = a.g[kind];
  const int nout = (epi == E_SILU) ? a.F : g.N;
  const int nseg = (nout + kSegN - 1) / kSegN;
  const int rows = (epi == E_HEADS && a.cfg_scale != 1.0f) ? a.B : a.R;
  const int ntask = rows * nseg;
  for (int t = cw * gridDim.x + blockIdx.x; t < ntask; t += kCW * gridDim.x) {
    const int m = t / nseg, seg = t % nseg;
    const int n0 = seg * kSegN + lane * 8;
    const bool on = n0 < nout;
    // E_QKV: position, RoPE cos / sin and the KV page depend on the row only: fetched BEFORE the partial sums, whose loads then
    // overlap them (three dependent L2 round trips after the sums otherwise)
    int q_pos = 0, q_page = 0;
    float4 q_c0 = make_float4(1.f, 0.f, 1.f, 0.f), q_c1 = q_c0;
    if (epi == E_QKV) {
      q_pos = a.lengths[m];
      const int qn_ = a.Hq * kHd, kn_ = a.Hkv * kHd;
      if (a.rope_interleaved && on && n0 < qn_ + kn_) {
        const float* rp_ = a.rope + (size_t)min(q_pos, a.rope_len - 1) * kHd + n0 % kHd;
        q_c0 = *reinterpret_cast<const float4*>(rp_); q_c1 = *reinterpret_cast<const float4*>(rp_ + 4);
      }
      if (on && n0 >= qn_) q_page = a.page_table[(size_t)m * a.max_pages + q_pos / ZB_PAGE_TOKENS];
    }
    float v[8];
    if (on) sum_partials(a, g, m, n0, v);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    if (epi == E_STORE) {
      if (on) {
        uint4 o;
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst + (size_t)m * g.N + n0) = o;
      }
    } else if (epi == E_SILU) {
      if (on) {
        float gt[8];
        sum_partials(a, g, m, a.F + n0, gt);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float r2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float yv = rbf(v[2 * j + e]), gg = rbf(gt[2 * j + e]);
            const float sg = rbf(gg / (1.0f + expf(-gg)));            // F.silu on bf16: fp32 math, bf16 result
            r2[e] = __fmul_rn(yv, sg);
          }
          o[j] = pack_bf16(r2[0], r2[1]);
        }
        *reinterpret_cast<uint4*>(a.h + (size_t)m * a.F + n0) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else if (epi == E_QKV) {
      const int qn = a.Hq * kHd, kn = a.Hkv * kHd;
      const int pos = q_pos;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rbf(v[j]);
      const bool rot = on && n0 < qn + kn;
      const int dh = n0 % kHd;                                // first of this lane's 8 features inside its head
      const float* rp = a.rope + (size_t)min(pos, a.rope_len - 1) * kHd;      // [hd/2][2] floats
      if (a.rope_interleaved) {
        if (rot) {                                            // pairs (2i, 2i+1), un-contracted fp32 (_torch.py:57-68)
          const float4 c0 = q_c0, c1 = q_c1;
          const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = o[2 * j], x1 = o[2 * j + 1], c = cs[2 * j], s = cs[2 * j + 1];
            o[2 * j] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
            o[2 * j + 1] = __fadd_rn(__fmul_rn(x1, c), __fmul_rn(x0, s));
          }
        }
      } else {
        // rotate-half: feature i pairs with i + 64 of the same head = lane +- 8 (all lanes take part in the shuffles)
        float other[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) other[j] = __shfl_xor_sync(0xffffffffu, o[j], 8);
        if (rot) {
          const bool first = dh < kHd / 2;
          const int i0 = dh % (kHd / 2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 cs = *reinterpret_cast<const float2*>(rp + 2 * (i0 + j));
            o[j] = first ? __fsub_rn(__fmul_rn(o[j], cs.x), __fmul_rn(other[j], cs.y))
                         : __fadd_rn(__fmul_rn(o[j], cs.x), __fmul_rn(other[j], cs.y));
          }
        }
      }
      if (on) {
        uint4 pk;
        pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]); pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
        if (n0 < qn) {
          *reinterpret_cast<uint4*>(a.q + (size_t)m * qn + n0) = pk;
        } else {
          const int kvsel = n0 < qn + kn ? 0 : 1;
          const int ci = n0 - qn - kvsel * kn;
          const int page = q_page;
          bf16* pb = kv_layer + (((size_t)page * 2 + kvsel) * a.Hkv + ci / kHd) * ZB_PAGE_TOKENS * kHd;
          *reinterpret_cast<uint4*>(pb + (size_t)(pos % ZB_PAGE_TOKENS) * kHd + (ci % kHd)) = pk;
        }
      }
    } else {  // E_HEADS
      if (on) {
        if (a.cfg_scale != 1.0f) {                            // u + (c - u) * s in fp32 (model.py:230-232)
          float uu[8];
          sum_partials(a, g, a.B + m, n0, uu);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j < g.N) {
              const float c = rbf(v[j]), w = rbf(uu[j]);
              a.logits[(size_t)m * a.QV + n0 + j] = __fadd_rn(w, __fmul_rn(__fsub_rn(c, w), a.cfg_scale));
            }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j < g.N) a.logits[(size_t)m * a.QV + n0 + j] = rbf(v[j]);
        }
      }
    }
  }
}

// ---- attention ----------------------------------------------------------------------------------------------------
// Merge the nsplit partials of one (row, kv head) pair in split order and write the attention output (bf16).  Called
// by the warp that delivered the pair's LAST partial (arrival counter), so no grid barrier separates attention and
// merge; the order of the merge is fixed, whoever runs it.  Partial record: [head][o[128], max, sum, pad] fp32.
__device__ __forceinline__ void merge_pair(const TcArgs& a, int pair, int nsplit, int lane) {
  const int G = a.G, r = pair / a.Hkv, g = pair % a.Hkv;
  float M[8], L[8], acc[8][4];
#pragma unroll
  for (int h = 0; h < 8; ++h) { M[h] = -INFINITY; L[h] = 0.f; acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f; }
  for (int sp = 0; sp < nsplit; ++sp) {
    const float* ps = a.attn_part + ((size_t)(pair * nsplit + sp) * G) * kPartStride;
    float ms[8], ls[8]; float4 ov[8];
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < G) {
        ms[h] = __ldcg(ps + h * kPartStride + kHd); ls[h] = __ldcg(ps + h * kPartStride + kHd + 1);
        ov[h] = __ldcg(reinterpret_cast<const float4*>(ps + h * kPartStride + lane * 4));
      }
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < G && ms[h] != -INFINITY) {                      // (a split without tokens delivers max = -inf, sum = 0)
        const float Mn = fmaxf(M[h], ms[h]);
        const float so = __expf(M[h] - Mn), sn = __expf(ms[h] - Mn);
        L[h] = fmaf(ls[h], sn, L[h] * so);
        acc[h][0] = fmaf(ov[h].x, sn, acc[h][0] * so); acc[h][1] = fmaf(ov[h].y, sn, acc[h][1] * so);
        acc[h][2] = fmaf(ov[h].z, sn, acc[h][2] * so); acc[h][3] = fmaf(ov[h].w, sn, acc[h][3] * so);
        M[h] = Mn;
      }
  }
#pragma unroll
  for (int h = 0; h < 8; ++h)
    if (h < G) {
      const float inv = 1.0f / L[h];
      uint2 o;
      o.x = pack_bf16(acc[h][0] * inv, acc[h][1] * inv); o.y = pack_bf16(acc[h][2] * inv, acc[h][3] * inv);
      *reinterpret_cast<uint2*>(a.ay + (size_t)r * a.Hq * kHd + (size_t)(g * G + h) * kHd + lane * 4) = o;
    }
}

// One warp: its parts.  Fragment columns n = 2*(lane%4) + {0,1} are query heads (n < G valid).
__device__ __forceinline__ void tc_attention(const TcArgs& a, const AttnSched& s, const bf16* kv_layer, unsigned char* ring, uint64_t* full_bar,
                                             uint64_t* empty_bar, volatile int* slot_use, int gst_base, int cw, int lane) {
  const int G = a.G, S = a.stages, nsplit = s.nsplit;
  const int lq = lane >> 2, lr = lane & 3;                    // fragment row / column-pair index
  const int n0 = 2 * lr;
  int i = 0;                                                  // tiles consumed by this warp in this phase
  for (int part = 0; part < s.np[cw]; ++part) {
    const int pair = s.pair[cw][part], sidx = s.sidx[cw][part];
    const int r = pair / a.Hkv, g = pair % a.Hkv;
    const int len = a.lengths[r];                             // tokens cached by earlier steps; this step's token sits at position len
    // q fragments (B operand of S^T = K q^T): b0 = q[h = lq][d = 16 ks + 2 lr + {0,1}], b1 = ... + 8
    uint32_t qf[8][2];
    {
      const bf16* qp = a.q + (size_t)r * a.Hq * kHd + (size_t)(g * G + lq) * kHd + 2 * lr;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        qf[ks][0] = lq < G ? __ldcg(reinterpret_cast<const unsigned*>(qp + 16 * ks)) : 0u;
        qf[ks][1] = lq < G ? __ldcg(reinterpret_cast<const unsigned*>(qp + 16 * ks + 8)) : 0u;
      }
    }
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
    for (int c = s.c0[cw][part]; c < s.c1[cw][part]; ++c, ++i) {
      const int seq = gst_base + attn_seq(s, cw, i);
      const int slot = seq % S;
      ring_wait_full(full_bar, slot_use, S, seq, lane == 0);
      const uint32_t kb = smem_u32(ring + (size_t)slot * kSlot), vb = kb + 16384;
      const int nvalid = min(kTileTok, len - c * kTileTok);   // may be <= 0 (empty cache): everything masked
      bool wrote = false;
      if (nvalid < kTileTok) {                                // V rows of tokens that do not exist yet may hold anything: 0 * NaN = NaN
        const int first = max(nvalid, 0);
        for (int idx = lane; idx < (kTileTok - first) * 16; idx += 32) {
          const int row = first + (idx >> 4), ch = idx & 15;
          asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(vb + (uint32_t)((ch >> 3) * 8192 + row * 128 + ((ch & 7) << 4))), "r"(0u) : "memory");
        }
        wrote = true;
        __syncwarp();
      }
      const int lim[2] = {nvalid, nvalid};
      attn_tile64(kb, vb, qf, a.scale, lim, o, mrun, lrun, lane);
      if (wrote) fence_proxy_async_smem();                    // generic writes before the next TMA fill of this slot
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
    if (sidx == nsplit - 1) {
      // last split of the pair: fold in this step's own token (K/V appended by the in_proj epilogue of this launch)
      const int page = a.page_table[(size_t)r * a.max_pages + len / ZB_PAGE_TOKENS];
      const bf16* kp = kv_layer + (((size_t)page * 2 + 0) * a.Hkv + g) * ZB_PAGE_TOKENS * kHd + (size_t)(len % ZB_PAGE_TOKENS) * kHd;
      const bf16* vp = kp + (size_t)a.Hkv * ZB_PAGE_TOKENS * kHd;
      float dot = 0.f;                                        // head lq, this lane's 32 of the 128 dims
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t k0 = __ldcg(reinterpret_cast<const unsigned*>(kp + 16 * ks + 2 * lr));
        const uint32_t k1 = __ldcg(reinterpret_cast<const unsigned*>(kp + 16 * ks + 8 + 2 * lr));
        dot = fmaf(bf16lo(qf[ks][0]), bf16lo(k0), dot); dot = fmaf(bf16hi(qf[ks][0]), bf16hi(k0), dot);
        dot = fmaf(bf16lo(qf[ks][1]), bf16lo(k1), dot); dot = fmaf(bf16hi(qf[ks][1]), bf16hi(k1), dot);
      }
      float vn[8][2];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        vn[dt][0] = __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(vp + 16 * dt + lq)) << 16);
        vn[dt][1] = __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(vp + 16 * dt + lq + 8)) << 16);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);           // lanes 4h .. 4h+3 hold head h's score
      float pn[2], fac[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float sn = __shfl_sync(0xffffffffu, dot, 4 * (n0 + j)) * a.scale;
        const float mn = fmaxf(mrun[j], sn);
        fac[j] = __expf(mrun[j] - mn);
        pn[j] = __expf(sn - mn);
        mrun[j] = mn;
        lrun[j] = lrun[j] * fac[j] + (lq == 0 ? pn[j] : 0.f);  // lrun is a per-lane partial over the fragment rows
      }
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        o[dt][0] = fmaf(pn[0], vn[dt][0], o[dt][0] * fac[0]); o[dt][1] = fmaf(pn[1], vn[dt][0], o[dt][1] * fac[1]);
        o[dt][2] = fmaf(pn[0], vn[dt][1], o[dt][2] * fac[0]); o[dt][3] = fmaf(pn[1], vn[dt][1], o[dt][3] * fac[1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      lrun[j] += __shfl_xor_sync(0xffffffffu, lrun[j], 4);
      lrun[j] += __shfl_xor_sync(0xffffffffu, lrun[j], 8);
      lrun[j] += __shfl_xor_sync(0xffffffffu, lrun[j], 16);
    }
    if (nsplit == 1) {
      // the warp has seen the whole pair: normalise and write the attention output row (bf16)
      bf16* ob = a.ay + (size_t)r * a.Hq * kHd + (size_t)(g * G) * kHd;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = n0 + j;
        if (n < G) {
          const float inv = 1.0f / lrun[j];
          bf16* on = ob + (size_t)n * kHd;
#pragma unroll
          for (int dt = 0; dt < 8; ++dt) { on[16 * dt + lq] = f2bf(o[dt][j] * inv); on[16 * dt + lq + 8] = f2bf(o[dt][2 + j] * inv); }
        }
      }
    } else {
      // publish the partial of this split: [head][o[128], max, sum]; the last arrival merges the pair
      float* out = a.attn_part + ((size_t)(pair * nsplit + sidx) * G) * kPartStride;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = n0 + j;
        if (n < G) {
          float* on = out + (size_t)n * kPartStride;
#pragma unroll
          for (int dt = 0; dt < 8; ++dt) { on[16 * dt + lq] = o[dt][j]; on[16 * dt + lq + 8] = o[dt][2 + j]; }
          if (lq == 0) { on[kHd] = mrun[j]; on[kHd + 1] = lrun[j]; }
        }
      }
      __threadfence();
      __syncwarp();
      unsigned old = 0;
      if (lane == 0) old = atomicAdd(a.pair_cnt + pair, 1u);
      old = __shfl_sync(0xffffffffu, old, 0);
      if ((int)old == nsplit - 1) {
        __threadfence();
        merge_pair(a, pair, nsplit, lane);
        if (lane == 0) a.pair_cnt[pair] = 0u;                 // next use: the next layer's attention, grid barriers away
      }
    }
  }
}

// =====================================================================================================================
__global__ void __launch_bounds__(kTcThreads, 1) decode_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_tc_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots], empty_bar[kMaxSlots], bfull[kMaxBSlots], bempty[kMaxBSlots], b_go, acc_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ volatile int slot_use[kMaxSlots];
  __shared__ AttnSched sched;
  __shared__ float red[2 * kCW];
  if (a.loop && (a.loop->done || a.loop->offset + 1 >= a.T_delayed)) return;     // same answer in every CTA (model.py:471-472)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_tc_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* bring = ring + (size_t)a.stages * kSlot;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); slot_use[s] = -1; }
    for (int s = 0; s < a.bstages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    mbar_init(&b_go, 1); mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  uint32_t ncols = 32;
  while ((int)ncols < a.Rp) ncols <<= 1;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 3) attn_schedule(a, sched, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const TcLayer* layers = a.layers;
  const TcLayer* acts = layers + a.n_layer;                   // map[0] = heads, map[B_*] = activation operands
  const int natt = attn_tiles_cta(sched);                     // attention tiles of this CTA per layer

  if (warp == 0) {
    // ===== TMA producer: the whole step's weights and cached K/V, in consumption order =====
    if (lane == 0) {
      int gst = 0;
      for (int li = 0; li < a.n_layer; ++li) {
        const TcLayer* L = layers + li;
        tc_prefetch(a, a.g[G_OUT], &L->map[M_OUT], G_OUT);
        tc_produce(a, a.g[G_QKV], &L->map[M_IN], G_QKV, ring, full_bar, empty_bar, gst);
        tc_produce_kv(a, sched, &L->map[M_KV], ring, full_bar, empty_bar, gst);
        tc_prefetch(a, a.g[G_FC1], &L->map[M_FC1], G_FC1);
        for (int rep = 0; rep < a.out_proj_repeats; ++rep) tc_produce(a, a.g[G_OUT], &L->map[M_OUT], G_OUT, ring, full_bar, empty_bar, gst);
        tc_prefetch(a, a.g[G_FC2], &L->map[M_FC2], G_FC2);
        tc_produce(a, a.g[G_FC1], &L->map[M_FC1], G_FC1, ring, full_bar, empty_bar, gst);
        if (li + 1 < a.n_layer) tc_prefetch(a, a.g[G_QKV], &L[1].map[M_IN], G_QKV);
        else tc_prefetch(a, a.g[G_HEADS], &acts->map[0], G_HEADS);
        tc_produce(a, a.g[G_FC2], &L->map[M_FC2], G_FC2, ring, full_bar, empty_bar, gst);
      }
      tc_produce(a, a.g[G_HEADS], &acts->map[0], G_HEADS, ring, full_bar, empty_bar, gst);
    }
  } else if (warp_id_uniform() == 1) {
    // ===== MMA issuer (whole warp, see tc_issue) =====
    {
      int gst = 0, bst = 0;
#define TC_ISSUE(kind) tc_issue(a, a.g[kind], ring, bring, full_bar, empty_bar, bfull, bempty, &acc_full, slot_use, tmem, gst, bst)
      for (int li = 0; li < a.n_layer; ++li) {
        TC_ISSUE(G_QKV);
        gst += natt;
        for (int rep = 0; rep < a.out_proj_repeats; ++rep) TC_ISSUE(G_OUT);
        TC_ISSUE(G_FC1);
        TC_ISSUE(G_FC2);
      }
      TC_ISSUE(G_HEADS);
#undef TC_ISSUE
    }
  } else if (warp == 2) {
    // ===== activation producer =====
    if (lane == 0) {
      int bst = 0, ngo = 0;
#define TC_PB(kind, m) tc_produce_b(a, a.g[kind], &acts->map[m], bring, bfull, bempty, &b_go, bst, ngo)
      for (int li = 0; li < a.n_layer; ++li) {
        TC_PB(G_QKV, B_XN);
        int src = B_AY;
        for (int rep = 0; rep < a.out_proj_repeats; ++rep) { TC_PB(G_OUT, src); src = (src == B_Y1) ? B_AY : B_Y1; }
        TC_PB(G_FC1, B_XN);
        TC_PB(G_FC2, B_H);
      }
      TC_PB(G_HEADS, B_XN);
#undef TC_PB
    }
  } else {
    // ===== compute warps =====
    const int ctid = threadIdx.x - 96, cw = ctid >> 5;
    // bar[1] = live steps this session has completed (written by CTA 0 after its last barrier, i.e. after every CTA
    // has read it): the barrier counter bar[0] keeps counting across steps
    const unsigned step0 = *reinterpret_cast<volatile unsigned*>(a.bar + 1);
    CState cs; cs.gst = 0; cs.nacc = 0; cs.epoch = step0 * (unsigned)a.nbar; cs.tl = nullptr; cs.ti = 0;
#define TC_STAMP() TL_STAMP(cs, ctid)
#define TC_GEMM(kind) tc_gemm_compute(a, kind, &b_go, &acc_full, tmem, reinterpret_cast<float*>(bring), cs, ctid)
#define TC_BARRIER() do { TC_STAMP(); grid_barrier(a, cs.epoch, ctid); TC_STAMP(); } while (0)
    tc_embed(a, layers[0].norm_w, layers[0].norm_b, red, cw, lane, ctid);
    TC_BARRIER();
    for (int li = 0; li < a.n_layer; ++li) {
      const TcLayer* L = layers + li;
      cs.tl = (a.timeline && li == a.n_layer / 2) ? a.timeline + (size_t)blockIdx.x * 128 : nullptr;    // one layer, every CTA
      cs.ti = 0;
      // in_proj (input: normed x) -> RoPE, q, paged KV append
      TC_GEMM(G_QKV);
      TC_BARRIER();
      tc_epi(a, G_QKV, E_QKV, nullptr, L->kv_layer, cw, lane);
      TC_BARRIER();
      // attention over the paged cache
      tc_attention(a, sched, L->kv_layer, ring, full_bar, empty_bar, slot_use, cs.gst, cw, lane);
      cs.gst += natt;
      TC_BARRIER();
      // out_proj (twice in the reference, _torch.py:419-420), the last pass adds the residual and applies norm2
      bf16* dst = a.y1;
      for (int rep = 0; rep < a.out_proj_repeats; ++rep) {
        const bool last = rep == a.out_proj_repeats - 1;
        TC_GEMM(G_OUT);
        TC_BARRIER();
        if (last) tc_epi_resid(a, G_OUT, L->norm2_w, L->norm2_b, red, cw, lane, ctid);
        else tc_epi(a, G_OUT, E_STORE, dst, nullptr, cw, lane);
        TC_BARRIER();
        dst = (dst == a.y1) ? a.ay : a.y1;
      }
      // fc1 -> value * silu(gate)
      TC_GEMM(G_FC1);
      TC_BARRIER();
      if (!a.fc1_direct) {
        tc_epi(a, G_FC1, E_SILU, nullptr, nullptr, cw, lane);
        TC_BARRIER();
      } else {
        cs.ti += 2;
      }
      // fc2 + residual, then the next layer's first norm (or the final norm)
      TC_GEMM(G_FC2);
      TC_BARRIER();
      const bool lastl = li + 1 == a.n_layer;
      tc_epi_resid(a, G_FC2, lastl ? a.normf_w : L[1].norm_w, lastl ? a.normf_b : L[1].norm_b, red, cw, lane, ctid);
      TC_BARRIER();
    }
    cs.tl = nullptr;
    // fused heads -> fp32 -> CFG mix
    TC_GEMM(G_HEADS);
    TC_BARRIER();
    tc_epi(a, G_HEADS, E_HEADS, nullptr, nullptr, cw, lane);
    if (cs.epoch != (step0 + 1) * (unsigned)a.nbar) asm volatile("trap;");      // host and kernel disagree on the barrier count
    if (blockIdx.x == 0 && ctid == 0) *reinterpret_cast<volatile unsigned*>(a.bar + 1) = step0 + 1;
#undef TC_STAMP
#undef TC_GEMM
#undef TC_BARRIER
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------ host side ---------------------------------------
inline int env_int_tc(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct TcPlan {
  TcGemm g[G_COUNT];
  int Rp, stages, bstages, bslot_bytes, na, fc1_direct;
  size_t smem, ws_bytes;
};

bool tc_plan(const zb_model_desc& d, int R, int grid, TcPlan* p) {
  memset(p, 0, sizeof(*p));
  const int qn = d.n_heads * d.head_dim, nqkv = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim, QV = d.n_codebooks * d.head_vocab;
  if (d.head_dim != kHd || d.d_model % 64 || d.d_ff % 64 || qn % 64 || R < 2 || R > 128 || (R & 1)) return false;
  const int G = d.n_heads / d.n_heads_kv;
  if (G < 1 || G > 8 || d.n_heads % d.n_heads_kv) return false;
  if (d.n_codebooks > 16 || d.d_model % 8 || R * d.n_heads_kv > 512) return false;
  p->Rp = R <= 16 ? 16 : R <= 32 ? 32 : R <= 64 ? 64 : 128;
  p->bslot_bytes = kKbSlot * p->Rp * 128;
  // activation ring (L2 hits, ~1.6 us under load) next to the weight / KV ring (HBM, ~3 us): 3 x 32 KB + 4 x 32 KB at 128 rows
  p->bstages = std::max(2, std::min(kMaxBSlots, env_int_tc("ZB_TC_BSTAGES", p->Rp >= 128 ? 3 : 4)));
  p->na = std::max(1, std::min(kCW, env_int_tc("ZB_TC_NA", 4)));
  size_t ws = 0;
  auto plan = [&](int kind, int N, int K) -> bool {
    TcGemm& g = p->g[kind];
    g.N = N; g.K = K; g.Nw = round_up(N, 8);
    if (K % 64) return false;
    const int nkb = K / 64;
    if (kind == G_FC1) {
      // value rows | gate rows of the same features in one tile; as many row blocks as there are CTAs when that fits
      const int F = N / 2;
      g.RBv = std::min(64, std::max(8, round_up((F + grid - 1) / grid, 8)));
      g.RB = 2 * g.RBv; g.nrb = (F + g.RBv - 1) / g.RBv;
    } else {
      g.RB = 128; g.nrb = (N + 127) / 128;
    }
    if (g.nrb > grid) return false;
    g.nks = std::max(1, std::min(grid / g.nrb, nkb / 2));
    if (g.nks > nkb) g.nks = nkb;
    ws = std::max(ws, (size_t)g.nks * R * g.Nw * sizeof(float));
    return true;
  };
  if (!plan(G_QKV, nqkv, d.d_model) || !plan(G_OUT, d.d_model, qn) || !plan(G_FC1, 2 * d.d_ff, d.d_model) || !plan(G_FC2, d.d_model, d.d_ff) ||
      !plan(G_HEADS, QV, d.d_model))
    return false;
  if (d.out_proj_repeats < 1 || d.out_proj_repeats > 2 || (d.out_proj_repeats == 2 && qn != d.d_model)) return false;
  p->ws_bytes = ws;
  // fc1 epilogue straight from tensor memory: no K split and the staging [2 RBv][Rp + 1] floats fits the activation ring
  p->fc1_direct = p->g[G_FC1].nks == 1 && (size_t)2 * p->g[G_FC1].RBv * (p->Rp + 1) * 4 <= (size_t)p->bstages * p->bslot_bytes &&
                  env_int_tc("ZB_TC_FC1_DIRECT", 1);
  const size_t fixed = (size_t)p->bstages * p->bslot_bytes + 1024 /* alignment */;
  const size_t avail = (size_t)227 * 1024 - 2048;           // static shared memory: barriers, attention schedule, reduction scratch
  if (avail < fixed + 2 * (size_t)kSlot) return false;
  p->stages = (int)std::min<size_t>(kMaxSlots, (avail - fixed) / kSlot);
  p->stages = std::max(2, std::min(p->stages, env_int_tc("ZB_TC_STAGES", p->stages)));
  p->smem = fixed + (size_t)p->stages * kSlot;
  return true;
}

zb_status tc_map_2d(zb_ctx* ctx, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  ZB_REQUIRE(ctx, enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ZB_REQUIRE(ctx, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] box %u", (int)r, (unsigned long long)rows,
             (unsigned long long)cols, box_rows);
  return ZB_OK;
}

struct TcArena { unsigned* pair_cnt; bf16 *x, *xn, *y1, *q, *ay, *h; float *ws, *attn_part; size_t bytes; };
TcArena tc_arena(const zb_model_desc& d, const TcPlan& p, int R, int grid, void* base) {
  const size_t qn = (size_t)d.n_heads * d.head_dim;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  const int G = d.n_heads / d.n_heads_kv;
  char* q = (char*)base;
  TcArena A;
  A.pair_cnt = (unsigned*)q; q += up(512 * sizeof(unsigned));     // zeroed with the arena at session start
  A.x = (bf16*)q; q += up((size_t)R * d.d_model * 2);
  A.xn = (bf16*)q; q += up((size_t)R * d.d_model * 2);
  A.y1 = (bf16*)q; q += up((size_t)R * d.d_model * 2);
  A.q = (bf16*)q; q += up((size_t)R * qn * 2);
  A.ay = (bf16*)q; q += up((size_t)R * qn * 2);
  A.h = (bf16*)q; q += up((size_t)R * d.d_ff * 2);
  A.ws = (float*)q; q += up(p.ws_bytes);
  A.attn_part = (float*)q; q += up((size_t)std::max(R * d.n_heads_kv, grid * kCW) * kMaxSplit * G * kPartStride * sizeof(float));
  A.bytes = (size_t)(q - (char*)base);
  return A;
}

static unsigned long long* g_tc_timeline = nullptr;

}  // namespace

// ---- interface to api.cu ------------------------------------------------------------------------------------------
bool zb_tc_supported(const zb_model* model, int R) {
  // read per session: tests compare the paths inside one process.  Default: the rows the FFMA2 persistent kernel
  // (decode.cu, R <= 4) does not serve; ZB_DECODE_TC=2 forces this kernel for every supported R, 0 switches it off.
  const int mode = env_int_tc("ZB_DECODE_TC", 1);
  if (!mode || model->n_mamba > 0) return false;
  if (mode == 1 && R <= 4) return false;
  TcPlan p;
  return tc_plan(model->d, R, model->ctx->num_sms, &p);
}

size_t zb_tc_table_bytes(const zb_model* model) { return (size_t)(model->d.n_layer + 1) * sizeof(TcLayer); }

// activation buffers of one generate session (plain bf16 / fp32, exchanged between CTAs through L2)
size_t zb_tc_arena_bytes(const zb_model* model, int R) {
  TcPlan p;
  if (!tc_plan(model->d, R, model->ctx->num_sms, &p)) return 0;
  return tc_arena(model->d, p, R, model->ctx->num_sms, nullptr).bytes + 1024;
}

zb_status zb_tc_table_build(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, int R, void* arena, void* host_buf) {
  const zb_model_desc& d = model->d;
  TcPlan p;
  ZB_REQUIRE(ctx, tc_plan(d, R, ctx->num_sms, &p), "tcgen05 decode step: unsupported shape");
  const size_t page_elems = (size_t)2 * d.n_heads_kv * ZB_PAGE_TOKENS * d.head_dim;
  const int qn = d.n_heads * d.head_dim, nqkv = (d.n_heads + 2 * d.n_heads_kv) * d.head_dim;
  TcLayer* out = (TcLayer*)host_buf;
  memset(out, 0, zb_tc_table_bytes(model));
  for (int li = 0; li < d.n_layer; ++li) {
    const zb_layer& L = model->layers[li];
    ZB_REQUIRE(ctx, L.kind == ZB_LAYER_ATTENTION, "tcgen05 decode step: layer %d is not an attention layer", li);
    TcLayer& t = out[li];
    t.kv_layer = (bf16*)cache->kv_pages + (size_t)model->attn_index[li] * cache->num_pages * page_elems;
    if (zb_status st = tc_map_2d(ctx, &t.map[M_IN], L.in_proj, nqkv, d.d_model, p.g[G_QKV].RB)) return st;
    if (zb_status st = tc_map_2d(ctx, &t.map[M_OUT], L.out_proj, d.d_model, qn, p.g[G_OUT].RB)) return st;
    if (zb_status st = tc_map_2d(ctx, &t.map[M_FC1], L.fc1, 2 * d.d_ff, d.d_model, p.g[G_FC1].RBv)) return st;
    if (zb_status st = tc_map_2d(ctx, &t.map[M_FC2], L.fc2, d.d_model, d.d_ff, p.g[G_FC2].RB)) return st;
    // K/V of one layer as rows of 128 dims: row = ((page * 2 + k|v) * Hkv + kv head) * 64 + token
    if (zb_status st = tc_map_2d(ctx, &t.map[M_KV], t.kv_layer, (uint64_t)cache->num_pages * 2 * d.n_heads_kv * ZB_PAGE_TOKENS, kHd, kTileTok)) return st;
    t.norm_w = (const bf16*)L.norm_w; t.norm_b = (const bf16*)L.norm_b; t.norm2_w = (const bf16*)L.norm2_w; t.norm2_b = (const bf16*)L.norm2_b;
  }
  TcLayer& t = out[d.n_layer];
  const uintptr_t ab = ((uintptr_t)arena + 1023) & ~(uintptr_t)1023;
  const TcArena A = tc_arena(d, p, R, ctx->num_sms, (void*)ab);
  if (zb_status st = tc_map_2d(ctx, &t.map[0], d.heads, d.n_codebooks * d.head_vocab, d.d_model, p.g[G_HEADS].RB)) return st;
  if (zb_status st = tc_map_2d(ctx, &t.map[B_XN], A.xn, R, d.d_model, p.Rp)) return st;
  if (zb_status st = tc_map_2d(ctx, &t.map[B_AY], A.ay, R, qn, p.Rp)) return st;
  if (zb_status st = tc_map_2d(ctx, &t.map[B_Y1], A.y1, R, d.d_model, p.Rp)) return st;
  return tc_map_2d(ctx, &t.map[B_H], A.h, R, d.d_ff, p.Rp);
}

zb_status zb_launch_decode_tc(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* table_dev, unsigned* bar, void* arena, int R,
                              float cfg_scale, float* logits, const int64_t* delayed, int T_delayed, const zb_loop_state* loop, cudaStream_t stream) {
  const zb_model_desc& d = model->d;
  TcPlan p;
  ZB_REQUIRE(ctx, tc_plan(d, R, ctx->num_sms, &p), "tcgen05 decode step: unsupported shape");
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.layers = (const TcLayer*)table_dev; a.n_layer = d.n_layer;
  for (int i = 0; i < G_COUNT; ++i) a.g[i] = p.g[i];
  a.R = R; a.B = cfg_scale != 1.0f ? R / 2 : R; a.Rp = p.Rp;
  a.D = d.d_model; a.F = d.d_ff; a.Hq = d.n_heads; a.Hkv = d.n_heads_kv; a.G = d.n_heads / d.n_heads_kv; a.eps = d.norm_eps; a.norm_kind = d.norm_kind;
  a.rope_interleaved = d.rope_interleaved; a.out_proj_repeats = d.out_proj_repeats;
  a.normf_w = (const bf16*)d.norm_f_w; a.normf_b = (const bf16*)d.norm_f_b; a.QV = d.n_codebooks * d.head_vocab; a.cfg_scale = cfg_scale; a.logits = logits;
  a.rope = d.rope_table; a.rope_len = d.rope_len; a.lengths = cache->lengths; a.page_table = cache->page_table; a.max_pages = cache->max_pages_per_row;
  for (int k = 0; k < d.n_codebooks; ++k) a.emb[k] = (const bf16*)model->emb[k];
  a.Q = d.n_codebooks; a.vocab = d.emb_vocab; a.delayed = delayed; a.T_delayed = T_delayed;
  {
    const uintptr_t ab = ((uintptr_t)arena + 1023) & ~(uintptr_t)1023;
    const TcArena A = tc_arena(d, p, R, ctx->num_sms, (void*)ab);
    a.x = A.x; a.xn = A.xn; a.y1 = A.y1; a.q = A.q; a.ay = A.ay; a.h = A.h; a.ws = A.ws; a.attn_part = A.attn_part; a.pair_cnt = A.pair_cnt;
  }
  a.bar = bar; a.loop = loop; a.fc1_direct = p.fc1_direct; a.l2pf = env_int_tc("ZB_TC_L2PF", 0);   // measured: no gain (the ring, not HBM latency, paces the main loops) and the K/V stream loses L2
  a.nbar = 2 + d.n_layer * (7 + 2 * d.out_proj_repeats - (p.fc1_direct ? 1 : 0));
  a.stages = p.stages; a.bstages = p.bstages; a.bslot_bytes = p.bslot_bytes; a.na = p.na;
  a.scale = 1.0f / sqrtf((float)d.head_dim); a.timeline = g_tc_timeline;
  ZB_CUDA(ctx, zb_ensure_smem(ctx, decode_tc_kernel, p.smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctx->num_sms); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = p.smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;          // every CTA must be resident: the phases meet at grid barriers
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  ZB_CUDA(ctx, cudaLaunchKernelEx(&cfg, decode_tc_kernel, a));
  ctx->launches++;
  return ZB_OK;
}

extern "C" ZB_API zb_status zb_debug_tc_timeline(unsigned long long* dev_buf) { g_tc_timeline = dev_buf; return ZB_OK; }
