// Persistent decode step (device side) - included by decode.cu.
//
// ONE cooperative launch runs embed -> n_layer x (in_proj, attention, out_proj x repeats, fc1, fc2) -> heads for up to
// 4 activation rows (zonos/backbone/_torch.py:307-328, :238; zonos/codec_utils.py:37,68-79; zonos/model.py:229-233).
// 148 CTAs (one per SM) stay resident.  The producer warp of every CTA streams that CTA's slice of ALL the step's
// weight matrices back to back through a shared-memory ring (cp.async.bulk + mbarrier): weights do not depend on
// activations, so HBM keeps streaming across phase boundaries.  Eight consumer warps run the phases.
//
// Three measured facts shape the code (scripts/timeline_mega.py, B200):
//  * Every matrix phase runs through ONE copy of the code (mega_consume, phase parameters are run-time values).  With
//    one inlined template instance per matrix the per-layer code was ~110 KB, every phase started cold in the
//    instruction cache, and a 60-instruction loop took 3 us the first time and 0.3 us when repeated - instruction
//    fetches queue behind the saturated weight stream like any other memory request.
//  * The warps share 4 issue slots, so every 100 instructions per thread cost a phase ~0.1 us: the dot products run on
//    the tensor cores (mma.sync m16n8k16, a few dozen instructions per 32 KB stage instead of ~180 with FFMA).
//  * Activations travel between CTAs as self-validating tagged words instead of through grid barriers (below).
// out_proj is applied twice by the reference (_torch.py:419-420): its slice is held in the ring between the two
// passes, so it is read from HBM once.
#pragma once

struct MegaLayer { const bf16 *norm_w, *norm_b, *in_proj, *out_proj, *norm2_w, *norm2_b, *fc1, *fc2; bf16* kv_layer; };

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
};

// A ring stage holds kMegaRows weight rows (one n8 MMA tile) x one k-block of KB = min(K, kMegaKB) elements.  Every row
// is its own bulk copy and the rows sit (row bytes + 16) apart, so the eight 16-byte row segments one ldmatrix phase
// reads fall into eight different bank groups.
constexpr int kMegaRows = 8, kMegaKB = 2048, kMegaWarpK = 256;
constexpr int kMW = 8;                    // consumer warps (+ 1 producer warp): 9 warps leave 168 registers per thread
constexpr int kMegaStageBytes = kMegaRows * (kMegaKB * 2 + 16), kMegaAttnBytes = 40 * 1024;

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
__device__ __forceinline__ uint2 pack_tagged(const uint4& v) {      // 4 tagged words -> 4 bf16
  return make_uint2((v.x >> 16) | (v.y & 0xffff0000u), (v.z >> 16) | (v.w & 0xffff0000u));
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_relaxed_v2(const uint32_t* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
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
__device__ __forceinline__ uint4 poll_v4(const uint32_t* p, uint32_t tag) {   // spin until the 4 words at p carry `tag`
  uint4 v = ld_relaxed_v4(p);
  for (unsigned spins = 0; !tags_ok(v, tag); ++spins) {
    if (spins > kMegaSpinLimit) asm volatile("trap;");
    v = ld_relaxed_v4(p);
  }
  return v;
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// D[16 x 8] += A[16 x 16] B[16 x 8] (bf16 in, fp32 accumulate): A = activation rows, B = 8 weight rows
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- phases ------------------------------------------------------------------------------------------------------
// kind of a matrix phase = what its epilogue does with the dot products
enum { MP_INPROJ = 0, MP_OUT = 1, MP_OUT_LAST = 2, MP_FC1 = 3, MP_FC2 = 4, MP_HEADS = 5, MP_ATTN = 6 };

// what a thread's in_proj epilogue item needs besides the dot products: position, RoPE cos/sin, KV page.  The same in
// every layer of a step, so it is fetched once per step (three dependent global loads otherwise trail every in_proj)
struct MegaQkvPre { int pos, page; float2 cs; };

__device__ __forceinline__ bool mega_pairs(int kind) { return kind == MP_INPROJ || kind == MP_FC1; }
// output units of a phase (a unit = the weight rows one epilogue item needs: a (value, gate) or RoPE pair, or one row)
__device__ __forceinline__ void mega_unit_rows(const MegaArgs& m, int kind, int u, int& n0, int& n1) {
  n0 = u; n1 = u + 1;
  if (kind == MP_FC1) { n1 = u + m.F; }
  else if (kind == MP_INPROJ) {
    if (!m.rope_interleaved && u < (m.Hq + m.Hkv) * (m.hd / 2)) {      // rotate-half pairs (i, i + hd/2)
      const int half = m.hd / 2;
      n0 = (u / half) * m.hd + (u % half); n1 = n0 + half;
    } else { n0 = 2 * u; n1 = 2 * u + 1; }
  }
}
// this CTA's slice of the N weight rows of a phase: units [u_begin, u_begin + nunits), nrows weight rows
__device__ __forceinline__ void mega_slice(const MegaArgs& m, int kind, int N, int& u_begin, int& nrows) {
  const unsigned nunits = (unsigned)(kind == MP_FC1 ? m.F : (kind == MP_INPROJ ? N / 2 : N));
  u_begin = (int)(blockIdx.x * nunits / gridDim.x);
  const int u_end = (int)((blockIdx.x + 1) * nunits / gridDim.x);
  nrows = (u_end - u_begin) * (mega_pairs(kind) ? 2 : 1);
}
__device__ __forceinline__ int mega_row_of_local(const MegaArgs& m, int kind, int u_begin, int lr) {
  if (!mega_pairs(kind)) return u_begin + lr;
  int n0, n1;
  mega_unit_rows(m, kind, u_begin + (lr >> 1), n0, n1);
  return (lr & 1) ? n1 : n0;
}

// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst): one stage per (group
// of kMegaRows rows, k-block), the k-blocks of a group back to back
__device__ __forceinline__ void mega_produce(const MegaArgs& m, int kind, const bf16* W, int N, int K, unsigned char* ring, uint64_t* full_bar,
                                             uint64_t* empty_bar, int S, int& gst, uint64_t pol, int lane) {
  const int KB = min(K, kMegaKB), NB = K / KB, pitch = KB * 2 + 16;
  int u_begin, nrows;
  mega_slice(m, kind, N, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  int slot = gst % S, parity = ((gst / S) - 1) & 1;
#pragma unroll 1
  for (int gi = 0; gi < ngroup; ++gi) {
    const bf16* src = W + (size_t)mega_row_of_local(m, kind, u_begin, min(gi * kMegaRows + (lane & 7), nrows - 1)) * K;
#pragma unroll 1
    for (int kb = 0; kb < NB; ++kb, ++gst) {
      if (gst >= S) mbar_wait(&empty_bar[slot], parity);
      if (lane == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)(kMegaRows * KB * 2));
      __syncwarp();
      if (lane < kMegaRows) {
        unsigned char* dst = ring + (size_t)slot * kMegaStageBytes + (size_t)lane * pitch;
        if (pol) bulk_g2s(dst, src + (size_t)kb * KB, (uint32_t)KB * 2, &full_bar[slot], pol);
        else bulk_g2s_nohint(dst, src + (size_t)kb * KB, (uint32_t)KB * 2, &full_bar[slot]);
      }
      if (++slot == S) { slot = 0; parity ^= 1; }
    }
  }
}

// consumers: one matrix phase.  Warp w owns the k-slice [w*256, w*256+256) of every k-block (KS = KB/256 warps take
// part).  The A operand of the MMA holds "virtual rows" rho = kb*R + i (k-block kb of activation row i), so the
// fragments of all k-blocks live in the same registers, spread over the lane groups; they are loaded from the tagged
// words directly in fragment layout (lane = 4*rho + c holds k = 16s + 2c, +1 and 16s + 2c + 8, +9 of every 16-k step
// s).  B fragments (8 weight rows x 16 k) come from the ring with ldmatrix; the lane-level reduction is the MMA itself,
// the KS*NB partials of an output meet in `part`.  release = false keeps the slots (out_proj "hold").
template <int R>
__device__ __forceinline__ void mega_consume(const MegaArgs& m, int kind, int N, int K, const bf16* nw, const bf16* nb, int norm_pending,
                                             const uint32_t* xt, int ldx, uint32_t* yt, int ldy, bf16* kv_layer, bool release,
                                             unsigned char* ring, float* part, uint64_t* full_bar, uint64_t* empty_bar, float (*red)[kMW][4],
                                             int S, int& gst, int warp, int lane, uint32_t tag_in, uint32_t tag_out, const MegaQkvPre& qkv_pre,
                                             unsigned long long* stamp) {
  constexpr int NSTEP = kMegaWarpK / 16;                        // 16-k MMA steps per warp and k-block
  constexpr bool kHi = R > 2;                                   // up to 16 virtual rows: rows 8..15 of the A tile in use
  const bool pairs = mega_pairs(kind), has_norm = nw != nullptr;
  const int KB = min(K, kMegaKB), NB = K / KB, pitch = KB * 2 + 16;
  const int KS = KB / kMegaWarpK, KST = KS * NB, VR = R * NB;
  const bool active = warp < KS;
  int u_begin, nrows;
  mega_slice(m, kind, N, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  const int g = lane >> 2, c = lane & 3;
  const bool lo_on = active && g < VR, hi_on = kHi && active && g + 8 < VR;
  const int lo_kb = g / R, lo_i = g % R, hi_kb = (g + 8) / R, hi_i = (g + 8) % R;
  const uint32_t* x_lo = xt + (size_t)lo_i * ldx + (size_t)lo_kb * KB + warp * kMegaWarpK + 2 * c;
  const uint32_t* x_hi = xt + (size_t)hi_i * ldx + (size_t)hi_kb * KB + warp * kMegaWarpK + 2 * c;

  // activations: spin on the operand loads themselves until every word carries the producing phase's tag
  uint32_t afr[NSTEP][kHi ? 4 : 2];
  for (unsigned spins = 0;; ++spins) {
    bool ok = true;
#pragma unroll
    for (int st = 0; st < NSTEP; ++st) {
      uint2 p0 = make_uint2(tag_in, tag_in), p1 = p0;
      if (lo_on) { p0 = ld_relaxed_v2(x_lo + st * 16); p1 = ld_relaxed_v2(x_lo + st * 16 + 8); }
      ok = ok && (((p0.x ^ tag_in) | (p0.y ^ tag_in) | (p1.x ^ tag_in) | (p1.y ^ tag_in)) & 0xffffu) == 0u;
      afr[st][0] = (p0.x >> 16) | (p0.y & 0xffff0000u);
      afr[st][kHi ? 2 : 1] = (p1.x >> 16) | (p1.y & 0xffff0000u);
      if (kHi) {
        uint2 q0 = make_uint2(tag_in, tag_in), q1 = q0;
        if (hi_on) { q0 = ld_relaxed_v2(x_hi + st * 16); q1 = ld_relaxed_v2(x_hi + st * 16 + 8); }
        ok = ok && (((q0.x ^ tag_in) | (q0.y ^ tag_in) | (q1.x ^ tag_in) | (q1.y ^ tag_in)) & 0xffffu) == 0u;
        afr[st][1] = (q0.x >> 16) | (q0.y & 0xffff0000u);
        afr[st][3] = (q1.x >> 16) | (q1.y & 0xffff0000u);
      }
    }
    if (ok) break;
    if (spins > kMegaSpinLimit) asm volatile("trap;");
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();

  if (has_norm) {                                               // LayerNorm / RMSNorm over the row (one k-block: virtual row = row)
    // row statistics: this lane holds 64 elements of row g; the 4 lanes of a row, then the KS warps
    float sacc = 0.f, qacc = 0.f;
#pragma unroll
    for (int st = 0; st < NSTEP; ++st) {
      const float x0 = bf16lo(afr[st][0]), x1 = bf16hi(afr[st][0]), x2 = bf16lo(afr[st][kHi ? 2 : 1]), x3 = bf16hi(afr[st][kHi ? 2 : 1]);
      sacc += (x0 + x1) + (x2 + x3);
      qacc = fmaf(x0, x0, qacc); qacc = fmaf(x1, x1, qacc); qacc = fmaf(x2, x2, qacc); qacc = fmaf(x3, x3, qacc);
    }
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1); qacc += __shfl_xor_sync(0xffffffffu, qacc, 1);
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 2); qacc += __shfl_xor_sync(0xffffffffu, qacc, 2);
    if (lo_on && c == 0) { red[0][warp][g] = sacc; red[1][warp][g] = qacc; }
    // the norm parameters sit in shared memory (copied there a layer ahead: a global load issued here would queue
    // behind the saturated weight stream for microseconds); this thread's copies are complete after the wait, all
    // threads' after the barrier
    if (norm_pending == 0) asm volatile("cp.async.wait_group 0;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");
    if (lo_on) {
      const float inv_k = 1.0f / (float)K;                      // K is a power of two: multiplying is exact
      float tot = 0.f, tsq = 0.f;
      unsigned long long* dbg = (m.timeline && blockIdx.x == 0 && threadIdx.x == 0 && kind == MP_FC1) ? m.timeline + 400 : nullptr;
      if (dbg) dbg[0] = gtime();
#pragma unroll 1
      for (int rep_ = 0; rep_ < 3; ++rep_) {                     // EXPERIMENT
        tot = 0.f; tsq = 0.f;
#pragma unroll 1
        for (int q = 0; q < KS; ++q) { tot += red[0][q][g]; tsq += red[1][q][g]; }
        asm volatile("" ::: "memory");
        if (dbg) dbg[1 + rep_] = gtime();
      }
      {                                                         // EXPERIMENT: different (cold) code of similar size
        float z = 1.f;
#pragma unroll 1
        for (int q = 0; q < KS; ++q) { z = z * red[1][q][g] + red[0][q][g] * 0.5f; z = fminf(z, 3.f); }
        asm volatile("" ::: "memory");
        if (dbg) dbg[4] = gtime() + (z == 12345.f);
      }
      const float mu = tot * inv_k;
      const float mean = (m.norm_kind == ZB_NORM_LAYERNORM) ? mu : 0.f;
      const float var = (m.norm_kind == ZB_NORM_LAYERNORM) ? fmaxf(tsq * inv_k - mu * mu, 0.f) : tsq * inv_k;
      const float rstd = rsqrtf(var + m.eps);
      const uint32_t nwp = smem_u32(nw) + (uint32_t)(warp * kMegaWarpK + 2 * c) * 2;
      const uint32_t nbp = nb ? smem_u32(nb) + (uint32_t)(warp * kMegaWarpK + 2 * c) * 2 : 0u;
#pragma unroll
      for (int st = 0; st < NSTEP; ++st) {
        const uint32_t g0 = lds32(nwp + st * 32), g1 = lds32(nwp + st * 32 + 16);
        const uint32_t b0 = nbp ? lds32(nbp + st * 32) : 0u, b1 = nbp ? lds32(nbp + st * 32 + 16) : 0u;
        const uint32_t w0 = afr[st][0], w1 = afr[st][kHi ? 2 : 1];
        const float y0 = (bf16lo(w0) - mean) * rstd * bf16lo(g0) + bf16lo(b0), y1 = (bf16hi(w0) - mean) * rstd * bf16hi(g0) + bf16hi(b0);
        const float y2 = (bf16lo(w1) - mean) * rstd * bf16lo(g1) + bf16lo(b1), y3 = (bf16hi(w1) - mean) * rstd * bf16hi(g1) + bf16hi(b1);
        afr[st][0] = pack_bf16(y0, y1);
        afr[st][kHi ? 2 : 1] = pack_bf16(y2, y3);
      }
    }
  } else {
    asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");   // everyone has left the previous phase's epilogue: `part` is free
  }

  // ---- operands of this thread's epilogue item (residual value): fetched NOW so the L2 round trip overlaps the
  // weight streaming instead of trailing it.  One output unit x activation row per thread. ----
  const bool cfg = (kind == MP_HEADS && m.cfg_scale != 1.0f);
  const int rows_out = cfg ? m.B : R;
  const int nu = pairs ? nrows >> 1 : nrows;
  const int et = threadIdx.x;
  const bool e_on = et < nu * rows_out;
  const int ej = e_on ? et / rows_out : 0, ei = e_on ? et - ej * rows_out : 0;
  int en0, en1;
  mega_unit_rows(m, kind, u_begin + ej, en0, en1);
  float pre_resid = 0.f;
  if (e_on && (kind == MP_OUT_LAST || kind == MP_FC2)) pre_resid = untag(ld_relaxed_u32(m.xt + (size_t)ei * m.D + en0));   // validated in an earlier phase

  // ---- stream the matrix ----
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(lane & 7) * pitch + (uint32_t)(warp * kMegaWarpK + (lane >> 3) * 8) * 2;
  uint32_t slot = (uint32_t)(gst % S), parity = (uint32_t)((gst / S) & 1);   // advanced incrementally: no division per stage
  uint32_t src = lane_base + slot * kMegaStageBytes, fb = full0 + slot * 8, eb = empty0 + slot * 8;
  // D[virtual row][n = 2c, 2c+1]: partial of weight rows 2c, 2c+1 of the group for k-slice (kb, warp)
  float* dst = part + ((size_t)(2 * c) * KST + warp) * R;
  const int dst_row = KST * R, dst_group = kMegaRows * KST * R;
  gst += ngroup * NB;
#pragma unroll 1
  for (int gi = 0; gi < ngroup; ++gi) {
#pragma unroll 1
    for (int kb = 0; kb < NB; ++kb) {
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      mbar_wait_u32(fb, parity);
      if (active) {
#pragma unroll
        for (int j = 0; j < NSTEP / 2; ++j) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(src + j * 64, b0, b1, b2, b3);
          mma_16816(d0, afr[2 * j][0], kHi ? afr[2 * j][1] : 0u, afr[2 * j][kHi ? 2 : 1], kHi ? afr[2 * j][3] : 0u, b0, b1);
          mma_16816(d1, afr[2 * j + 1][0], kHi ? afr[2 * j + 1][1] : 0u, afr[2 * j + 1][kHi ? 2 : 1], kHi ? afr[2 * j + 1][3] : 0u, b2, b3);
        }
      }
      __syncwarp();
      if (release && lane == 0) mbar_arrive_u32(eb);
      if (++slot == (uint32_t)S) { slot = 0; parity ^= 1u; src = lane_base; fb = full0; eb = empty0; }
      else { src += kMegaStageBytes; fb += 8; eb += 8; }
      // the rows of this k-block are the virtual rows kb*R .. kb*R + R-1
      if (lo_on && lo_kb == kb) { dst[kb * KS * R + lo_i] = d0[0] + d1[0]; dst[kb * KS * R + lo_i + dst_row] = d0[1] + d1[1]; }
      if (kHi && hi_on && hi_kb == kb) { dst[kb * KS * R + hi_i] = d0[2] + d1[2]; dst[kb * KS * R + hi_i + dst_row] = d0[3] + d1[3]; }
    }
    dst += dst_group;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");

  // ---- epilogue.  Rounding points are the reference's: bf16 Linear output first, then the fused op ----
  if (e_on) {
    float v0 = 0.f, v1 = 0.f, u0 = 0.f;
    const float* s0 = part + (size_t)(pairs ? 2 * ej : ej) * KST * R + ei;
#pragma unroll 1
    for (int q = 0; q < KST; ++q) {
      v0 += s0[q * R];
      if (pairs) v1 += s0[(KST + q) * R];
      if (cfg) u0 += s0[q * R + m.B];
    }
    if (kind == MP_OUT_LAST || kind == MP_FC2) {               // residual add (_torch.py:322,326)
      st_relaxed_u32(yt + (size_t)ei * ldy + en0, tag_word(pre_resid + rbf(v0), tag_out));
    } else if (kind == MP_OUT) {
      st_relaxed_u32(yt + (size_t)ei * ldy + en0, tag_word(v0, tag_out));
    } else if (kind == MP_FC1) {                                // value * silu(gate); F.silu on bf16: fp32 math, bf16 result (_torch.py:473-474)
      const float yv = rbf(v0), gt = rbf(v1);
      const float sg = rbf(gt / (1.0f + expf(-gt)));
      st_relaxed_u32(yt + (size_t)ei * ldy + en0, tag_word(__fmul_rn(yv, sg), tag_out));
    } else if (kind == MP_INPROJ) {
      const int qn = m.Hq * m.hd, kn = m.Hkv * m.hd;
      float o0 = rbf(v0), o1 = rbf(v1);
      if (en0 < qn + kn) {                                     // q or k: rotate with separate fp32 mul / sub / add like the reference's eager ops (_torch.py:57-68)
        const float r0 = __fsub_rn(__fmul_rn(o0, qkv_pre.cs.x), __fmul_rn(o1, qkv_pre.cs.y));
        const float r1 = __fadd_rn(__fmul_rn(o1, qkv_pre.cs.x), __fmul_rn(o0, qkv_pre.cs.y));
        o0 = r0; o1 = r1;
      }
      if (en0 < qn) {
        st_relaxed_u32(m.qt + (size_t)ei * qn + en0, tag_word(o0, tag_out));
        st_relaxed_u32(m.qt + (size_t)ei * qn + en1, tag_word(o1, tag_out));
      } else {
        const int kvsel = en0 < qn + kn ? 0 : 1;
        const int c0i = en0 - qn - kvsel * kn, c1i = en1 - qn - kvsel * kn;
        // this step's attention reads the new token from the tagged side buffer; the cache copy is for later steps
        st_relaxed_u32(m.kvt + ((size_t)ei * 2 + kvsel) * kn + c0i, tag_word(o0, tag_out));
        st_relaxed_u32(m.kvt + ((size_t)ei * 2 + kvsel) * kn + c1i, tag_word(o1, tag_out));
        bf16* pb = kv_layer + ((size_t)qkv_pre.page * 2 + kvsel) * m.Hkv * ZB_PAGE_TOKENS * m.hd;
        const int tk = qkv_pre.pos % ZB_PAGE_TOKENS;
        pb[((size_t)(c0i / m.hd) * ZB_PAGE_TOKENS + tk) * m.hd + (c0i % m.hd)] = f2bf(o0);
        pb[((size_t)(c1i / m.hd) * ZB_PAGE_TOKENS + tk) * m.hd + (c1i % m.hd)] = f2bf(o1);
      }
    } else {                                                    // heads: fp32 logits, u + (c - u) * s in fp32 (model.py:230-232)
      if (cfg) {
        const float cv = rbf(v0), uv = rbf(u0);
        m.logits[(size_t)ei * m.QV + en0] = __fadd_rn(uv, __fmul_rn(__fsub_rn(cv, uv), m.cfg_scale));
      } else {
        m.logits[(size_t)ei * m.QV + en0] = rbf(v0);
      }
    }
  }
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
  if (t.n_old > 0) {
    bf16* ks = reinterpret_cast<bf16*>(scratch);
    bf16* vs = ks + kCH * kKStride;
    const bf16* kp = kv_layer + (((size_t)t.page * 2 + 0) * m.Hkv + t.g) * kCH * kHD;
    const bf16* vp = kv_layer + (((size_t)t.page * 2 + 1) * m.Hkv + t.g) * kCH * kHD;
#pragma unroll 1
    for (int c = threadIdx.x; c < kCH * kHD / 8; c += kMW * 32) {
      const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
      if (tok < t.n_old) {
        cp_async16(ks + tok * kKStride + d8, kp + tok * kHD + d8);
        cp_async16(vs + tok * kHD + d8, vp + tok * kHD + d8);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");            // always one group: see cp.async.wait_group 1 in the in_proj norm
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
  // `prefetched`; this step's own token (index kv_len-1) comes from the tagged side buffer
  const int n_old = prefetched ? max(0, min(nk, kv_len - 1 - k0)) : 0;
  const int tok_new = kv_len - 1 - k0;                                  // this step's token, if it falls into this split
#pragma unroll 1
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
    bf16* dst = kvsel ? vs + tok_new * kHD : ks + tok_new * kKStride;
    *reinterpret_cast<uint2*>(dst + lane * 4) = pack_tagged(nv);
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
#pragma unroll 4
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
#pragma unroll 4
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
    float M = -INFINITY;
#pragma unroll 1
    for (int sp = 0; sp < nact; ++sp) M = fmaxf(M, __ldcg(base + (size_t)sp * kPart + kHD));
    float L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int sp = 0; sp < nact; ++sp) {
      const float* ps_ = base + (size_t)sp * kPart;
      const float w = __expf(__ldcg(ps_ + kHD) - M);
      L = fmaf(__ldcg(ps_ + kHD + 1), w, L);
      const float4 ov = __ldcg(reinterpret_cast<const float4*>(ps_ + lane * 4));
      acc[0] = fmaf(ov.x, w, acc[0]); acc[1] = fmaf(ov.y, w, acc[1]);
      acc[2] = fmaf(ov.z, w, acc[2]); acc[3] = fmaf(ov.w, w, acc[3]);
    }
    const float inv = 1.0f / L;
    const uint4 outv = make_uint4(tag_word(acc[0] * inv, tag_out), tag_word(acc[1] * inv, tag_out), tag_word(acc[2] * inv, tag_out),
                                  tag_word(acc[3] * inv, tag_out));
    st_relaxed_v4(m.ayt + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4, outv);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kMW * 32) : "memory");       // scratch free for the next unit
}

template <int R>
__global__ void __launch_bounds__((kMW + 1) * 32, 1) decode_step_kernel(const __grid_constant__ MegaArgs m) {
  extern __shared__ __align__(128) unsigned char smem_m[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ float red[2][kMW][4];
  if (loop_idle(m.loop, m.T_delayed)) return;                 // same answer in every CTA: the loop state only changes in the sampler
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = m.ring_stages;
  unsigned char* ring = smem_m;
  float* part = reinterpret_cast<float*>(smem_m + (size_t)S * kMegaStageBytes);
  unsigned char* attn_scratch = smem_m + (size_t)S * kMegaStageBytes + m.part_bytes;
  bf16* nbuf = reinterpret_cast<bf16*>(attn_scratch + kMegaAttnBytes);   // [2 buffers][weight | bias][D]: norm parameters, copied a layer ahead
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kMW); }
    mbar_fence_init();
  }
  __syncthreads();
  const int qn = m.Hq * m.hd, nqkv = (m.Hq + 2 * m.Hkv) * m.hd;
  const int reps = m.out_proj_repeats, ppl = 4 + reps, nph = 2 + m.n_layer * ppl;

  if (warp == kMW) {
    // ===== producer: the whole step's weights, in consumption order =====
    const uint64_t pol = m.evict_first ? l2_evict_first_policy() : 0ull;
    int gst = 0;
    MegaLayer L = m.layers[0], Lnext = L;
#pragma unroll 1
    for (int li = 0; li < m.n_layer; ++li, L = Lnext) {
      if (li + 1 < m.n_layer) Lnext = m.layers[li + 1];       // the next layer's pointers are on their way while this one streams
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int kind = j == 0 ? MP_INPROJ : (j == 1 ? MP_OUT : (j == 2 ? MP_FC1 : MP_FC2));
        const bf16* W = j == 0 ? L.in_proj : (j == 1 ? L.out_proj : (j == 2 ? L.fc1 : L.fc2));
        const int N = j == 0 ? nqkv : (j == 2 ? 2 * m.F : m.D), K = j == 1 ? qn : (j == 3 ? m.F : m.D);
        mega_produce(m, kind, W, N, K, ring, full_bar, empty_bar, S, gst, pol, lane);
      }
    }
    mega_produce(m, MP_HEADS, m.heads, m.QV, m.D, ring, full_bar, empty_bar, S, gst, pol, lane);
    return;
  }

  // ===== consumers =====
  const unsigned epoch = m.sync[1];                         // written by this session's previous live step
  int gst = 0, gst_out = 0;
  const MegaAttnMeta ameta = mega_attention_meta(m, blockIdx.x, R * m.Hkv * m.nsplit);
  const bool stamping = m.timeline && blockIdx.x == 0;
  if (m.steplog && blockIdx.x == 0 && threadIdx.x == 0 && m.loop) m.steplog[2 * min(m.loop->steps, 4000)] = gtime();
  if (stamping && threadIdx.x == 0) m.timeline[0] = gtime();
  // norm parameters -> shared memory: 16-byte cp.async chunks spread over the consumer threads
  auto norm_prefetch = [&](int buf, const bf16* w, const bf16* b) {
    const int chunks = m.D / 8;
    bf16* dstw = nbuf + (size_t)buf * 2 * m.D;
#pragma unroll 1
    for (int q = threadIdx.x; q < 2 * chunks; q += kMW * 32) {
      const bf16* src = q < chunks ? w + q * 8 : (b ? b + (q - chunks) * 8 : nullptr);
      if (src) cp_async16(dstw + q * 8, src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  MegaLayer L = m.layers[0], Lnext = L;
  norm_prefetch(0, L.norm_w, L.norm_b);
  norm_prefetch(1, L.norm2_w, L.norm2_b);
  // step constants of this thread's in_proj epilogue item (same index math as the epilogue)
  MegaQkvPre qkv_pre; qkv_pre.pos = 0; qkv_pre.page = 0; qkv_pre.cs = make_float2(1.f, 0.f);
  {
    int u_begin, nrows;
    mega_slice(m, MP_INPROJ, nqkv, u_begin, nrows);
    const int et = threadIdx.x;
    if (et < (nrows >> 1) * R) {
      const int ej = et / R, ei = et % R;
      int en0, en1;
      mega_unit_rows(m, MP_INPROJ, u_begin + ej, en0, en1);
      qkv_pre.pos = m.lengths[ei];
      if (en0 < qn + m.Hkv * m.hd) {
        const int ri = m.rope_interleaved ? (en0 % m.hd) / 2 : (en0 % m.hd);
        qkv_pre.cs = *reinterpret_cast<const float2*>(m.rope + ((size_t)min(qkv_pre.pos, m.rope_len - 1) * (m.hd / 2) + ri) * 2);
      }
      if (en0 >= qn) qkv_pre.page = m.page_table[(size_t)ei * m.max_pages + qkv_pre.pos / ZB_PAGE_TOKENS];
    }
  }
  // phase 0: codebook embedding sum (sequential bf16 adds, codec_utils.py:37) for this CTA's columns, both CFG rows
  {
    const uint32_t tag0 = mega_tag(epoch, nph, 0);
    const int d_begin = (int)((long long)blockIdx.x * m.D / gridDim.x), d_end = (int)((long long)(blockIdx.x + 1) * m.D / gridDim.x);
    const long long col = m.loop ? (long long)m.loop->offset : 0;
#pragma unroll 1
    for (int t = threadIdx.x; t < (d_end - d_begin) * m.B; t += kMW * 32) {
      const int b = t / (d_end - d_begin), dd = d_begin + t % (d_end - d_begin);
      float acc = 0.f;
#pragma unroll 1
      for (int k = 0; k < m.Q; ++k) {
        long long id = m.delayed[((size_t)b * m.Q + k) * m.T_delayed + col];
        id = id < 0 ? 0 : (id >= m.vocab ? m.vocab - 1 : id);
        acc = rbf(acc + bf2f(m.emb[k][(size_t)id * m.D + dd]));
      }
      st_relaxed_u32(m.xt + (size_t)b * m.D + dd, tag_word(acc, tag0));
      st_relaxed_u32(m.xt + (size_t)(m.B + b) * m.D + dd, tag_word(acc, tag0));
    }
  }
  if (stamping && threadIdx.x == 0) m.timeline[1] = gtime();

  // phases 1 .. nph-1: per layer in_proj, attention, out_proj x reps, fc1, fc2; then the heads.  One loop body for all.
  int li = 0, j = 0;
#pragma unroll 1
  for (int ph = 1; ph < nph; ++ph) {
    const bool heads = ph == nph - 1;
    if (!heads && j == 0 && li + 1 < m.n_layer) Lnext = m.layers[li + 1];   // pointers of the next layer: loaded a layer before they are needed
    const uint32_t tag_in = mega_tag(epoch, nph, ph - 1), tag_out = mega_tag(epoch, nph, ph);
    unsigned long long* slot = (stamping && 2 * ph + 1 < 512) ? &m.timeline[2 * ph] : nullptr;
    if (!heads && j == 1) {
      // attention over the paged cache
#pragma unroll 1
      for (int unit = blockIdx.x; unit < R * m.Hkv * m.nsplit; unit += gridDim.x)
        mega_attention_unit(m, L.kv_layer, unit, attn_scratch, warp, lane, unit == (int)blockIdx.x ? &ameta : nullptr, tag_in, tag_out,
                            unit == (int)blockIdx.x ? slot : nullptr);
      // buffer 0 is free since the in_proj phase: the next layer's first norm (or the final norm) starts its way to shared memory
      if (li + 1 < m.n_layer) norm_prefetch(0, Lnext.norm_w, Lnext.norm_b); else norm_prefetch(0, m.normf_w, m.normf_b);
    } else {
      // matrix phase: parameters by position in the layer
      int kind, N, K, ldx, ldy = 0, pending = 0;
      const bf16 *nw = nullptr, *nb = nullptr;
      const uint32_t* xt; uint32_t* yt = nullptr;
      bool release = true;
      if (heads) { kind = MP_HEADS; N = m.QV; K = m.D; nw = nbuf; nb = m.normf_b ? nbuf + m.D : nullptr; xt = m.xt; ldx = m.D; }
      else if (j == 0) {
        kind = MP_INPROJ; N = nqkv; K = m.D; nw = nbuf; nb = L.norm_b ? nbuf + m.D : nullptr; xt = m.xt; ldx = m.D; pending = 1;
        mega_attention_prefetch(m, L.kv_layer, ameta, attn_scratch);      // K/V of earlier tokens: in flight while in_proj streams
      } else if (j < 2 + reps) {
        const int r = j - 2;
        const bool last = r == reps - 1;
        kind = last ? MP_OUT_LAST : MP_OUT; N = m.D; K = qn; xt = (r & 1) ? m.y1t : m.ayt; ldx = qn;
        yt = last ? m.xt : ((r & 1) ? m.ayt : m.y1t); ldy = m.D; release = last;
        if (r == 0) gst_out = gst; else gst = gst_out;                    // the held out_proj slice is consumed again
      } else if (j == 2 + reps) {
        kind = MP_FC1; N = 2 * m.F; K = m.D; nw = nbuf + 2 * m.D; nb = L.norm2_b ? nbuf + 3 * m.D : nullptr; xt = m.xt; ldx = m.D; yt = m.ht; ldy = m.F;
      } else { kind = MP_FC2; N = m.D; K = m.F; xt = m.ht; ldx = m.F; yt = m.xt; ldy = m.D; }
      mega_consume<R>(m, kind, N, K, nw, nb, pending, xt, ldx, yt, ldy, L.kv_layer, release, ring, part, full_bar, empty_bar, red, S, gst, warp, lane,
                      tag_in, tag_out, qkv_pre, slot);
      if (kind == MP_FC1 && li + 1 < m.n_layer) norm_prefetch(1, Lnext.norm2_w, Lnext.norm2_b);
    }
    if (slot && threadIdx.x == 0) slot[1] = gtime();
    if (++j == ppl) { j = 0; ++li; L = Lnext; }
  }
  // CTA 0 can only get here after it consumed outputs of every CTA, i.e. after every CTA read the epoch
  if (blockIdx.x == 0 && threadIdx.x == 0) m.sync[1] = epoch + 1;
  if (m.steplog && blockIdx.x == 0 && threadIdx.x == 0 && m.loop) m.steplog[2 * min(m.loop->steps, 4000) + 1] = gtime();
}
