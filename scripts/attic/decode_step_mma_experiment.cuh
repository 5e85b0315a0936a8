// Persistent decode step (device side) - included by decode.cu.
//
// ONE cooperative launch runs embed -> n_layer x (in_proj, attention, out_proj x repeats, fc1, fc2) -> heads for up to
// 4 activation rows (zonos/backbone/_torch.py:307-328, :238; zonos/codec_utils.py:37,68-79; zonos/model.py:229-233).
// 148 CTAs (one per SM) stay resident.  The producer warp of every CTA streams that CTA's slice of ALL the step's
// weight matrices back to back through a shared-memory ring (cp.async.bulk + mbarrier) - weights do not depend on
// activations, so HBM keeps streaming across phase boundaries.  The 16 consumer warps run every matrix phase through
// ONE code path (mega_consume, phase descriptor filled at run time): the whole per-layer loop is a few thousand
// instructions and stays resident in the instruction cache - with one inlined copy per matrix the kernel spent
// microseconds per phase waiting for instruction fetches behind the saturated weight stream.
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
  int ring_stages, part_bytes, stage_bytes_unused;
  unsigned long long* timeline;   // debug: globaltimer stamps of CTA 0 (2 per phase: inputs ready, work done)
};

// A ring stage holds kMegaRows weight rows (one n8 MMA tile) x one k-block of min(K, kMegaKB) elements; every row is
// its own bulk copy and the rows sit (row bytes + 16) apart, so the eight 16-byte row segments one ldmatrix phase
// reads fall into eight different bank groups.
constexpr int kMegaRows = 8, kMegaKB = 2048, kMegaWarpK = 128, kMegaMaxNB = 4, kMegaMaxG = 2;
constexpr int kMegaStageBytes = kMegaRows * (kMegaKB * 2 + 16);

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
// D[16 x 8] += A[16 x 16] B[16 x 8]: A = activation rows (only rows 0..R-1 are non-zero: a1 = a3 = 0), B = 8 weight rows
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %5}, {%7, %8}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}

// ---- phases ------------------------------------------------------------------------------------------------------
enum { MP_INPROJ = 0, MP_ATTN = 1, MP_OUT = 2, MP_OUT_LAST = 3, MP_FC1 = 4, MP_FC2 = 5, MP_HEADS = 6 };

struct MegaPhase {
  int kind;
  const bf16* W; int N, K;              // weight [N, K]
  const bf16 *nw, *nb;                  // norm in front of the matrix (INPROJ, FC1, HEADS)
  const uint32_t* xt; int ldx;          // tagged input rows
  uint32_t* yt; int ldy;                // tagged output rows (OUT*, FC1, FC2)
  bf16* kv_layer;                       // INPROJ / ATTN
  bool streams;                         // the producer streams W for this phase (false: attention, held out_proj pass)
  bool first_out, release;              // out_proj hold: the first pass records the ring position, only the last releases
};

__device__ __forceinline__ int mega_num_phases(const MegaArgs& m) { return 2 + m.n_layer * (4 + m.out_proj_repeats); }

// phase ph = 1 .. nph-1 (phase 0 is the embedding sum)
__device__ __forceinline__ void mega_phase(const MegaArgs& m, int ph, int nph, MegaPhase& p) {
  const int qn = m.Hq * m.hd;
  p.nw = nullptr; p.nb = nullptr; p.yt = nullptr; p.ldy = 0; p.kv_layer = nullptr; p.streams = true; p.first_out = false; p.release = true;
  if (ph == nph - 1) {
    p.kind = MP_HEADS; p.W = m.heads; p.N = m.QV; p.K = m.D; p.nw = m.normf_w; p.nb = m.normf_b; p.xt = m.xt; p.ldx = m.D;
    return;
  }
  const int ppl = 4 + m.out_proj_repeats, li = (ph - 1) / ppl, j = (ph - 1) % ppl;
  const MegaLayer& L = m.layers[li];
  p.kv_layer = L.kv_layer;
  if (j == 0) {
    p.kind = MP_INPROJ; p.W = L.in_proj; p.N = (m.Hq + 2 * m.Hkv) * m.hd; p.K = m.D; p.nw = L.norm_w; p.nb = L.norm_b; p.xt = m.xt; p.ldx = m.D;
  } else if (j == 1) {
    p.kind = MP_ATTN; p.W = nullptr; p.N = 0; p.K = 0; p.xt = m.qt; p.ldx = qn; p.streams = false;
  } else if (j < 2 + m.out_proj_repeats) {
    const int r = j - 2;
    const bool last = r == m.out_proj_repeats - 1;
    p.kind = last ? MP_OUT_LAST : MP_OUT; p.W = L.out_proj; p.N = m.D; p.K = qn;
    p.xt = (r & 1) ? m.y1t : m.ayt; p.ldx = qn;
    p.yt = last ? m.xt : ((r & 1) ? m.ayt : m.y1t); p.ldy = m.D;
    p.streams = r == 0; p.first_out = r == 0; p.release = last;
  } else if (j == 2 + m.out_proj_repeats) {
    p.kind = MP_FC1; p.W = L.fc1; p.N = 2 * m.F; p.K = m.D; p.nw = L.norm2_w; p.nb = L.norm2_b; p.xt = m.xt; p.ldx = m.D; p.yt = m.ht; p.ldy = m.F;
  } else {
    p.kind = MP_FC2; p.W = L.fc2; p.N = m.D; p.K = m.F; p.xt = m.ht; p.ldx = m.F; p.yt = m.xt; p.ldy = m.D;
  }
}

// output units of a phase (a unit = the weight rows one epilogue item needs: a (value, gate) or RoPE pair, or one row)
__device__ __forceinline__ bool mega_pairs(int kind) { return kind == MP_INPROJ || kind == MP_FC1; }
__device__ __forceinline__ int mega_units(const MegaArgs& m, const MegaPhase& p) {
  return p.kind == MP_FC1 ? m.F : (p.kind == MP_INPROJ ? p.N / 2 : p.N);
}
__device__ __forceinline__ void mega_unit_rows(const MegaArgs& m, int kind, int u, int& n0, int& n1) {
  if (kind == MP_FC1) { n0 = u; n1 = u + m.F; }
  else if (kind == MP_INPROJ) {
    if (!m.rope_interleaved && u < (m.Hq + m.Hkv) * (m.hd / 2)) {      // rotate-half pairs (i, i + hd/2)
      const int half = m.hd / 2;
      n0 = (u / half) * m.hd + (u % half); n1 = n0 + half;
    } else { n0 = 2 * u; n1 = 2 * u + 1; }
  } else { n0 = u; n1 = u + 1; }
}
__device__ __forceinline__ void mega_slice(const MegaArgs& m, const MegaPhase& p, int& u_begin, int& nrows) {
  const int nunits = mega_units(m, p);
  u_begin = (int)((long long)blockIdx.x * nunits / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * nunits / gridDim.x);
  nrows = (u_end - u_begin) * (mega_pairs(p.kind) ? 2 : 1);
}
__device__ __forceinline__ int mega_row_of_local(const MegaArgs& m, int kind, int u_begin, int lr) {
  if (mega_pairs(kind)) {
    int n0, n1;
    mega_unit_rows(m, kind, u_begin + (lr >> 1), n0, n1);
    return (lr & 1) ? n1 : n0;
  }
  return u_begin + lr;
}

// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst): one stage per
// (k-block, group of kMegaRows rows), k-block outermost
__device__ __forceinline__ void mega_produce(const MegaArgs& m, const MegaPhase& p, unsigned char* ring, uint64_t* full_bar, uint64_t* empty_bar,
                                             int S, int& gst, int lane) {
  const int K = p.K, KB = min(K, kMegaKB), NB = K / KB, pitch = KB * 2 + 16;
  int u_begin, nrows;
  mega_slice(m, p, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  int slot = gst % S, parity = ((gst / S) - 1) & 1;
  for (int kb = 0; kb < NB; ++kb)
    for (int gi = 0; gi < ngroup; ++gi, ++gst) {
      if (gst >= S) mbar_wait(&empty_bar[slot], parity);
      if (lane == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)(kMegaRows * KB * 2));
      __syncwarp();
      if (lane < kMegaRows) {
        const bf16* src = p.W + (size_t)mega_row_of_local(m, p.kind, u_begin, min(gi * kMegaRows + lane, nrows - 1)) * K + (size_t)kb * KB;
        bulk_g2s_nohint(ring + (size_t)slot * kMegaStageBytes + (size_t)lane * pitch, src, (uint32_t)KB * 2, &full_bar[slot]);
      }
      if (++slot == S) { slot = 0; parity ^= 1; }
    }
}

// consumers: one matrix phase on the tensor cores (mma.sync m16n8k16, fp32 accumulate).  Warp w owns the k-slice
// [w*128, w*128+128) of every k-block: its A fragments (the activation rows, bf16) stay in registers while the block
// streams, B fragments (8 weight rows x 16 k) come from the ring with ldmatrix, and the tile's lane-level reduction is
// the MMA itself; the KS = KB/128 warp partials of each output meet in `part`.
template <int R>
__device__ __forceinline__ void mega_consume(const MegaArgs& m, const MegaPhase& p, unsigned char* ring, float* part, bf16* staging,
                                             uint64_t* full_bar, uint64_t* empty_bar, float (*red)[kW3][4], int S, int& gst, int warp, int lane,
                                             uint32_t tag_in, uint32_t tag_out, unsigned long long* stamp) {
  constexpr int NSTEP = kMegaWarpK / 16;
  const int kind = p.kind;
  const bool pairs = mega_pairs(kind), has_norm = p.nw != nullptr;
  const int K = p.K, KB = min(K, kMegaKB), NB = K / KB, pitch = KB * 2 + 16;
  const int KS = KB / kMegaWarpK;                               // warps that take part (all 16 for KB = 2048)
  const bool active = warp < KS;
  int u_begin, nrows;
  mega_slice(m, p, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  const int kw = warp * kMegaWarpK + lane * 4;                  // this lane's 4 elements of the warp's slice, per k-block
  const int g = lane >> 2, c = lane & 3;
  bf16* stg = staging + (size_t)warp * R * kMegaWarpK;          // per-warp fragment staging
  uint2 nwr = make_uint2(0, 0), nbr = make_uint2(0, 0);          // norm parameters: in flight together with the activations
  if (has_norm && active) {
    nwr = *reinterpret_cast<const uint2*>(p.nw + kw);
    if (p.nb) nbr = *reinterpret_cast<const uint2*>(p.nb + kw);
  }

  // activations: spin on the operand loads themselves until every word carries the producing phase's tag
  uint2 xp[kMegaMaxNB][R];                                      // this lane's 4 elements per (k-block, row) as bf16 pairs
  if (active) {
    for (unsigned spins = 0;; ++spins) {
      bool ok = true;
#pragma unroll
      for (int kb = 0; kb < kMegaMaxNB; ++kb)
        if (kb < NB) {
#pragma unroll
          for (int i = 0; i < R; ++i) {
            const uint4 v = ld_relaxed_v4(p.xt + (size_t)i * p.ldx + (size_t)kb * KB + kw);
            ok = ok && tags_ok(v, tag_in);
            xp[kb][i] = pack_tagged(v);
          }
        }
      if (ok) break;
      if (spins > kMegaSpinLimit) asm volatile("trap;");
    }
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();

  if (has_norm) {                                               // LayerNorm / RMSNorm over the row (K == KB here)
    float xf[R][4], mean[R], rstd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { xf[i][0] = bf16lo(xp[0][i].x); xf[i][1] = bf16hi(xp[0][i].x); xf[i][2] = bf16lo(xp[0][i].y); xf[i][3] = bf16hi(xp[0][i].y); }
    if (active) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        float sacc = 0.f, qacc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { sacc += xf[i][e]; qacc = fmaf(xf[i][e], xf[i][e], qacc); }
        sacc = warp_sum(sacc);
        qacc = warp_sum(qacc);
        if (lane == 0) { red[0][warp][i] = sacc; red[1][warp][i] = qacc; }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
    const float inv_k = 1.0f / (float)K;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float tot = 0.f, tsq = 0.f;
      for (int q = 0; q < KS; ++q) { tot += red[0][q][i]; tsq += red[1][q][i]; }
      const float mu = tot * inv_k;
      mean[i] = (m.norm_kind == ZB_NORM_LAYERNORM) ? mu : 0.f;
      const float var = (m.norm_kind == ZB_NORM_LAYERNORM) ? fmaxf(tsq * inv_k - mu * mu, 0.f) : tsq * inv_k;
      rstd[i] = rsqrtf(var + m.eps);
    }
    const float g4[4] = {bf16lo(nwr.x), bf16hi(nwr.x), bf16lo(nwr.y), bf16hi(nwr.y)};
    const float b4[4] = {bf16lo(nbr.x), bf16hi(nbr.x), bf16lo(nbr.y), bf16hi(nbr.y)};
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float y4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) y4[e] = (xf[i][e] - mean[i]) * rstd[i] * g4[e] + b4[e];
      xp[0][i] = make_uint2(pack_bf16(y4[0], y4[1]), pack_bf16(y4[2], y4[3]));
    }
  } else {
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // everyone has left the previous phase's epilogue: `part` is free
  }

  // ---- operands of this thread's epilogue (residual value, RoPE cos/sin, KV page): fetched NOW so their L2 round
  // trips overlap the weight streaming instead of trailing it ----
  const bool cfg = (kind == MP_HEADS && m.cfg_scale != 1.0f);
  const int rows_out = cfg ? m.B : R;
  const int nu = pairs ? nrows / 2 : nrows;
  const int et = threadIdx.x;                                 // one epilogue item per thread (host guarantees nu*rows_out <= 512)
  const bool e_on = et < nu * rows_out;
  const int ej = e_on ? et / rows_out : 0, ei = e_on ? et % rows_out : 0;
  int en0, en1;
  mega_unit_rows(m, kind, u_begin + ej, en0, en1);
  const int qn = m.Hq * m.hd, kn = m.Hkv * m.hd;
  float pre_resid = 0.f;
  float2 pre_cs = make_float2(1.f, 0.f);
  int pre_pos = 0, pre_page = 0;
  if (e_on) {
    if (kind == MP_OUT_LAST || kind == MP_FC2) pre_resid = untag(ld_relaxed_u32(m.xt + (size_t)ei * m.D + en0));   // validated in an earlier phase
    if (kind == MP_INPROJ) {
      pre_pos = m.lengths[ei];
      if (en0 < qn + kn) {
        const int ri = m.rope_interleaved ? (en0 % m.hd) / 2 : (en0 % m.hd);
        pre_cs = *reinterpret_cast<const float2*>(m.rope + ((size_t)min(pre_pos, m.rope_len - 1) * (m.hd / 2) + ri) * 2);
      }
      if (en0 >= qn) pre_page = m.page_table[(size_t)ei * m.max_pages + pre_pos / ZB_PAGE_TOKENS];
    }
  }

  // ---- stream the matrix ----
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(lane & 7) * pitch + (uint32_t)(warp * kMegaWarpK + (lane >> 3) * 8) * 2;
  uint32_t slot = (uint32_t)(gst % S), parity = (uint32_t)((gst / S) & 1);   // advance incrementally: no division in the stage loop
  uint32_t src = lane_base + slot * kMegaStageBytes, fb = full0 + slot * 8, eb = empty0 + slot * 8;
  const bool storer = active && g < R, release = p.release;
  float* dst = part + ((size_t)(2 * c) * KS + warp) * R + g;   // D[row g][n = 2c, 2c+1] of the current row group
  const int dst_row = KS * R, dst_group = kMegaRows * KS * R;
  float d0[kMegaMaxG][4], d1[kMegaMaxG][4];                     // NB > 1: the (<= kMegaMaxG) row groups accumulate over the k-blocks
#pragma unroll
  for (int q = 0; q < kMegaMaxG; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) { d0[q][e] = 0.f; d1[q][e] = 0.f; }
  for (int kb = 0; kb < NB; ++kb) {
    // A fragments of this k-block: through the staging area from "4 consecutive k per lane" to the MMA layout
    // (lane = 4*row + c holds k = 16s + 2c, +1 and 16s + 2c + 8, +9 of every 16-k step s)
    uint32_t afr[NSTEP][2];
    if (active) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        uint2 v = xp[0][i];
#pragma unroll
        for (int q = 1; q < kMegaMaxNB; ++q) if (kb == q) v = xp[q][i];
        *reinterpret_cast<uint2*>(stg + i * kMegaWarpK + lane * 4) = v;
      }
    }
    __syncwarp();
#pragma unroll
    for (int st = 0; st < NSTEP; ++st) {
      afr[st][0] = storer ? *reinterpret_cast<const uint32_t*>(stg + g * kMegaWarpK + st * 16 + 2 * c) : 0u;
      afr[st][1] = storer ? *reinterpret_cast<const uint32_t*>(stg + g * kMegaWarpK + st * 16 + 2 * c + 8) : 0u;
    }
    __syncwarp();
    for (int gi = 0; gi < ngroup; ++gi) {
      const int q = (NB > 1) ? gi : 0;
      float e0[4], e1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { e0[e] = (q == 1) ? d0[1][e] : d0[0][e]; e1[e] = (q == 1) ? d1[1][e] : d1[0][e]; }
      mbar_wait_u32(fb, parity);
      if (active) {
#pragma unroll
        for (int j = 0; j < NSTEP / 2; ++j) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(src + j * 64, b0, b1, b2, b3);
          mma_16816(e0, afr[2 * j][0], afr[2 * j][1], b0, b1);
          mma_16816(e1, afr[2 * j + 1][0], afr[2 * j + 1][1], b2, b3);
        }
      }
      __syncwarp();
      if (release && lane == 0) mbar_arrive_u32(eb);
      if (++slot == (uint32_t)S) { slot = 0; parity ^= 1u; src = lane_base; fb = full0; eb = empty0; }
      else { src += kMegaStageBytes; fb += 8; eb += 8; }
      if (NB == 1) {                                            // the group is complete after its only stage
        if (storer) { dst[0] = e0[0] + e1[0]; dst[dst_row] = e0[1] + e1[1]; }
        dst += dst_group;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (q == 1) { d0[1][e] = e0[e]; d1[1][e] = e1[e]; } else { d0[0][e] = e0[e]; d1[0][e] = e1[e]; }
        }
      }
    }
  }
  if (NB > 1 && storer) {
#pragma unroll
    for (int q = 0; q < kMegaMaxG; ++q)
      if (q < ngroup) { dst[q * dst_group] = d0[q][0] + d1[q][0]; dst[q * dst_group + dst_row] = d0[q][1] + d1[q][1]; }
  }
  gst += ngroup * NB;
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");

  // ---- epilogue: one output unit x activation row per thread.  Rounding points are the reference's: bf16 Linear
  // output first, then the fused op ----
  if (e_on) {
    float v0 = 0.f, v1 = 0.f, u0 = 0.f;
    const float* s0 = part + (size_t)(pairs ? 2 * ej : ej) * KS * R;
    for (int q = 0; q < KS; ++q) v0 += s0[q * R + ei];
    if (pairs) { const float* s1 = s0 + (size_t)KS * R; for (int q = 0; q < KS; ++q) v1 += s1[q * R + ei]; }
    if (cfg) for (int q = 0; q < KS; ++q) u0 += s0[q * R + m.B + ei];
    if (kind == MP_OUT_LAST || kind == MP_FC2) {               // residual add (_torch.py:322,326)
      st_relaxed_u32(p.yt + (size_t)ei * p.ldy + en0, tag_word(pre_resid + rbf(v0), tag_out));
    } else if (kind == MP_OUT) {
      st_relaxed_u32(p.yt + (size_t)ei * p.ldy + en0, tag_word(v0, tag_out));
    } else if (kind == MP_FC1) {                                // value * silu(gate), F.silu on bf16: fp32 math, bf16 result (_torch.py:473-474)
      const float yv = rbf(v0), gt = rbf(v1);
      const float sg = rbf(gt / (1.0f + expf(-gt)));
      st_relaxed_u32(p.yt + (size_t)ei * p.ldy + en0, tag_word(__fmul_rn(yv, sg), tag_out));
    } else if (kind == MP_INPROJ) {
      float o0 = rbf(v0), o1 = rbf(v1);
      if (en0 < qn + kn) {                                     // q or k: rotate with separate fp32 mul / sub / add like the reference's eager ops (_torch.py:57-68)
        const float r0 = __fsub_rn(__fmul_rn(o0, pre_cs.x), __fmul_rn(o1, pre_cs.y));
        const float r1 = __fadd_rn(__fmul_rn(o1, pre_cs.x), __fmul_rn(o0, pre_cs.y));
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
        bf16* pb = p.kv_layer + ((size_t)pre_page * 2 + kvsel) * m.Hkv * ZB_PAGE_TOKENS * m.hd;
        const int tk = pre_pos % ZB_PAGE_TOKENS;
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
struct MegaAttnMeta { int n_old, page, g; };                  // step constants of this CTA's first attention unit
__device__ __forceinline__ MegaAttnMeta mega_attention_meta(const MegaArgs& m, int unit, int nunits) {
  MegaAttnMeta t; t.n_old = 0; t.page = 0; t.g = 0;
  if (unit >= nunits) return t;
  const int split = unit % m.nsplit, r = unit / (m.nsplit * m.Hkv);
  t.g = (unit / m.nsplit) % m.Hkv;
  const int kv_len = m.lengths[r] + 1;
  t.n_old = max(0, min(kCH, kv_len - 1 - split * kCH));
  if (t.n_old > 0) t.page = m.page_table[(size_t)r * m.max_pages + split];
  return t;
}
__device__ __forceinline__ void mega_attention_prefetch(const MegaArgs& m, const bf16* kv_layer, const MegaAttnMeta& t, unsigned char* scratch) {
  if (t.n_old <= 0) return;
  bf16* ks = reinterpret_cast<bf16*>(scratch);
  bf16* vs = ks + kCH * kKStride;
  const bf16* kp = kv_layer + (((size_t)t.page * 2 + 0) * m.Hkv + t.g) * kCH * kHD;
  const bf16* vp = kv_layer + (((size_t)t.page * 2 + 1) * m.Hkv + t.g) * kCH * kHD;
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += kW3 * 32) {
    const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
    if (tok < t.n_old) {
      cp_async16(ks + tok * kKStride + d8, kp + tok * kHD + d8);
      cp_async16(vs + tok * kHD + d8, vp + tok * kHD + d8);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void mega_attention_unit(const MegaArgs& m, const bf16* kv_layer, int unit, unsigned char* scratch, int warp, int lane,
                                                    bool prefetched, uint32_t tag_in, uint32_t tag_out, unsigned long long* stamp) {
  bf16* ks = reinterpret_cast<bf16*>(scratch);                         // [64][136]
  bf16* vs = ks + kCH * kKStride;                                      // [64][128]
  float* qs = reinterpret_cast<float*>(vs + kCH * kHD);                // [8][128]
  float* ps = qs + 8 * kHD;                                            // [8][64]
  int* s_last = reinterpret_cast<int*>(ps + 8 * kCH);
  const int G = m.Hq / m.Hkv;
  const int split = unit % m.nsplit, g = (unit / m.nsplit) % m.Hkv, r = unit / (m.nsplit * m.Hkv);
  const int kv_len = m.lengths[r] + 1;
  const int nact = (kv_len + kCH - 1) / kCH;
  if (split >= nact) return;                                            // uniform for the CTA
  const int k0 = split * kCH, nk = min(kCH, kv_len - k0);
  const int page = m.page_table[(size_t)r * m.max_pages + split];
  const bf16* kp = kv_layer + (((size_t)page * 2 + 0) * m.Hkv + g) * kCH * kHD;
  const bf16* vp = kv_layer + (((size_t)page * 2 + 1) * m.Hkv + g) * kCH * kHD;
  // tokens cached by earlier steps were prefetched (mega_attention_prefetch, before the in_proj phase) when
  // `prefetched`; this step's own token (index kv_len-1) comes from the tagged side buffer
  const int n_old = prefetched ? max(0, min(nk, kv_len - 1 - k0)) : 0;
  const int tok_new = kv_len - 1 - k0;                                  // this step's token, if it falls into this split
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += kW3 * 32) {
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
  } else if (warp >= 8 && warp < 10 && tok_new >= 0 && tok_new < nk) {  // K (warp 8) and V (warp 9) of this step's token
    const int kvsel = warp - 8;
    const uint4 nv = poll_v4(m.kvt + ((size_t)r * 2 + kvsel) * m.Hkv * kHD + (size_t)g * kHD + lane * 4, tag_in);
    bf16* dst = kvsel ? vs + tok_new * kHD : ks + tok_new * kKStride;
    *reinterpret_cast<uint2*>(dst + lane * 4) = pack_tagged(nv);
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
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
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
  if (threadIdx.x == 0) {
    // one acq_rel RMW publishes this CTA's partials (cumulative over the CTA barrier) and acquires the others'
    int32_t* cnt = m.attn_counters + (size_t)r * m.Hkv + g;
    int prev;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(prev) : "l"(cnt) : "memory");
    *s_last = (prev == nact - 1);
    if (*s_last) *cnt = 0;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
  if (*s_last && warp < G) {
    const float* base = m.attn_part + (((size_t)r * m.Hq + head) * m.nsplit) * kPart;
    float M = -INFINITY;
    for (int sp = 0; sp < nact; ++sp) M = fmaxf(M, __ldcg(base + (size_t)sp * kPart + kHD));
    float L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
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
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");       // scratch free for the next unit
}

template <int R>
__global__ void __launch_bounds__((kW3 + 1) * 32, 1) decode_step_kernel(const __grid_constant__ MegaArgs m) {
  extern __shared__ __align__(128) unsigned char smem_m[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ float red[2][kW3][4];
  if (loop_idle(m.loop, m.T_delayed)) return;                 // same answer in every CTA: the loop state only changes in the sampler
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = m.ring_stages;
  unsigned char* ring = smem_m;
  float* part = reinterpret_cast<float*>(smem_m + (size_t)S * kMegaStageBytes);
  bf16* staging = reinterpret_cast<bf16*>(smem_m + (size_t)S * kMegaStageBytes + m.part_bytes);
  unsigned char* attn_scratch = reinterpret_cast<unsigned char*>(staging) + (size_t)kW3 * R * kMegaWarpK * sizeof(bf16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kW3); }
    mbar_fence_init();
  }
  __syncthreads();
  const int nph = mega_num_phases(m);
  MegaPhase p;

  if (warp == kW3) {
    // ===== producer: the whole step's weights, in consumption order =====
    int gst = 0;
#pragma unroll 1
    for (int ph = 1; ph < nph; ++ph) {
      mega_phase(m, ph, nph, p);
      if (p.streams) mega_produce(m, p, ring, full_bar, empty_bar, S, gst, lane);
    }
    return;
  }

  // ===== consumers =====
  const unsigned epoch = m.sync[1];                         // written by this session's previous live step
  int gst = 0, gst_out = 0;
  const MegaAttnMeta ameta = mega_attention_meta(m, blockIdx.x, R * m.Hkv * m.nsplit);
  const bool stamping = m.timeline && blockIdx.x == 0;
  if (stamping && threadIdx.x == 0) m.timeline[0] = gtime();
  // phase 0: codebook embedding sum (sequential bf16 adds, codec_utils.py:37) for this CTA's columns, both CFG rows
  {
    const uint32_t tag0 = mega_tag(epoch, nph, 0);
    const int d_begin = (int)((long long)blockIdx.x * m.D / gridDim.x), d_end = (int)((long long)(blockIdx.x + 1) * m.D / gridDim.x);
    const long long col = m.loop ? (long long)m.loop->offset : 0;
    for (int t = threadIdx.x; t < (d_end - d_begin) * m.B; t += kW3 * 32) {
      const int b = t / (d_end - d_begin), dd = d_begin + t % (d_end - d_begin);
      float acc = 0.f;
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

#pragma unroll 1
  for (int ph = 1; ph < nph; ++ph) {
    mega_phase(m, ph, nph, p);
    const uint32_t tag_in = mega_tag(epoch, nph, ph - 1), tag_out = mega_tag(epoch, nph, ph);
    unsigned long long* slot = (stamping && 2 * ph + 1 < 512) ? &m.timeline[2 * ph] : nullptr;
    if (p.kind == MP_ATTN) {
      for (int unit = blockIdx.x; unit < R * m.Hkv * m.nsplit; unit += gridDim.x)
        mega_attention_unit(m, p.kv_layer, unit, attn_scratch, warp, lane, unit == (int)blockIdx.x, tag_in, tag_out,
                            unit == (int)blockIdx.x ? slot : nullptr);
    } else {
      // K/V of earlier tokens for this layer's attention: in flight while in_proj streams
      if (p.kind == MP_INPROJ) mega_attention_prefetch(m, p.kv_layer, ameta, attn_scratch);
      if (p.first_out) gst_out = gst;
      mega_consume<R>(m, p, ring, part, staging, full_bar, empty_bar, red, S, gst, warp, lane, tag_in, tag_out, slot);
      if (!p.release) gst = gst_out;                          // the held out_proj slice is consumed again by the next pass
    }
    if (slot && threadIdx.x == 0) slot[1] = gtime();
  }
  // CTA 0 can only get here after it consumed outputs of every CTA, i.e. after every CTA read the epoch
  if (blockIdx.x == 0 && threadIdx.x == 0) m.sync[1] = epoch + 1;
}
