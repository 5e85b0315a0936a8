p='zonos_b200/csrc/decode.cu'; s=open(p).read()

def rep(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:80])
    s = s.replace(old, new)

# ---- 1. MegaArgs + sync helpers
a_=s.index("struct MegaArgs {"); b_=s.index("template <int EPI>\n__device__ __forceinline__ void mega_slice")
new='''struct MegaArgs {
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
  int ring_stages, part_bytes;
  unsigned long long* timeline;   // debug: globaltimer stamps of CTA 0 (2 per phase: inputs ready, work done)
};

constexpr int kMegaStageBytes = 32 * 1024;

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

'''
s=s[:a_]+new+s[b_:]

# ---- 2. mega_consume: signature, activation load, epilogue
rep('''                                             float (*red)[kW3][4], int S, int& gst, bool release, int warp, int lane) {''',
'''                                             float (*red)[kW3][4], int S, int& gst, bool release, int warp, int lane,
                                             const uint32_t* xt, uint32_t tag_in, uint32_t* yt, uint32_t tag_out, const uint32_t* rt, uint32_t* qt,
                                             uint32_t* kvt, unsigned long long* stamp) {''')
a_=s.index("  float xf[R][NC * 8];\n#pragma unroll\n  for (int i = 0; i < R; ++i)\n#pragma unroll\n    for (int c = 0; c < NC; ++c) {\n      const uint4 v = (i < a.M) ? __ldcg(reinterpret_cast<const uint4*>(a.x + (size_t)i * a.ldx + koff + c * 256 + lane * 8))", s.index("mega_consume("))
b_=s.index("  if (PRO == PRO_NORM) {\n    float mean[R], rstd[R];", a_)
new='''  // activations: spin on the operand loads themselves until every word carries the producing phase's tag
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
  if (stamp && threadIdx.x == 0) *stamp = gtime();
'''
s=s[:a_]+new+s[b_:]
rep("    if (EPI == EPI_RESID) pre_resid = ldcg_bf16(a.resid + (size_t)ei * a.ldr + en0);\n    if (EPI == EPI_QKV) {\n      pre_pos",
    "    if (EPI == EPI_RESID) pre_resid = untag(ld_relaxed_u32(rt + (size_t)ei * a.ldr + en0));   // validated by this CTA in an earlier phase\n    if (EPI == EPI_QKV) {\n      pre_pos")
rep('''    if (EPI == EPI_RESID) {
      a.y[(size_t)ei * a.ldy + en0] = f2bf(pre_resid + rbf(v0));
    } else if (EPI == EPI_QKV) {''',
'''    if (EPI == EPI_RESID) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(pre_resid + rbf(v0), tag_out));
    } else if (EPI == EPI_STORE) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(v0, tag_out));
    } else if (EPI == EPI_SILU) {                            // same ops as gemv_epilogue<EPI_SILU>
      const float yv = rbf(v0), g = rbf(v1);
      const float sg = rbf(g / (1.0f + expf(-g)));
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(__fmul_rn(yv, sg), tag_out));
    } else if (EPI == EPI_QKV) {''')
rep('''      if (en0 < qn_) {
        a.q_out[(size_t)ei * qn_ + en0] = f2bf(o0);
        a.q_out[(size_t)ei * qn_ + en1] = f2bf(o1);
      } else {
        const int kvsel = en0 < qn_ + kn_ ? 0 : 1;
        const int c0i = en0 - qn_ - kvsel * kn_, c1i = en1 - qn_ - kvsel * kn_;
        bf16* pb = a.kv_layer + ((size_t)pre_page * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;''',
'''      if (en0 < qn_) {
        st_relaxed_u32(qt + (size_t)ei * qn_ + en0, tag_word(o0, tag_out));
        st_relaxed_u32(qt + (size_t)ei * qn_ + en1, tag_word(o1, tag_out));
      } else {
        const int kvsel = en0 < qn_ + kn_ ? 0 : 1;
        const int c0i = en0 - qn_ - kvsel * kn_, c1i = en1 - qn_ - kvsel * kn_;
        // this step's attention reads the new token from the tagged side buffer; the cache copy is for later steps
        st_relaxed_u32(kvt + ((size_t)ei * 2 + kvsel) * kn_ + c0i, tag_word(o0, tag_out));
        st_relaxed_u32(kvt + ((size_t)ei * 2 + kvsel) * kn_ + c1i, tag_word(o1, tag_out));
        bf16* pb = a.kv_layer + ((size_t)pre_page * 2 + kvsel) * a.Hkv * ZB_PAGE_TOKENS * a.hd;''')

# ---- 3. attention unit
rep('''__device__ __forceinline__ void mega_attention_unit(const MegaArgs& m, const bf16* kv_layer, int unit, unsigned char* scratch, int warp, int lane,
                                                    bool prefetched) {''',
'''__device__ __forceinline__ void mega_attention_unit(const MegaArgs& m, const bf16* kv_layer, int unit, unsigned char* scratch, int warp, int lane,
                                                    bool prefetched, uint32_t tag_in, uint32_t tag_out, unsigned long long* stamp) {''')
rep('''  const int n_old = prefetched ? max(0, min(nk, kv_len - 1 - k0)) : 0;
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += kW3 * 32) {
    const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
    if (tok >= n_old && tok < nk) {''',
'''  const int n_old = prefetched ? max(0, min(nk, kv_len - 1 - k0)) : 0;
  const int tok_new = kv_len - 1 - k0;                                  // this step's token, if it falls into this split
  for (int c = threadIdx.x; c < kCH * kHD / 8; c += kW3 * 32) {
    const int tok = c / (kHD / 8), d8 = (c % (kHD / 8)) * 8;
    if (tok >= n_old && tok < nk && tok != tok_new) {''')
rep('''  asm volatile("cp.async.commit_group;\\ncp.async.wait_group 0;" ::: "memory");
  const int head = g * G + warp;
  if (warp < G) {
    const uint2 qv = __ldcg(reinterpret_cast<const uint2*>(m.q + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4));
    qs[warp * kHD + lane * 4 + 0] = bf16lo(qv.x); qs[warp * kHD + lane * 4 + 1] = bf16hi(qv.x);
    qs[warp * kHD + lane * 4 + 2] = bf16lo(qv.y); qs[warp * kHD + lane * 4 + 3] = bf16hi(qv.y);
  }
''',
'''  asm volatile("cp.async.commit_group;" ::: "memory");
  const int head = g * G + warp;
  if (warp < G) {                                                       // q of this step: tagged words from the in_proj phase
    const uint4 qv = poll_v4(m.qt + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4, tag_in);
    *reinterpret_cast<float4*>(&qs[warp * kHD + lane * 4]) = make_float4(untag(qv.x), untag(qv.y), untag(qv.z), untag(qv.w));
  } else if (warp >= 8 && warp < 10 && tok_new >= 0 && tok_new < nk) {  // K (warp 8) and V (warp 9) of this step's token
    const int kvsel = warp - 8;
    const uint4 nv = poll_v4(m.kvt + ((size_t)r * 2 + kvsel) * m.Hkv * kHD + (size_t)g * kHD + lane * 4, tag_in);
    uint2 pk;
    pk.x = (nv.x >> 16) | (nv.y & 0xffff0000u);
    pk.y = (nv.z >> 16) | (nv.w & 0xffff0000u);
    bf16* dst = kvsel ? vs + tok_new * kHD : ks + tok_new * kKStride;
    *reinterpret_cast<uint2*>(dst + lane * 4) = pk;
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();
  asm volatile("cp.async.wait_group 0;" ::: "memory");
''')
rep('''    const float inv = 1.0f / L;
    uint2 outv;
    outv.x = pack_bf16(acc[0] * inv, acc[1] * inv);
    outv.y = pack_bf16(acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<uint2*>(m.attn_y + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4) = outv;
  }''',
'''    const float inv = 1.0f / L;
    const uint4 outv = make_uint4(tag_word(acc[0] * inv, tag_out), tag_word(acc[1] * inv, tag_out), tag_word(acc[2] * inv, tag_out),
                                  tag_word(acc[3] * inv, tag_out));
    st_relaxed_v4(m.ayt + (size_t)r * m.Hq * kHD + (size_t)head * kHD + lane * 4, outv);
  }''')

# ---- 4. kernel body (consumers)
k_=s.index("decode_step_kernel(const __grid_constant__ MegaArgs m)")
a_=s.index("  // ===== consumers =====", k_)
b_=s.index("// ------------------------------------------------------------------ plain norm ---------------", a_)
new='''  // ===== consumers =====
  const unsigned epoch = m.sync[1];                         // written by this session's previous live step
  const int nph = 2 + m.n_layer * (4 + m.out_proj_repeats);
  int ph = 0, gst = 0, stamp_i = 0;
  const MegaAttnMeta ameta = mega_attention_meta(m, blockIdx.x, R * m.Hkv * m.nsplit);
  const bool stamping = m.timeline && blockIdx.x == 0;
#define MEGA_STAMP() do { if (stamping && threadIdx.x == 0 && stamp_i < 126) m.timeline[stamp_i] = gtime(); ++stamp_i; } while (0)
#define MEGA_STAMP_SLOT() ((stamping && stamp_i < 126) ? &m.timeline[stamp_i++] : (++stamp_i, (unsigned long long*)nullptr))
#define TAG(p) mega_tag(epoch, nph, (p))
  MEGA_STAMP();
  // phase 0: codebook embedding sum (sequential bf16 adds, codec_utils.py:37) for this CTA's columns, both CFG rows
  {
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
      st_relaxed_u32(m.xt + (size_t)b * m.D + dd, tag_word(acc, TAG(0)));
      st_relaxed_u32(m.xt + (size_t)(m.B + b) * m.D + dd, tag_word(acc, TAG(0)));
    }
  }
  MEGA_STAMP();
  ph = 1;

  for (int li = 0; li < m.n_layer; ++li) {
    const MegaLayer& L = m.layers[li];
    // A: norm -> in_proj -> RoPE -> KV append (+ q)
    mega_fill(a, m, R);
    a.W = L.in_proj; a.N = nqkv; a.K = m.D; a.ldx = m.D; a.nw = L.norm_w; a.nb = L.norm_b; a.kv_layer = L.kv_layer;
    mega_attention_prefetch(m, L.kv_layer, ameta, attn_scratch);
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      mega_consume<R, 1, 4, PRO_NORM, EPI_QKV>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), nullptr, TAG(ph),
                                               nullptr, m.qt, m.kvt, slot);
    }
    MEGA_STAMP(); ++ph;
    // B: attention over the paged cache
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      for (int unit = blockIdx.x; unit < R * m.Hkv * m.nsplit; unit += gridDim.x)
        mega_attention_unit(m, L.kv_layer, unit, attn_scratch, warp, lane, unit == (int)blockIdx.x, TAG(ph - 1), TAG(ph),
                            unit == (int)blockIdx.x ? slot : nullptr);
    }
    MEGA_STAMP(); ++ph;
    // C/D: out_proj (twice in the reference); the slice stays in the ring between the passes
    {
      const int gst0 = gst;
      const uint32_t* src = m.ayt;
      for (int rep = 0; rep < m.out_proj_repeats; ++rep) {
        const bool last = rep == m.out_proj_repeats - 1;
        mega_fill(a, m, R);
        a.W = L.out_proj; a.N = m.D; a.K = qn; a.ldx = qn; a.ldy = m.D; a.ldr = m.D;
        int g2 = gst0;
        unsigned long long* slot = MEGA_STAMP_SLOT();
        if (last) {
          mega_consume<R, 1, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, g2, true, warp, lane, src, TAG(ph - 1), m.xt, TAG(ph),
                                                     m.xt, nullptr, nullptr, slot);
        } else {
          uint32_t* dst = (src == m.y1t) ? m.ayt : m.y1t;
          mega_consume<R, 1, 4, PRO_NONE, EPI_STORE>(a, ring, part, full_bar, empty_bar, red, S, g2, false, warp, lane, src, TAG(ph - 1), dst, TAG(ph),
                                                     nullptr, nullptr, nullptr, slot);
          src = dst;
        }
        gst = g2;
        MEGA_STAMP(); ++ph;
      }
    }
    // E: norm2 -> fc1 -> value * silu(gate)
    mega_fill(a, m, R);
    a.W = L.fc1; a.N = 2 * m.F; a.K = m.D; a.ldx = m.D; a.nw = L.norm2_w; a.nb = L.norm2_b; a.ldy = m.F;
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      mega_consume<R, 1, 4, PRO_NORM, EPI_SILU>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), m.ht, TAG(ph),
                                                nullptr, nullptr, nullptr, slot);
    }
    MEGA_STAMP(); ++ph;
    // F: fc2 + residual
    mega_fill(a, m, R);
    a.W = L.fc2; a.N = m.D; a.K = m.F; a.ldx = m.F; a.ldy = m.D; a.ldr = m.D;
    {
      unsigned long long* slot = MEGA_STAMP_SLOT();
      if (m.F == 8192)
        mega_consume<R, 2, 2, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);
      else
        mega_consume<R, 1, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);
    }
    MEGA_STAMP(); ++ph;
  }
  // heads: final norm -> fused heads -> fp32 -> CFG mix
  mega_fill(a, m, R);
  a.W = m.heads; a.N = m.QV; a.K = m.D; a.ldx = m.D; a.nw = m.normf_w; a.nb = m.normf_b;
  {
    unsigned long long* slot = MEGA_STAMP_SLOT();
    mega_consume<R, 1, 4, PRO_NORM, EPI_HEADS>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.xt, TAG(ph - 1), nullptr, 0u,
                                               nullptr, nullptr, nullptr, slot);
  }
  MEGA_STAMP();
  // CTA 0 can only get here after it consumed outputs of every CTA, i.e. after every CTA read the epoch
  if (blockIdx.x == 0 && threadIdx.x == 0) m.sync[1] = epoch + 1;
#undef MEGA_STAMP
#undef MEGA_STAMP_SLOT
#undef TAG
}

'''
s=s[:a_]+new+s[b_:]

# ---- 5. host launcher
rep("                                bf16* x, int R, int max_kv_len,", "                                uint32_t* arena, int R, int max_kv_len,")
rep("zb_status zb_launch_decode_step(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* mega_layers_dev, unsigned* bar,",
    "zb_status zb_launch_decode_step(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* mega_layers_dev, unsigned* sync,")
rep("  m.x = x; m.q = s.q; m.attn_y = s.attn_y; m.y1 = s.y1; m.h = s.h; m.attn_part = s.part; m.attn_counters = ctx->counters; m.nsplit = nsplit;\n  m.scale = 1.0f / sqrtf((float)d.head_dim); m.loop = loop; m.bar = bar; m.timeline = g_timeline;",
'''  {  // tagged activation words: the session's own zero-initialised arena (zb_mega_arena_bytes)
    const size_t qn_ = (size_t)d.n_heads * d.head_dim, kn_ = (size_t)d.n_heads_kv * d.head_dim;
    uint32_t* p = arena;
    m.xt = p; p += (size_t)R * d.d_model;
    m.qt = p; p += (size_t)R * qn_;
    m.ayt = p; p += (size_t)R * qn_;
    m.y1t = p; p += (size_t)R * d.d_model;
    m.ht = p; p += (size_t)R * d.d_ff;
    m.kvt = p; p += (size_t)R * 2 * kn_;
  }
  m.attn_part = s.part; m.attn_counters = ctx->counters; m.nsplit = nsplit;
  m.scale = 1.0f / sqrtf((float)d.head_dim); m.loop = loop; m.sync = sync; m.timeline = g_timeline;''')
old="size_t zb_mega_layers_bytes(const zb_model* model) { return (size_t)model->d.n_layer * sizeof(MegaLayer); }"
rep(old, old+'''
size_t zb_mega_arena_bytes(const zb_model* model, int R) {
  const zb_model_desc& d = model->d;
  const size_t qn = (size_t)d.n_heads * d.head_dim, kn = (size_t)d.n_heads_kv * d.head_dim;
  return ((size_t)R * (2 * d.d_model + 2 * qn + d.d_ff + 2 * kn)) * sizeof(uint32_t);
}''')
open(p,'w').write(s)
print("ok")
