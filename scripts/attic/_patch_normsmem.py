p='zonos_b200/csrc/decode.cu'; s=open(p).read()
def rep(old,new,cnt=1):
    global s
    assert s.count(old)==cnt,(s.count(old),old[:80]); s=s.replace(old,new)

# --- MegaArgs: evict hint flag
rep("  int ring_stages, part_bytes;\n  unsigned long long* timeline;   // debug: globaltimer stamps of CTA 0 (2 per phase: inputs ready, work done)",
    "  int ring_stages, part_bytes, evict_first;\n  unsigned long long* timeline;   // debug: globaltimer stamps of CTA 0 (2 per phase: inputs ready, work done)")

# --- producer: optional hint
rep('''      if (lane == 0) bulk_g2s(dst, a.W + (size_t)(u_begin + r0) * K, (uint32_t)kMegaStageBytes, &full_bar[slot], pol);''',
    '''      if (lane == 0) {
        if (pol) bulk_g2s(dst, a.W + (size_t)(u_begin + r0) * K, (uint32_t)kMegaStageBytes, &full_bar[slot], pol);
        else bulk_g2s_nohint(dst, a.W + (size_t)(u_begin + r0) * K, (uint32_t)kMegaStageBytes, &full_bar[slot]);
      }''')
rep('''        bulk_g2s(dst + (size_t)q * row_bytes, a.W + (size_t)row_of_local<EPI>(a, u_begin, lr) * K, row_bytes, &full_bar[slot], pol);
      }
    }
  }
}

// consumers: one matrix phase.''','''        const bf16* src = a.W + (size_t)row_of_local<EPI>(a, u_begin, lr) * K;
        if (pol) bulk_g2s(dst + (size_t)q * row_bytes, src, row_bytes, &full_bar[slot], pol);
        else bulk_g2s_nohint(dst + (size_t)q * row_bytes, src, row_bytes, &full_bar[slot]);
      }
    }
  }
}

// consumers: one matrix phase.''')

# --- consume: norm params from shared memory, read after the stats barrier
a_=s.index("  uint4 nwr[NC], nbr[NC];                                      // norm parameters: in flight together with the activations", s.index("__device__ __forceinline__ void mega_consume("))
b_=s.index("  // activations: spin on the operand loads themselves until every word carries the producing phase's tag", a_)
s=s[:a_]+s[b_:]
rep('''    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
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
      const uint32_t gv[4] = {nwr[c].x, nwr[c].y, nwr[c].z, nwr[c].w}, bv[4] = {nbr[c].x, nbr[c].y, nbr[c].z, nbr[c].w};''',
'''    // the norm parameters sit in shared memory (copied there a layer ahead: a global load issued here would queue
    // behind the saturated weight stream for microseconds); this thread's copies are complete after the wait, all
    // threads' after the barrier
    if (norm_pending == 0) asm volatile("cp.async.wait_group 0;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
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
      const size_t kk = koff + c * 256 + lane * 8;
      const uint4 nwv = *reinterpret_cast<const uint4*>(a.nw + kk);
      const uint4 nbv = a.nb ? *reinterpret_cast<const uint4*>(a.nb + kk) : make_uint4(0, 0, 0, 0);
      const uint32_t gv[4] = {nwv.x, nwv.y, nwv.z, nwv.w}, bv[4] = {nbv.x, nbv.y, nbv.z, nbv.w};''')
rep("                                             uint32_t* kvt, unsigned long long* stamp) {\n  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);\n  constexpr int Kc = NC * 256;",
    "                                             uint32_t* kvt, unsigned long long* stamp, int norm_pending = 0) {\n  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);\n  constexpr int Kc = NC * 256;")

# --- attention prefetch always commits a group
rep('''__device__ __forceinline__ void mega_attention_prefetch(const MegaArgs& m, const bf16* kv_layer, const MegaAttnMeta& t, unsigned char* scratch) {
  if (t.n_old <= 0) return;''','''__device__ __forceinline__ void mega_attention_prefetch(const MegaArgs& m, const bf16* kv_layer, const MegaAttnMeta& t, unsigned char* scratch) {
  if (t.n_old <= 0) { asm volatile("cp.async.commit_group;" ::: "memory"); return; }     // always one group: see cp.async.wait_group 1 in the in_proj norm''')

# --- kernel: norm buffers + prefetch
rep('''  unsigned char* attn_scratch = smem_m + (size_t)S * kMegaStageBytes + m.part_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kW3); }''',
'''  unsigned char* attn_scratch = smem_m + (size_t)S * kMegaStageBytes + m.part_bytes;
  bf16* nbuf = reinterpret_cast<bf16*>(attn_scratch + kMegaAttnBytes);   // [2 buffers][weight | bias][D]: norm parameters, copied a layer ahead
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kW3); }''')
rep('''    for (int li = 0; li < m.n_layer; ++li) {
      const MegaLayer& L = m.layers[li];
      mega_fill(a, m, R);
      a.W = L.in_proj; a.N = nqkv; a.K = m.D;
      mega_produce<EPI_QKV>(a, ring, full_bar, empty_bar, S, gst, pol, lane);''',
'''    MegaLayer L = m.layers[0], Lnext = L;
    for (int li = 0; li < m.n_layer; ++li, L = Lnext) {
      if (li + 1 < m.n_layer) Lnext = m.layers[li + 1];       // the next layer's pointers are on their way while this one streams
      mega_fill(a, m, R);
      a.W = L.in_proj; a.N = nqkv; a.K = m.D;
      mega_produce<EPI_QKV>(a, ring, full_bar, empty_bar, S, gst, pol, lane);''')
rep("    const uint64_t pol = l2_evict_first_policy();\n    int gst = 0;\n    MegaLayer L", "    const uint64_t pol = m.evict_first ? l2_evict_first_policy() : 0ull;\n    int gst = 0;\n    MegaLayer L")
# consumer side
rep('''  const MegaAttnMeta ameta = mega_attention_meta(m, blockIdx.x, R * m.Hkv * m.nsplit);
  const bool stamping = m.timeline && blockIdx.x == 0;''','''  const MegaAttnMeta ameta = mega_attention_meta(m, blockIdx.x, R * m.Hkv * m.nsplit);
  const bool stamping = m.timeline && blockIdx.x == 0;
  // norm parameters -> shared memory: one 16-byte cp.async per consumer thread and buffer half
  auto norm_prefetch = [&](int buf, const bf16* w, const bf16* b) {
    const int chunks = m.D / 8;
    bf16* dstw = nbuf + (size_t)buf * 2 * m.D;
    for (int q = threadIdx.x; q < 2 * chunks; q += kW3 * 32) {
      const bf16* src = q < chunks ? w + q * 8 : (b ? b + (q - chunks) * 8 : nullptr);
      if (src) cp_async16(dstw + q * 8, src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  norm_prefetch(0, m.layers[0].norm_w, m.layers[0].norm_b);
  norm_prefetch(1, m.layers[0].norm2_w, m.layers[0].norm2_b);''')
rep('''    a.W = L.in_proj; a.N = nqkv; a.K = m.D; a.ldx = m.D; a.nw = L.norm_w; a.nb = L.norm_b; a.kv_layer = L.kv_layer;''',
    '''    a.W = L.in_proj; a.N = nqkv; a.K = m.D; a.ldx = m.D; a.nw = nbuf; a.nb = L.norm_b ? nbuf + m.D : nullptr; a.kv_layer = L.kv_layer;''')
open(p,'w').write(s)
print('ok')
