p='zonos_b200/csrc/decode.cu'; s=open(p).read()

# ---------- constants + helpers
old="constexpr int kMegaStageBytes = 32 * 1024, kMegaAttnBytes = 40 * 1024;"
new='''// A ring stage holds kMegaRows weight rows (one n8 MMA tile) x one k-block of KB = min(K, kMegaKB) elements.  Every row
// is its own bulk copy and the rows sit (row bytes + 16) apart, so the eight 16-byte row segments one ldmatrix phase
// reads fall into eight different bank groups.
constexpr int kMegaRows = 8, kMegaKB = 2048, kMegaWarpK = 128;
constexpr int kMegaStageBytes = kMegaRows * (kMegaKB * 2 + 16), kMegaAttnBytes = 40 * 1024;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// D[16 x 8] += A[16 x 16] B[16 x 8] (bf16 in, fp32 accumulate): A = activation rows, B = 8 weight rows
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 ld_relaxed_v2(const uint32_t* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}'''
assert s.count(old)==1; s=s.replace(old,new)

# ---------- producer
a_=s.index("// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst)")
b_=s.index("// consumers: one matrix phase.  wait_full / release control the out_proj")
new='''// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst): one stage per (group
// of kMegaRows rows, k-block), the k-blocks of a group back to back
template <int EPI>
__device__ __forceinline__ void mega_produce(const GemvArgs& a, unsigned char* ring, uint64_t* full_bar, uint64_t* empty_bar, int S, int& gst,
                                             uint64_t pol, int lane) {
  const int K = a.K, KB = min(K, kMegaKB), NB = K / KB, pitch = KB * 2 + 16;
  int u_begin, nrows;
  mega_slice<EPI>(a, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  int slot = gst % S, parity = ((gst / S) - 1) & 1;
  for (int gi = 0; gi < ngroup; ++gi) {
    const bf16* src = a.W + (size_t)row_of_local<EPI>(a, u_begin, min(gi * kMegaRows + (lane & 7), nrows - 1)) * K;
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

'''
s=s[:a_]+new+s[b_:]

# ---------- consumer prologue
a_=s.index("// consumers: one matrix phase.  wait_full / release control the out_proj")
b_=s.index("  // ---- operands of this thread's epilogue (residual value, RoPE cos/sin, KV page): fetched NOW so their L2 round", a_)
new=r'''// consumers: one matrix phase on the tensor cores (mma.sync m16n8k16, fp32 accumulate): the stage loop is a few dozen
// instructions per warp instead of ~180 with FFMA - the 16 warps share 4 issue slots, so every 100 instructions per
// thread cost the phase 0.2 us.  Warp w owns the k-slice [w*128, w*128+128) of every k-block (KS = KB/128 warps take
// part).  The A operand holds "virtual rows" rho = kb*R + i (k-block kb of activation row i), so the fragments of all
// k-blocks live in the same registers, spread over the lane groups; B fragments (8 weight rows x 16 k) come from the
// ring with ldmatrix; the lane-level reduction is the MMA itself, the KS*NB partials of an output meet in `part`.
// release = false keeps the slots (out_proj "hold": the first pass leaves the slice in the ring for the second).
template <int R, int NB, int PRO, int EPI>
__device__ __forceinline__ void mega_consume(const GemvArgs& a, unsigned char* ring, float* part, uint64_t* full_bar, uint64_t* empty_bar,
                                             float (*red)[kW3][4], int S, int& gst, bool release, int warp, int lane,
                                             const uint32_t* xt, uint32_t tag_in, uint32_t* yt, uint32_t tag_out, const uint32_t* rt, uint32_t* qt,
                                             uint32_t* kvt, unsigned long long* stamp, int norm_pending = 0, const MegaQkvPre* qkv_pre = nullptr, unsigned long long* dbg = nullptr) {
#define DBG(i) do { if (dbg && threadIdx.x == 0) dbg[i] = gtime(); } while (0)
  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);
  constexpr int NSTEP = kMegaWarpK / 16;                        // 16-k MMA steps per warp and k-block
  constexpr int VR = R * NB;                                    // virtual A rows
  constexpr bool kHi = VR > 8;                                  // rows 8..15 of the A tile in use
  static_assert(VR <= 16, "at most 16 virtual rows");
  static_assert(NB == 1 || PRO == PRO_NONE, "the norm prologue covers one k-block");
  const int K = a.K, KB = K / NB, pitch = KB * 2 + 16;
  const int KS = KB / kMegaWarpK, KST = KS * NB;
  const bool active = warp < KS;
  int u_begin, nrows;
  mega_slice<EPI>(a, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  const int g = lane >> 2, c = lane & 3;
  const bool lo_on = active && g < VR, hi_on = active && kHi && g + 8 < VR;
  // element offset of this lane's first pair inside its virtual rows' k-slice (pairs at +16s and +16s+8)
  const uint32_t* x_lo = xt + (size_t)(g % R) * a.ldx + (size_t)(g / R) * KB + warp * kMegaWarpK + 2 * c;
  const uint32_t* x_hi = xt + (size_t)((g + 8) % R) * a.ldx + (size_t)((g + 8) / R) * KB + warp * kMegaWarpK + 2 * c;

  // activations, already in MMA fragment layout: spin on the operand loads themselves until every word carries the
  // producing phase's tag
  uint32_t afr[NSTEP][kHi ? 4 : 2];
  float xf[(PRO == PRO_NORM) ? NSTEP * 4 : 1];                  // PRO_NORM: fp32 copy for the norm (one virtual row per lane)
  for (unsigned spins = 0;; ++spins) {
    bool ok = true;
#pragma unroll
    for (int st = 0; st < NSTEP; ++st) {
      if (lo_on) {
        const uint2 p0 = ld_relaxed_v2(x_lo + st * 16), p1 = ld_relaxed_v2(x_lo + st * 16 + 8);
        ok = ok && ((p0.x & 0xffffu) == tag_in) && ((p0.y & 0xffffu) == tag_in) && ((p1.x & 0xffffu) == tag_in) && ((p1.y & 0xffffu) == tag_in);
        afr[st][0] = (p0.x >> 16) | (p0.y & 0xffff0000u);
        afr[st][kHi ? 2 : 1] = (p1.x >> 16) | (p1.y & 0xffff0000u);
        if (PRO == PRO_NORM) { xf[st * 4 + 0] = untag(p0.x); xf[st * 4 + 1] = untag(p0.y); xf[st * 4 + 2] = untag(p1.x); xf[st * 4 + 3] = untag(p1.y); }
      } else {
        afr[st][0] = 0u; afr[st][kHi ? 2 : 1] = 0u;
        if (PRO == PRO_NORM) { xf[st * 4 + 0] = 0.f; xf[st * 4 + 1] = 0.f; xf[st * 4 + 2] = 0.f; xf[st * 4 + 3] = 0.f; }
      }
      if (kHi) {
        if (hi_on) {
          const uint2 p0 = ld_relaxed_v2(x_hi + st * 16), p1 = ld_relaxed_v2(x_hi + st * 16 + 8);
          ok = ok && ((p0.x & 0xffffu) == tag_in) && ((p0.y & 0xffffu) == tag_in) && ((p1.x & 0xffffu) == tag_in) && ((p1.y & 0xffffu) == tag_in);
          afr[st][1] = (p0.x >> 16) | (p0.y & 0xffff0000u);
          afr[st][3] = (p1.x >> 16) | (p1.y & 0xffff0000u);
        } else { afr[st][1] = 0u; afr[st][3] = 0u; }
      }
    }
    if (ok) break;
    if (spins > kMegaSpinLimit) asm volatile("trap;");
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();
  DBG(0);
  if (PRO == PRO_NORM) {
    // row statistics: this lane holds 32 elements of row g; the 4 lanes of a row, then the KS warps
    float sacc = 0.f, qacc = 0.f;
#pragma unroll
    for (int e = 0; e < NSTEP * 4; ++e) { sacc += xf[e]; qacc = fmaf(xf[e], xf[e], qacc); }
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1); qacc += __shfl_xor_sync(0xffffffffu, qacc, 1);
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 2); qacc += __shfl_xor_sync(0xffffffffu, qacc, 2);
    if (lo_on && c == 0) { red[0][warp][g] = sacc; red[1][warp][g] = qacc; }
    DBG(1);
    // the norm parameters sit in shared memory (copied there a layer ahead: a global load issued here would queue
    // behind the saturated weight stream for microseconds); this thread's copies are complete after the wait, all
    // threads' after the barrier
    if (norm_pending == 0) asm volatile("cp.async.wait_group 0;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
    DBG(2);
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");
    DBG(3);
    if (lo_on) {
      const float inv_k = 1.0f / (float)K;                      // K is a power of two: multiplying is exact
      float tot = 0.f, tsq = 0.f;
      for (int q = 0; q < KS; ++q) { tot += red[0][q][g]; tsq += red[1][q][g]; }
      const float mu = tot * inv_k;
      const float mean = (a.norm_kind == ZB_NORM_LAYERNORM) ? mu : 0.f;
      const float var = (a.norm_kind == ZB_NORM_LAYERNORM) ? fmaxf(tsq * inv_k - mu * mu, 0.f) : tsq * inv_k;
      const float rstd = rsqrtf(var + a.eps);
      const bf16* nwp = a.nw + warp * kMegaWarpK + 2 * c;
      const bf16* nbp = a.nb ? a.nb + warp * kMegaWarpK + 2 * c : nullptr;
#pragma unroll
      for (int st = 0; st < NSTEP; ++st) {
        const uint32_t g0 = *reinterpret_cast<const uint32_t*>(nwp + st * 16), g1 = *reinterpret_cast<const uint32_t*>(nwp + st * 16 + 8);
        const uint32_t b0 = nbp ? *reinterpret_cast<const uint32_t*>(nbp + st * 16) : 0u, b1 = nbp ? *reinterpret_cast<const uint32_t*>(nbp + st * 16 + 8) : 0u;
        const float y0 = (xf[st * 4 + 0] - mean) * rstd * bf16lo(g0) + bf16lo(b0), y1 = (xf[st * 4 + 1] - mean) * rstd * bf16hi(g0) + bf16hi(b0);
        const float y2 = (xf[st * 4 + 2] - mean) * rstd * bf16lo(g1) + bf16lo(b1), y3 = (xf[st * 4 + 3] - mean) * rstd * bf16hi(g1) + bf16hi(b1);
        afr[st][0] = pack_bf16(y0, y1);
        afr[st][kHi ? 2 : 1] = pack_bf16(y2, y3);
      }
    }
  } else {
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // everyone has left the previous phase's epilogue: `part` is free
  }
  DBG(4);

'''
s=s[:a_]+new+s[b_:]

# ---------- stage loop
a_=s.index("  constexpr int V = RW * R;\n  const int my_idx = multi_reduce_index<V>(lane);", s.index("template <int R, int NB, int PRO, int EPI>"))
b_=s.index("  DBG(6);\n  asm volatile(\"bar.sync 1, %0;\" ::\"n\"(kW3 * 32) : \"memory\");\n  DBG(7);", a_)
new=r'''  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(lane & 7) * pitch + (uint32_t)(warp * kMegaWarpK + (lane >> 3) * 8) * 2;
  uint32_t slot = (uint32_t)(gst % S), phase = (uint32_t)((gst / S) & 1);   // advanced incrementally: no division per stage
  uint32_t src = lane_base + slot * kMegaStageBytes, fb = full0 + slot * 8, eb = empty0 + slot * 8;
  // D[virtual row][n = 2c, 2c+1]: partial of weight rows 2c, 2c+1 of the group for k-slice (kb, warp)
  float* dst = part + ((size_t)(2 * c) * KST + warp) * R;
  const int dst_row = KST * R, dst_group = kMegaRows * KST * R;
  gst += ngroup * NB;
  for (int gi = 0; gi < ngroup; ++gi) {
#pragma unroll
    for (int kb = 0; kb < NB; ++kb) {
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      mbar_wait_u32(fb, phase);
      if (gi == 0 && kb == 0) DBG(5);
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
      if (++slot == (uint32_t)S) { slot = 0; phase ^= 1u; src = lane_base; fb = full0; eb = empty0; }
      else { src += kMegaStageBytes; fb += 8; eb += 8; }
      // the rows of this k-block: virtual rows kb*R .. kb*R + R-1
      if (active) {
        const int lo_i = g - kb * R, hi_i = g + 8 - kb * R;
        if (lo_i >= 0 && lo_i < R) { dst[kb * KS * R + lo_i] = d0[0] + d1[0]; dst[kb * KS * R + lo_i + dst_row] = d0[1] + d1[1]; }
        if (kHi && hi_i >= 0 && hi_i < R) { dst[kb * KS * R + hi_i] = d0[2] + d1[2]; dst[kb * KS * R + hi_i + dst_row] = d0[3] + d1[3]; }
      }
    }
    dst += dst_group;
  }
'''
s=s[:a_]+new+s[b_:]

# ---------- epilogue partial sums use KST
e_=s.index("  if (e_on) {\n    float v0 = 0.f, v1 = 0.f, u0 = 0.f;", s.index("template <int R, int NB, int PRO, int EPI>"))
f_=s.index("    if (EPI == EPI_RESID) {\n      st_relaxed_u32(yt", e_)
new='''  if (e_on) {
    float v0 = 0.f, v1 = 0.f, u0 = 0.f;
    if (kPairs) {
      const float* s0 = part + (size_t)(2 * ej) * KST * R;
      const float* s1 = s0 + (size_t)KST * R;
      for (int q = 0; q < KST; ++q) { v0 += s0[q * R + ei]; v1 += s1[q * R + ei]; }
    } else {
      const float* s0 = part + (size_t)ej * KST * R;
      for (int q = 0; q < KST; ++q) { v0 += s0[q * R + ei]; if (cfg) u0 += s0[q * R + a.B + ei]; }
    }
'''
s=s[:e_]+new+s[f_:]

# ---------- call sites
def rep(old,new,cnt=1):
    global s
    assert s.count(old)==cnt,(s.count(old),old[:80]); s=s.replace(old,new)
rep("mega_consume<R, 1, 4, PRO_NORM, EPI_QKV>", "mega_consume<R, 1, PRO_NORM, EPI_QKV>")
rep("mega_consume<R, 1, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, g2,", "mega_consume<R, 1, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, g2,")
rep("mega_consume<R, 1, 4, PRO_NONE, EPI_STORE>", "mega_consume<R, 1, PRO_NONE, EPI_STORE>")
rep("mega_consume<R, 1, 4, PRO_NORM, EPI_SILU>", "mega_consume<R, 1, PRO_NORM, EPI_SILU>")
rep("mega_consume<R, 1, 4, PRO_NORM, EPI_HEADS>", "mega_consume<R, 1, PRO_NORM, EPI_HEADS>")
rep("""      if (m.F == 8192)
        mega_consume<R, 2, 2, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);
      else
        mega_consume<R, 1, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);""",
"""      if (m.F == 4 * kMegaKB)
        mega_consume<R, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                m.xt, nullptr, nullptr, slot);
      else if (m.F == 2 * kMegaKB)
        mega_consume<R, 2, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                m.xt, nullptr, nullptr, slot);
      else
        mega_consume<R, 1, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                m.xt, nullptr, nullptr, slot);""")

# ---------- host: supported shapes, partial buffer
rep('''  auto k_ok = [](int K) { return K == 256 || K == 512 || K == 1024 || K == 2048 || K == 4096; };''',
    '''  auto k_ok = [](int K) { return K == 256 || K == 512 || K == 1024 || K == 2048; };   // one k-block, 128 k per warp''')
rep("k_ok(d.d_model) && k_ok(qn) && (k_ok(d.d_ff) || d.d_ff == 8192) &&", "k_ok(d.d_model) && k_ok(qn) && (k_ok(d.d_ff) || d.d_ff == 4096 || d.d_ff == 8192) &&")
a_=s.index("  // partial-sum buffer: the largest padded row count x k-slices over all matrices of the step")
b_=s.index("  pb = (pb + 1023) / 1024 * 1024;")
new='''  // partial-sum buffer: the largest padded row count x k-slices over all matrices of the step
  auto part_need = [&](int nunits, bool pairs, int K) {
    const int KST = K / kMegaWarpK;                             // KS warps x NB k-blocks
    const int rows = ((nunits + grid - 1) / grid) * (pairs ? 2 : 1);
    return (size_t)((rows + kMegaRows - 1) / kMegaRows * kMegaRows) * KST * R * sizeof(float);
  };
  const int qn = d.n_heads * d.head_dim;
  size_t pb = part_need((d.n_heads + 2 * d.n_heads_kv) * d.head_dim / 2, true, d.d_model);
  pb = std::max(pb, part_need(d.d_model, false, qn));
  pb = std::max(pb, part_need(d.d_ff, true, d.d_model));
  pb = std::max(pb, part_need(d.d_model, false, d.d_ff));
  pb = std::max(pb, part_need(m.QV, false, d.d_model));
'''
s=s[:a_]+new+s[b_:]
open(p,'w').write(s)
print("ok")
