p='zonos_b200/csrc/decode.cu'; s=open(p).read()
def rep(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:90])
    s = s.replace(old, new)

rep("constexpr int kMegaStageBytes = 32 * 1024;",
'''// A ring stage holds kMegaRows weight rows (one n8 MMA tile) x one k-block of min(K, kMegaKB) elements; every row is
// its own bulk copy and the rows sit (row bytes + 16) apart, so the eight 16-byte row segments one ldmatrix phase
// reads fall into eight different bank groups.
constexpr int kMegaRows = 8, kMegaKB = 2048, kMegaWarpK = 128;
constexpr int kMegaStageBytes = kMegaRows * (kMegaKB * 2 + 16);

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// D[16 x 8] += A[16 x 16] B[16 x 8]: A = activation rows (only rows 0..R-1 are non-zero: a1 = a3 = 0), B = 8 weight rows
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %5}, {%7, %8}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}''')

a_=s.index("// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst)")
b_=s.index("// K/V of the tokens cached by EARLIER steps for this CTA's first attention unit of the layer")
new=r'''// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst): for every group of
// kMegaRows rows, one stage per k-block
template <int EPI>
__device__ __forceinline__ void mega_produce(const GemvArgs& a, unsigned char* ring, uint64_t* full_bar, uint64_t* empty_bar, int S, int& gst,
                                             uint64_t pol, int lane) {
  const int K = a.K, KB = min(K, kMegaKB), NB = K / KB, pitch = KB * 2 + 16;
  int u_begin, nrows;
  mega_slice<EPI>(a, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  const bf16* src = nullptr;
  for (int gi = 0; gi < ngroup; ++gi) {
    if (lane < kMegaRows) src = a.W + (size_t)row_of_local<EPI>(a, u_begin, min(gi * kMegaRows + lane, nrows - 1)) * K;
    for (int kb = 0; kb < NB; ++kb, ++gst) {
      const int slot = gst % S;
      if (gst >= S) mbar_wait(&empty_bar[slot], ((gst / S) - 1) & 1);
      if (lane == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)(kMegaRows * KB * 2));
      __syncwarp();
      if (lane < kMegaRows)
        bulk_g2s(ring + (size_t)slot * kMegaStageBytes + (size_t)lane * pitch, src + (size_t)kb * KB, (uint32_t)KB * 2, &full_bar[slot], pol);
    }
  }
}

// consumers: one matrix phase on the tensor cores (mma.sync m16n8k16, fp32 accumulate).  Warp w owns the k-slice
// [w*128, w*128+128) of every k-block: its A fragments (the activation rows, bf16) stay in registers for the whole
// phase, B fragments (8 weight rows x 16 k) come from the ring with ldmatrix, and the tile's lane-level reduction is
// the MMA itself; the KS = KB/128 warp partials of each output meet in `part`.  `release` = false keeps the slots
// (out_proj "hold": the first pass leaves its slice in the ring for the second).
template <int R, int NB, int PRO, int EPI>
__device__ __forceinline__ void mega_consume(const GemvArgs& a, unsigned char* ring, float* part, uint64_t* full_bar, uint64_t* empty_bar,
                                             float (*red)[kW3][4], int S, int& gst, bool release, int warp, int lane,
                                             const uint32_t* xt, uint32_t tag_in, uint32_t* yt, uint32_t tag_out, const uint32_t* rt, uint32_t* qt,
                                             uint32_t* kvt, unsigned long long* stamp) {
  constexpr bool kPairs = (EPI == EPI_SILU || EPI == EPI_QKV);
  constexpr int NSTEP = kMegaWarpK / 16;
  const int K = a.K, KB = K / NB, pitch = KB * 2 + 16;
  const int KS = KB / kMegaWarpK;                               // warps that take part (all 16 for KB = 2048)
  const bool active = warp < KS;
  int u_begin, nrows;
  mega_slice<EPI>(a, u_begin, nrows);
  const int ngroup = (nrows + kMegaRows - 1) / kMegaRows;
  const int kw = warp * kMegaWarpK + lane * 4;                  // this lane's 4 elements of the warp's slice, per k-block
  uint2 nwr[NB], nbr[NB];                                       // norm parameters: in flight together with the activations
  if (PRO == PRO_NORM && active) {
#pragma unroll
    for (int kb = 0; kb < NB; ++kb) {
      nwr[kb] = *reinterpret_cast<const uint2*>(a.nw + (size_t)kb * KB + kw);
      nbr[kb] = a.nb ? *reinterpret_cast<const uint2*>(a.nb + (size_t)kb * KB + kw) : make_uint2(0, 0);
    }
  }

  // activations: spin on the operand loads themselves until every word carries the producing phase's tag
  float xf[R][NB][4];
  if (active) {
    for (unsigned spins = 0;; ++spins) {
      bool ok = true;
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int kb = 0; kb < NB; ++kb) {
          const uint4 v = ld_relaxed_v4(xt + (size_t)i * a.ldx + (size_t)kb * KB + kw);
          ok = ok && tags_ok(v, tag_in);
          xf[i][kb][0] = untag(v.x); xf[i][kb][1] = untag(v.y); xf[i][kb][2] = untag(v.z); xf[i][kb][3] = untag(v.w);
        }
      if (ok) break;
      if (spins > kMegaSpinLimit) asm volatile("trap;");
    }
  }
  if (stamp && threadIdx.x == 0) *stamp = gtime();
  if (PRO == PRO_NORM) {
    float mean[R], rstd[R];
    if (active) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        float sacc = 0.f, qacc = 0.f;
#pragma unroll
        for (int kb = 0; kb < NB; ++kb)
#pragma unroll
          for (int e = 0; e < 4; ++e) { sacc += xf[i][kb][e]; qacc = fmaf(xf[i][kb][e], xf[i][kb][e], qacc); }
        sacc = warp_sum(sacc);
        qacc = warp_sum(qacc);
        if (lane == 0) { red[0][warp][i] = sacc; red[1][warp][i] = qacc; }
      }
    }
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
    if (active) {
#pragma unroll
      for (int kb = 0; kb < NB; ++kb) {
        const float g4[4] = {bf16lo(nwr[kb].x), bf16hi(nwr[kb].x), bf16lo(nwr[kb].y), bf16hi(nwr[kb].y)};
        const float b4[4] = {bf16lo(nbr[kb].x), bf16hi(nbr[kb].x), bf16lo(nbr[kb].y), bf16hi(nbr[kb].y)};
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) xf[i][kb][e] = (xf[i][kb][e] - mean[i]) * rstd[i] * g4[e] + b4[e];   // rounded to bf16 below
      }
    }
  } else {
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // `part` (staging below) is free: everyone left the previous epilogue
  }

  // A fragments: through a per-warp staging area (aliases `part`) from "4 consecutive k per lane" to the MMA layout
  // (lane = 4*row + c holds k = 16s + 2c, +1 and 16s + 2c + 8, +9 of every 16-k step s)
  uint32_t afr[NB][NSTEP][2];
  {
    bf16* stg = reinterpret_cast<bf16*>(part) + (size_t)warp * R * kMegaWarpK;
    const int g = lane >> 2, c = lane & 3;
#pragma unroll
    for (int kb = 0; kb < NB; ++kb) {
      if (active) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          uint2 pk;
          pk.x = pack_bf16(xf[i][kb][0], xf[i][kb][1]);
          pk.y = pack_bf16(xf[i][kb][2], xf[i][kb][3]);
          *reinterpret_cast<uint2*>(stg + i * kMegaWarpK + lane * 4) = pk;
        }
      }
      __syncwarp();
#pragma unroll
      for (int st = 0; st < NSTEP; ++st) {
        afr[kb][st][0] = (active && g < R) ? *reinterpret_cast<const uint32_t*>(stg + g * kMegaWarpK + st * 16 + 2 * c) : 0u;
        afr[kb][st][1] = (active && g < R) ? *reinterpret_cast<const uint32_t*>(stg + g * kMegaWarpK + st * 16 + 2 * c + 8) : 0u;
      }
      __syncwarp();
    }
  }

  // ---- operands of this thread's epilogue (residual value, RoPE cos/sin, KV page): fetched NOW so their L2 round
  // trips overlap the weight streaming instead of trailing it ----
  const bool cfg = (EPI == EPI_HEADS && a.cfg_scale != 1.0f);
  const int rows_out = cfg ? a.B : a.M;
  const int nu = kPairs ? nrows / 2 : nrows;
  const int et = threadIdx.x;                                 // one epilogue item per thread (host guarantees nu*rows_out <= 512)
  const bool e_on = et < nu * rows_out;
  const int ej = e_on ? et / rows_out : 0, ei = e_on ? et % rows_out : 0;
  int en0 = 0, en1 = 1;
  if (kPairs) unit_rows<EPI>(a, u_begin + ej, en0, en1); else { en0 = u_begin + ej; en1 = en0 + 1; }
  float pre_resid = 0.f;
  float2 pre_cs = make_float2(1.f, 0.f);
  int pre_pos = 0, pre_page = 0;
  if (e_on) {
    if (EPI == EPI_RESID) pre_resid = untag(ld_relaxed_u32(rt + (size_t)ei * a.ldr + en0));   // validated by this CTA in an earlier phase
    if (EPI == EPI_QKV) {
      pre_pos = a.lengths[ei];
      const int qn_ = a.Hq * a.hd, kn_ = a.Hkv * a.hd;
      if (en0 < qn_ + kn_) {
        const int ri = a.rope_interleaved ? (en0 % a.hd) / 2 : (en0 % a.hd);
        pre_cs = *reinterpret_cast<const float2*>(a.rope + ((size_t)min(pre_pos, a.rope_len - 1) * (a.hd / 2) + ri) * 2);
      }
      if (en0 >= qn_) pre_page = a.page_table[(size_t)ei * a.max_pages + pre_pos / ZB_PAGE_TOKENS];
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // every warp has its fragments: `part` may take partial sums

  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(lane & 7) * pitch + (uint32_t)(warp * kMegaWarpK + (lane >> 3) * 8) * 2;
  const int g = lane >> 2, c = lane & 3;
  for (int gi = 0; gi < ngroup; ++gi) {
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kb = 0; kb < NB; ++kb, ++gst) {
      const uint32_t slot = (uint32_t)(gst % S), phase = (uint32_t)((gst / S) & 1);
      mbar_wait_u32(full0 + slot * 8, phase);
      if (active) {
        const uint32_t src = lane_base + slot * kMegaStageBytes;
#pragma unroll
        for (int j = 0; j < NSTEP / 2; ++j) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(src + j * 64, b0, b1, b2, b3);
          mma_16816(d0, afr[kb][2 * j][0], afr[kb][2 * j][1], b0, b1);
          mma_16816(d1, afr[kb][2 * j + 1][0], afr[kb][2 * j + 1][1], b2, b3);
        }
      }
      __syncwarp();
      if (release && lane == 0) mbar_arrive_u32(empty0 + slot * 8);
    }
    if (active && g < R) {                                     // D[row g][n = 2c, 2c+1]
      float* dst = part + ((size_t)(gi * kMegaRows + 2 * c) * KS + warp) * R + g;
      dst[0] = d0[0] + d1[0];
      dst[(size_t)KS * R] = d0[1] + d1[1];
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");

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
    if (EPI == EPI_RESID) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(pre_resid + rbf(v0), tag_out));
    } else if (EPI == EPI_STORE) {
      st_relaxed_u32(yt + (size_t)ei * a.ldy + en0, tag_word(v0, tag_out));
    } else if (EPI == EPI_SILU) {                            // same ops as gemv_epilogue<EPI_SILU>
      const float yv = rbf(v0), g_ = rbf(v1);
      const float sg = rbf(g_ / (1.0f + expf(-g_)));
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
}

'''
s=s[:a_]+new+s[b_:]

# call sites
rep("mega_consume<R, 1, 4, PRO_NORM, EPI_QKV>", "mega_consume<R, 1, PRO_NORM, EPI_QKV>")
rep("mega_consume<R, 1, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, g2,", "mega_consume<R, 1, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, g2,")
rep("mega_consume<R, 1, 4, PRO_NONE, EPI_STORE>", "mega_consume<R, 1, PRO_NONE, EPI_STORE>")
rep("mega_consume<R, 1, 4, PRO_NORM, EPI_SILU>", "mega_consume<R, 1, PRO_NORM, EPI_SILU>")
rep("mega_consume<R, 1, 4, PRO_NORM, EPI_HEADS>", "mega_consume<R, 1, PRO_NORM, EPI_HEADS>")
rep('''      if (m.F == 8192)
        mega_consume<R, 2, 2, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);
      else
        mega_consume<R, 1, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                   m.xt, nullptr, nullptr, slot);''',
'''      if (m.F == 4 * kMegaKB)
        mega_consume<R, 4, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                m.xt, nullptr, nullptr, slot);
      else if (m.F == 2 * kMegaKB)
        mega_consume<R, 2, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                m.xt, nullptr, nullptr, slot);
      else
        mega_consume<R, 1, PRO_NONE, EPI_RESID>(a, ring, part, full_bar, empty_bar, red, S, gst, true, warp, lane, m.ht, TAG(ph - 1), m.xt, TAG(ph),
                                                m.xt, nullptr, nullptr, slot);''')

# host: supported shapes and partial-sum buffer
rep('''  auto k_ok = [](int K) { return K == 256 || K == 512 || K == 1024 || K == 2048 || K == 4096; };''',
    '''  auto k_ok = [](int K) { return K == 256 || K == 512 || K == 1024 || K == 2048; };   // one k-block, 128 k per warp''')
rep("k_ok(d.d_model) && k_ok(qn) && (k_ok(d.d_ff) || d.d_ff == 8192) &&", "k_ok(d.d_model) && k_ok(qn) && (k_ok(d.d_ff) || d.d_ff == 4096 || d.d_ff == 8192) &&")
a_=s.index("  // partial-sum buffer: the largest padded row count x k-slices over all matrices of the step")
b_=s.index("  pb = (pb + 1023) / 1024 * 1024;")
new='''  // partial-sum buffer: the largest padded row count x k-slices over all matrices of the step (it also stages the
  // activation fragments of every phase: 16 warps x R rows x 128 bf16)
  auto part_need = [&](int nunits, bool pairs, int K) {
    const int KS = std::min(K, kMegaKB) / kMegaWarpK;
    const int rows = ((nunits + grid - 1) / grid) * (pairs ? 2 : 1);
    return (size_t)((rows + kMegaRows - 1) / kMegaRows * kMegaRows) * KS * R * sizeof(float);
  };
  const int qn = d.n_heads * d.head_dim;
  size_t pb = part_need((d.n_heads + 2 * d.n_heads_kv) * d.head_dim / 2, true, d.d_model);
  pb = std::max(pb, part_need(d.d_model, false, qn));
  pb = std::max(pb, part_need(d.d_ff, true, d.d_model));
  pb = std::max(pb, part_need(d.d_model, false, d.d_ff));
  pb = std::max(pb, part_need(m.QV, false, d.d_model));
  pb = std::max(pb, (size_t)kW3 * R * kMegaWarpK * sizeof(bf16));
'''
s=s[:a_]+new+s[b_:]
open(p,'w').write(s)
print('ok')
