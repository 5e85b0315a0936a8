p='zonos_b200/csrc/decode.cu'; s=open(p).read()
# ---- producer: k-block outer, row group inner
a_=s.index("  const bf16* src = nullptr;\n  for (int gi = 0; gi < ngroup; ++gi) {\n    if (lane < kMegaRows) src = a.W")
b_=s.index("// consumers: one matrix phase on the tensor cores")
new='''  for (int kb = 0; kb < NB; ++kb)
    for (int gi = 0; gi < ngroup; ++gi, ++gst) {
      const int slot = gst % S;
      if (gst >= S) mbar_wait(&empty_bar[slot], ((gst / S) - 1) & 1);
      if (lane == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)(kMegaRows * KB * 2));
      __syncwarp();
      if (lane < kMegaRows) {
        const bf16* src = a.W + (size_t)row_of_local<EPI>(a, u_begin, min(gi * kMegaRows + lane, nrows - 1)) * K + (size_t)kb * KB;
        bulk_g2s(ring + (size_t)slot * kMegaStageBytes + (size_t)lane * pitch, src, (uint32_t)KB * 2, &full_bar[slot], pol);
      }
    }
}

'''
s=s[:a_]+new+s[b_:]
s=s.replace("// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst): for every group of\n// kMegaRows rows, one stage per k-block",
            "// producer: stream this CTA's slice of one matrix through the ring (global stage counter gst): one stage per\n// (k-block, group of kMegaRows rows), k-block outermost")

# ---- consumer body between the norm-parameter loads and the epilogue-operand prefetch
a_=s.index("  uint2 nwr[NB], nbr[NB];                                       // norm parameters")
b_=s.index("  // ---- operands of this thread's epilogue (residual value, RoPE cos/sin, KV page): fetched NOW")
new=r'''  static_assert(NB == 1 || PRO == PRO_NONE, "the norm prologue covers one k-block");
  const int g = lane >> 2, c = lane & 3;
  bf16* stg = reinterpret_cast<bf16*>(part) + (size_t)warp * R * kMegaWarpK;    // per-warp fragment staging (aliases `part`)
  uint2 nwr = make_uint2(0, 0), nbr = make_uint2(0, 0);          // norm parameters: in flight together with the activations
  if (PRO == PRO_NORM && active) {
    nwr = *reinterpret_cast<const uint2*>(a.nw + kw);
    if (a.nb) nbr = *reinterpret_cast<const uint2*>(a.nb + kw);
  }

  // activations of k-block 0: spin on the operand loads themselves until every word carries the producing phase's tag
  uint4 raw[R];
  auto load_block = [&](int kb) {
#pragma unroll
    for (int i = 0; i < R; ++i) raw[i] = ld_relaxed_v4(xt + (size_t)i * a.ldx + (size_t)kb * KB + kw);
  };
  auto block_ok = [&]() {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < R; ++i) ok = ok && tags_ok(raw[i], tag_in);
    return ok;
  };
  auto poll_block = [&](int kb) {
    for (unsigned spins = 0; !block_ok(); ++spins) {
      if (spins > kMegaSpinLimit) asm volatile("trap;");
      load_block(kb);
    }
  };
  if (active) { load_block(0); poll_block(0); }
  if (stamp && threadIdx.x == 0) *stamp = gtime();

  uint2 xp[R];                                                  // this lane's 4 elements per row as packed bf16 pairs
  if (PRO == PRO_NORM) {
    float xf[R][4], mean[R], rstd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { xf[i][0] = untag(raw[i].x); xf[i][1] = untag(raw[i].y); xf[i][2] = untag(raw[i].z); xf[i][3] = untag(raw[i].w); }
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
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float tot = 0.f, tsq = 0.f;
      for (int q = 0; q < KS; ++q) { tot += red[0][q][i]; tsq += red[1][q][i]; }
      const float mu = tot / (float)K;
      mean[i] = (a.norm_kind == ZB_NORM_LAYERNORM) ? mu : 0.f;
      const float var = (a.norm_kind == ZB_NORM_LAYERNORM) ? fmaxf(tsq / (float)K - mu * mu, 0.f) : tsq / (float)K;
      rstd[i] = rsqrtf(var + a.eps);
    }
    const float g4[4] = {bf16lo(nwr.x), bf16hi(nwr.x), bf16lo(nwr.y), bf16hi(nwr.y)};
    const float b4[4] = {bf16lo(nbr.x), bf16hi(nbr.x), bf16lo(nbr.y), bf16hi(nbr.y)};
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float y4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) y4[e] = (xf[i][e] - mean[i]) * rstd[i] * g4[e] + b4[e];
      xp[i] = make_uint2(pack_bf16(y4[0], y4[1]), pack_bf16(y4[2], y4[3]));
    }
  } else {
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // `part` (staging below) is free: everyone left the previous epilogue
#pragma unroll
    for (int i = 0; i < R; ++i) xp[i] = make_uint2((raw[i].x >> 16) | (raw[i].y & 0xffff0000u), (raw[i].z >> 16) | (raw[i].w & 0xffff0000u));
  }

  // A fragments of one k-block: through the staging area from "4 consecutive k per lane" to the MMA layout
  // (lane = 4*row + c holds k = 16s + 2c, +1 and 16s + 2c + 8, +9 of every 16-k step s)
  uint32_t afr[NSTEP][2];
  auto build_fragments = [&]() {
    if (active) {
#pragma unroll
      for (int i = 0; i < R; ++i) *reinterpret_cast<uint2*>(stg + i * kMegaWarpK + lane * 4) = xp[i];
    }
    __syncwarp();
#pragma unroll
    for (int st = 0; st < NSTEP; ++st) {
      afr[st][0] = (active && g < R) ? *reinterpret_cast<const uint32_t*>(stg + g * kMegaWarpK + st * 16 + 2 * c) : 0u;
      afr[st][1] = (active && g < R) ? *reinterpret_cast<const uint32_t*>(stg + g * kMegaWarpK + st * 16 + 2 * c + 8) : 0u;
    }
    __syncwarp();
  };
  build_fragments();

'''
s=s[:a_]+new+s[b_:]

# ---- stage loop
a_=s.index("  asm volatile(\"bar.sync 1, %0;\" ::\"n\"(kW3 * 32) : \"memory\");   // every warp has its fragments: `part` may take partial sums")
b_=s.index("  gst += ngroup * NB;\n")+len("  gst += ngroup * NB;\n")
new=r'''  if (NB == 1) asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // every warp has its fragments: `part` may take partial sums

  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t lane_base = smem_u32(ring) + (uint32_t)(lane & 7) * pitch + (uint32_t)(warp * kMegaWarpK + (lane >> 3) * 8) * 2;
  // slot / parity / addresses advance incrementally (no division in the stage loop)
  uint32_t slot = (uint32_t)(gst % S), phase = (uint32_t)((gst / S) & 1);
  uint32_t src = lane_base + slot * kMegaStageBytes, fb = full0 + slot * 8, eb = empty0 + slot * 8;
  const bool storer = active && g < R;
  float* dst = part + ((size_t)(2 * c) * KS + warp) * R + g;   // D[row g][n = 2c, 2c+1] of the current row group
  const int dst_row = KS * R, dst_group = kMegaRows * KS * R;
  auto stage_mma = [&](float (&d0)[4], float (&d1)[4]) {
    mbar_wait_u32(fb, phase);
    if (active) {
#pragma unroll
      for (int j = 0; j < NSTEP / 2; ++j) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(src + j * 64, b0, b1, b2, b3);
        mma_16816(d0, afr[2 * j][0], afr[2 * j][1], b0, b1);
        mma_16816(d1, afr[2 * j + 1][0], afr[2 * j + 1][1], b2, b3);
      }
    }
    __syncwarp();
    if (release && lane == 0) mbar_arrive_u32(eb);
    if (++slot == (uint32_t)S) { slot = 0; phase ^= 1u; src = lane_base; fb = full0; eb = empty0; }
    else { src += kMegaStageBytes; fb += 8; eb += 8; }
  };
  if (NB == 1) {
    for (int gi = 0; gi < ngroup; ++gi) {
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      stage_mma(d0, d1);
      if (storer) { dst[0] = d0[0] + d1[0]; dst[dst_row] = d0[1] + d1[1]; }
      dst += dst_group;
    }
  } else {
    // several k-blocks (fc2): k-block outermost, so one block's fragments are live at a time; the (at most kMaxG) row
    // groups keep their accumulators across the blocks.  The next block's activations are fetched while this one runs.
    constexpr int kMaxG = 2;
    float d0[kMaxG][4], d1[kMaxG][4];
#pragma unroll
    for (int q = 0; q < kMaxG; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) { d0[q][e] = 0.f; d1[q][e] = 0.f; }
    for (int kb = 0; kb < NB; ++kb) {
      if (kb > 0) build_fragments();
      if (active && kb + 1 < NB) load_block(kb + 1);
#pragma unroll
      for (int q = 0; q < kMaxG; ++q)
        if (q < ngroup) stage_mma(d0[q], d1[q]);
      if (active && kb + 1 < NB) {
        poll_block(kb + 1);
#pragma unroll
        for (int i = 0; i < R; ++i) xp[i] = make_uint2((raw[i].x >> 16) | (raw[i].y & 0xffff0000u), (raw[i].z >> 16) | (raw[i].w & 0xffff0000u));
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kW3 * 32) : "memory");   // all fragment staging is over: `part` may take partial sums
#pragma unroll
    for (int q = 0; q < kMaxG; ++q)
      if (q < ngroup && storer) { dst[q * dst_group] = d0[q][0] + d1[q][0]; dst[q * dst_group + dst_row] = d0[q][1] + d1[q][1]; }
  }
  gst += ngroup * NB;
'''
s=s[:a_]+new+s[b_:]
# remove the now-duplicated g/c declaration if any remained after the epilogue prefetch block
s=s.replace("  const int g = lane >> 2, c = lane & 3;\n  // slot / parity / addresses advance incrementally","  // slot / parity / addresses advance incrementally")
open(p,'w').write(s)
print("ok")
