// DAC 44.1 kHz ENCODE: waveform -> 9 codebook indices per frame (the audio-prefix path).
//
// Replaces `DACAutoencoder.encode` (zonos/autoencoder.py:104-117) = transformers `DacModel.encode(wav).audio_codes`
// (modeling_dac.py:581-640): DacEncoder (:442-473: Conv1d 1->64 k7, four blocks of three residual units + Snake +
// strided Conv1d, Snake, Conv1d 1024->1024 k3) and the residual vector quantiser (:281-343, :122-170).  The reference
// runs this in fp32 without autocast, once per utterance (3 s of prefix audio = 0.19 TFLOP), and its output is INTEGER
// (nearest code): everything here stays in fp32 with fused multiply-adds in a fixed order, so a code can only differ
// from the reference's where two codes score within fp32 summation noise of each other (the oracle reports the margins).
// Activations are channels-first [B][C][L] like torch's, so time is the coalesced direction.
//
//   snake_kernel      x + (alpha + 1e-9)^-1 sin(alpha x)^2, elementwise (:85-99)
//   enc_conv_kernel   Conv1d with stride / dilation / zero padding, bias and optional residual add: one CTA per
//                     (64 output channels x 64 time steps x utterance), 8 input channels per shared-memory step
//   rvq_kernel        one codebook of the residual quantiser per launch, one CTA per (frame, utterance): in_proj
//                     (1024 -> 8), cosine-nearest code, out_proj of the straight-through value, residual update
#include "internal.h"

namespace {

constexpr int kCoT = 64, kTT = 64, kCiT = 8, kEncThreads = 256;

struct SnakeArgs { const float* x; const float* alpha; float* y; int C; long long L; long long total; };
__global__ void __launch_bounds__(256) snake_kernel(SnakeArgs a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / a.L) % a.C);
    const float al = a.alpha[c], v = a.x[i];
    const float s = sinf(al * v);
    a.y[i] = v + (1.0f / (al + 1e-9f)) * (s * s);
  }
}

struct EncConvArgs {
  const float* x; const float* w; const float* bias; const float* resid; float* y;
  int Cin, Cout, K, stride, dil, pad; long long Lin, Lout;
};

__global__ void __launch_bounds__(kEncThreads) enc_conv_kernel(EncConvArgs a) {
  extern __shared__ float smem_ec[];
  const int XW = (kTT - 1) * a.stride + (a.K - 1) * a.dil + 1;       // input window of one time tile
  float* xs = smem_ec;                                               // [kCiT][XW]
  float* ws = smem_ec + kCiT * XW;                                   // [kCoT][kCiT][K]
  const int b = blockIdx.z, co0 = blockIdx.y * kCoT;
  const long long t0 = (long long)blockIdx.x * kTT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;            // 16 time groups x 16 channel groups
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long in0 = t0 * a.stride - a.pad;                       // first input sample of the window
  const float* xb = a.x + (size_t)b * a.Cin * a.Lin;
  for (int c0 = 0; c0 < a.Cin; c0 += kCiT) {
    const int nci = min(kCiT, a.Cin - c0);
    for (int i = threadIdx.x; i < nci * XW; i += kEncThreads) {
      const int ci = i / XW, p = i - ci * XW;
      const long long s = in0 + p;
      xs[ci * XW + p] = (s >= 0 && s < a.Lin) ? xb[(size_t)(c0 + ci) * a.Lin + s] : 0.f;
    }
    for (int i = threadIdx.x; i < kCoT * nci * a.K; i += kEncThreads) {
      const int co = i / (nci * a.K), r = i - co * (nci * a.K), ci = r / a.K, kk = r - ci * a.K;
      ws[(co * kCiT + ci) * a.K + kk] = (co0 + co < a.Cout) ? a.w[((size_t)(co0 + co) * a.Cin + c0 + ci) * a.K + kk] : 0.f;
    }
    __syncthreads();
    for (int ci = 0; ci < nci; ++ci)
      for (int kk = 0; kk < a.K; ++kk) {
        float wv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = ws[((ty * 4 + i) * kCiT + ci) * a.K + kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xs[ci * XW + (tx + 16 * j) * a.stride + kk * a.dil];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
      }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.Cout) continue;
    const float bv = a.bias[co];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long t = t0 + tx + 16 * j;
      if (t >= a.Lout) continue;
      const size_t o = ((size_t)b * a.Cout + co) * a.Lout + t;
      float v = acc[i][j] + bv;
      if (a.resid) v = a.resid[o] + v;                               // hidden_state + output_tensor (:205)
      a.y[o] = v;
    }
  }
}

struct RvqArgs {
  float* residual;                 // [B][D][T], updated in place
  const float *in_w, *in_b, *cb, *out_w, *out_b;   // [8][D], [8], [V][8], [D][8], [D]
  int64_t* codes;                  // [B][Q][T]
  int q, Q, D, V, T;
};

// codebook_dim is 8 (descript/dac_44khz); the kernel is written for it
__global__ void __launch_bounds__(256) rvq_kernel(RvqArgs a) {
  __shared__ float red[8][8];
  __shared__ float s_proj[8], s_enc[8], s_qr[8];
  __shared__ float s_best[8];
  __shared__ int s_idx[8];
  const int t = blockIdx.x, b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* res = a.residual + (size_t)b * a.D * a.T + t;
  // in_proj: a 1x1 conv, D -> 8 (:143)
  float p[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) p[o] = 0.f;
  for (int c = threadIdx.x; c < a.D; c += 256) {
    const float v = res[(size_t)c * a.T];
#pragma unroll
    for (int o = 0; o < 8; ++o) p[o] = fmaf(a.in_w[(size_t)o * a.D + c], v, p[o]);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    p[o] = warp_sum(p[o]);
    if (lane == 0) red[o][warp] = p[o];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = a.in_b[threadIdx.x];
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    s_proj[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                            // F.normalize(encodings) (:161): x / max(|x|, 1e-12)
    float n2 = 0.f;
    for (int j = 0; j < 8; ++j) n2 = fmaf(s_proj[j], s_proj[j], n2);
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    for (int j = 0; j < 8; ++j) s_enc[j] = s_proj[j] * inv;
  }
  __syncthreads();
  float e[8], l2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { e[j] = s_enc[j]; l2 = fmaf(e[j], e[j], l2); }
  // dist = -(|e|^2 - 2 e.c) + |c|^2 on unit vectors, argmax, first index on ties (:165-168)
  float best = -INFINITY;
  int bidx = 0;
  for (int v = threadIdx.x; v < a.V; v += 256) {
    float c[8], n2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = a.cb[(size_t)v * 8 + j]; n2 = fmaf(c[j], c[j], n2); }
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    float dot = 0.f, cn2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float cn = c[j] * inv; dot = fmaf(e[j], cn, dot); cn2 = fmaf(cn, cn, cn2); }
    const float d = -(l2 - 2.0f * dot) + cn2;
    if (d > best) { best = d; bidx = v; }                            // ascending v per thread: the first maximum wins
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  if (lane == 0) { s_best[warp] = best; s_idx[warp] = bidx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bidx)) { best = s_best[w]; bidx = s_idx[w]; }
    a.codes[((size_t)b * a.Q + a.q) * a.T + t] = bidx;
    for (int j = 0; j < 8; ++j) {
      const float pj = s_proj[j], qj = a.cb[(size_t)bidx * 8 + j];
      s_qr[j] = pj + (qj - pj);                                      // straight-through expression, forward value (:146-147)
    }
  }
  __syncthreads();
  // out_proj (1x1 conv, 8 -> D) and residual update (:321)
  for (int c = threadIdx.x; c < a.D; c += 256) {
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o = fmaf(a.out_w[(size_t)c * 8 + j], s_qr[j], o);
    res[(size_t)c * a.T] -= o + a.out_b[c];
  }
}

}  // namespace

extern "C" ZB_API size_t zb_dac_encode_workspace_bytes(int32_t B, int64_t L) {
  // three ping-pong activation buffers of the widest layer (64 channels at full rate = 128 at half rate) + the latents
  const size_t act = (size_t)B * 64 * (size_t)L * sizeof(float);
  const size_t lat = (size_t)B * 1024 * (size_t)(L / 512) * sizeof(float);
  return 3 * ((act + 255) / 256 * 256) + ((lat + 255) / 256 * 256);
}

extern "C" ZB_API zb_status zb_dac_encode(zb_ctx* ctx, const zb_dac_enc_desc* d, const float* wav, int32_t B, int64_t L, int64_t* codes, void* workspace,
                                          size_t workspace_bytes, zb_stream stream_) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, d && wav && codes && workspace, "zb_dac_encode: null argument");
  ZB_REQUIRE(ctx, B >= 1 && L >= 512 && L % 512 == 0, "zb_dac_encode: %lld samples (a positive multiple of 512 expected, zonos/autoencoder.py:98-100)", (long long)L);
  ZB_REQUIRE(ctx, d->n_blocks == 4 && d->codebook_dim == 8 && d->latent_dim == 1024 && d->hidden == 64, "zb_dac_encode: not the descript/dac_44khz shape");
  ZB_REQUIRE(ctx, d->n_tensors == 2 + d->n_blocks * 21 + 3 + d->n_codebooks * 5, "zb_dac_encode: %d tensors", d->n_tensors);
  ZB_REQUIRE(ctx, workspace_bytes >= zb_dac_encode_workspace_bytes(B, L), "zb_dac_encode: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t act = ((size_t)B * 64 * (size_t)L * sizeof(float) + 255) / 256 * 256;
  float* buf[3] = {(float*)workspace, (float*)((char*)workspace + act), (float*)((char*)workspace + 2 * act)};
  float* lat = (float*)((char*)workspace + 3 * act);
  const float* const* T = d->tensors;
  int ti = 0;

  auto conv = [&](const float* x, float* y, const float* w, const float* bias, const float* resid, int Cin, int Cout, int K, int stride, int dil, int pad,
                  long long Lin) -> long long {
    EncConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.resid = resid; a.y = y; a.Cin = Cin; a.Cout = Cout; a.K = K; a.stride = stride; a.dil = dil; a.pad = pad; a.Lin = Lin;
    a.Lout = (Lin + 2 * pad - dil * (K - 1) - 1) / stride + 1;
    const int XW = (kTT - 1) * stride + (K - 1) * dil + 1;
    const size_t smem = ((size_t)kCiT * XW + (size_t)kCoT * kCiT * K) * sizeof(float);
    if (zb_ensure_smem(ctx, enc_conv_kernel, smem) != cudaSuccess) return -1;
    enc_conv_kernel<<<dim3((unsigned)((a.Lout + kTT - 1) / kTT), (Cout + kCoT - 1) / kCoT, B), kEncThreads, smem, stream>>>(a);
    ctx->launches++;
    return a.Lout;
  };
  auto snake = [&](const float* x, float* y, const float* alpha, int C, long long Lc) {
    SnakeArgs a;
    a.x = x; a.alpha = alpha; a.y = y; a.C = C; a.L = Lc; a.total = (long long)B * C * Lc;
    snake_kernel<<<(unsigned)std::min<long long>((a.total + 255) / 256, 148 * 16), 256, 0, stream>>>(a);
    ctx->launches++;
  };

  long long Lc = L;
  int ch = d->hidden;
  // encoder.conv1: the waveform is [B][1][L]
  if (conv(wav, buf[0], T[ti], T[ti + 1], nullptr, 1, ch, 7, 1, 1, 3, Lc) < 0) return zb_fail(ctx, ZB_ERR_CUDA, "zb_dac_encode: shared memory");
  ti += 2;
  float *x = buf[0], *ta = buf[1], *tb = buf[2];
  const int dils[3] = {1, 3, 9};
  for (int blk = 0; blk < d->n_blocks; ++blk) {
    const int s = d->strides[blk];
    for (int j = 0; j < 3; ++j) {                                    // residual unit (:187-207)
      snake(x, ta, T[ti], ch, Lc);
      conv(ta, tb, T[ti + 1], T[ti + 2], nullptr, ch, ch, 7, 1, dils[j], 3 * dils[j], Lc);
      snake(tb, ta, T[ti + 3], ch, Lc);
      conv(ta, x, T[ti + 4], T[ti + 5], x, ch, ch, 1, 1, 1, 0, Lc); // in place: every output reads only its own residual element
      ti += 6;
    }
    snake(x, ta, T[ti], ch, Lc);
    const long long Ln = conv(ta, tb, T[ti + 1], T[ti + 2], nullptr, ch, 2 * ch, 2 * s, s, 1, (s + 1) / 2, Lc);
    ti += 3;
    std::swap(x, tb);
    Lc = Ln; ch *= 2;
  }
  ZB_REQUIRE(ctx, Lc == L / 512 && ch == 1024, "zb_dac_encode: internal shape error (%lld frames, %d channels)", Lc, ch);
  snake(x, ta, T[ti], ch, Lc);
  conv(ta, lat, T[ti + 1], T[ti + 2], nullptr, ch, d->latent_dim, 3, 1, 1, 1, Lc);
  ti += 3;
  for (int q = 0; q < d->n_codebooks; ++q) {
    RvqArgs r;
    r.residual = lat; r.in_w = T[ti]; r.in_b = T[ti + 1]; r.cb = T[ti + 2]; r.out_w = T[ti + 3]; r.out_b = T[ti + 4];
    r.codes = codes; r.q = q; r.Q = d->n_codebooks; r.D = d->latent_dim; r.V = d->codebook_size; r.T = (int)Lc;
    rvq_kernel<<<dim3((unsigned)Lc, B), 256, 0, stream>>>(r);
    ctx->launches++;
    ti += 5;
  }
  ZB_CUDA(ctx, cudaGetLastError());
  return ZB_OK;
}
