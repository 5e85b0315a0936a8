// DAC 44.1 kHz decoder: codes -> waveform.
//
// Replaces DACAutoencoder.decode (zonos/autoencoder.py:119-140), whose arithmetic lives in the un-vendored
// dependency `transformers` (modeling_dac.py:345-369 from_codes, :85-99 Snake1d, :173-207 residual unit,
// :234-262 decoder block, :405-439 decoder).
//
// Design: activations are channels-last bf16 [B][L][C] (the K-major operand layout of an implicit GEMM);
// every Conv1d / ConvTranspose1d is ONE implicit-GEMM launch
//     out_row[j][n] = sum_tap sum_ci in[j + shift(tap)][ci] * W[tap][n][ci]
// (a stride-s transposed conv with kernel 2s is the 2-tap GEMM with n = phase*Cout + co), with bias,
// residual add, the NEXT layer's Snake activation and the bf16 stores fused into the epilogue, so no
// elementwise kernel ever touches HBM.  The 9 codebook lookups + 1x1 projections are one gather-sum over a
// table precomputed at create time.  Numerics follow the reference's CUDA autocast path: bf16 conv operands,
// fp32 accumulation, fp32 Snake, bf16 residual stream.
#include <algorithm>

#include "internal.h"

namespace {

struct ConvArgs {
  const bf16* in;      // [B][Lin][Cin]
  const bf16* w;       // [taps][N][Cin]
  const float* bias;   // [Cout]
  int B, Lin, Cin, Cout, N, taps, dil, pad, ups, rows, Lout;
  const bf16* resid;   // [B][Lout][Cout] or null
  bf16* out_raw;       // [B][Lout][Cout] or null
  bf16* out_act;       // [B][Lout][Cout] or null: snake(out, alpha)
  const float* alpha;  // [Cout] or null (identity)
};

__device__ __forceinline__ float snake_f(float x, float alpha) {
  const float s = sinf(alpha * x);
  return x + (1.0f / (alpha + 1e-9f)) * (s * s);
}

constexpr int BM = 64, BN = 64, BK = 32;

// SIMT implicit GEMM (first correct path; the tcgen05 version keeps this interface).
__global__ void __launch_bounds__(256) conv_gemm_kernel(ConvArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bf16* inb = a.in + (size_t)b * a.Lin * a.Cin;
  for (int tap = 0; tap < a.taps; ++tap) {
    const int shift = a.ups ? -tap : tap * a.dil - a.pad;
    const bf16* wt = a.w + (size_t)tap * a.N * a.Cin;
    for (int c0 = 0; c0 < a.Cin; c0 += BK) {
      // A tile: 64 rows x 32 channels, B tile: 64 cols x 32 channels (8 bf16 per thread each)
      {
        const int row = threadIdx.x / 4, c8 = (threadIdx.x % 4) * 8;
        const int jj = j0 + row + shift;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (jj >= 0 && jj < a.Lin && j0 + row < a.rows) v = *reinterpret_cast<const uint4*>(inb + (size_t)jj * a.Cin + c0 + c8);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { As[c8 + 2 * q][row] = bf16lo(w4[q]); As[c8 + 2 * q + 1][row] = bf16hi(w4[q]); }
        const int n = n0 + row;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (n < a.N) u = *reinterpret_cast<const uint4*>(wt + (size_t)n * a.Cin + c0 + c8);
        const uint32_t u4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { Bs[c8 + 2 * q][row] = bf16lo(u4[q]); Bs[c8 + 2 * q + 1][row] = bf16hi(u4[q]); }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- fused epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty * 4 + i;
    if (j >= a.rows) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int n = n0 + tx * 4 + jj;
      if (n >= a.N) continue;
      int time = j, co = n;
      if (a.ups) { const int ph = n / a.Cout; co = n % a.Cout; time = a.ups * j + ph - a.pad; }
      if (time < 0 || time >= a.Lout) continue;
      const size_t o = ((size_t)b * a.Lout + time) * a.Cout + co;
      float v = rbf(acc[i][jj] + a.bias[co]);                       // conv output is bf16 under autocast
      if (a.resid) v = rbf(bf2f(a.resid[o]) + v);                   // bf16 residual add (modeling_dac.py:206)
      if (a.out_raw) a.out_raw[o] = f2bf(v);
      if (a.out_act) a.out_act[o] = f2bf(a.alpha ? snake_f(v, a.alpha[co]) : v);
    }
  }
}

// z[b][t][:] = sum_k table[k][codes[b][k][t]][:]   (fp32 sequential sum, bf16 store)
__global__ void __launch_bounds__(256) from_codes_kernel(const int64_t* codes, const float* table, int B, int Q, int T,
                                                         int C, int vocab, bf16* z) {
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < Q; ++k) {
      long long id = codes[((size_t)b * Q + k) * T + t];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
      s += table[((size_t)k * vocab + id) * C + c];
    }
    z[((size_t)b * T + t) * C + c] = f2bf(s);
  }
}

// table[k][code][c] = bias_k[c] + sum_d W_k[c][d] * E_k[code][d]
__global__ void build_table_kernel(const float* E, const float* W, const float* bias, int vocab, int dim, int C, float* table) {
  const int code = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int d = 0; d < dim; ++d) s = fmaf(W[(size_t)c * dim + d], E[(size_t)code * dim + d], s);
    table[(size_t)code * C + c] = s + bias[c];
  }
}

// w_out[tap][n][ci] (bf16) from torch Conv1d weight [Cout][Cin][K]
__global__ void relayout_conv_kernel(const float* w, int Cout, int Cin, int K, bf16* out) {
  const size_t total = (size_t)K * Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % Cin; const int co = (i / Cin) % Cout; const int k = i / ((size_t)Cin * Cout);
    out[i] = f2bf(w[((size_t)co * Cin + ci) * K + k]);
  }
}
// w_out[tap][ph*Cout+co][ci] from ConvTranspose1d weight [Cin][Cout][2s]; tap 0 <-> k = ph, tap 1 <-> k = ph + s
__global__ void relayout_convT_kernel(const float* w, int Cin, int Cout, int s, bf16* out) {
  const size_t total = (size_t)2 * s * Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % Cin; const size_t r = i / Cin; const int n = r % ((size_t)s * Cout); const int tap = r / ((size_t)s * Cout);
    const int ph = n / Cout, co = n % Cout;
    out[i] = f2bf(w[((size_t)ci * Cout + co) * (2 * s) + ph + tap * s]);
  }
}

// final Conv1d(C -> 1, k7, pad 3) + tanh  (one warp per output sample)
__global__ void __launch_bounds__(256) final_conv_kernel(const bf16* in, const bf16* w /*[7][C]*/, const float* bias, int B, int L,
                                                         int C, float* wav) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int b = warp / L, t = warp % L;
  float s = 0.f;
  for (int k = 0; k < 7; ++k) {
    const int tt = t + k - 3;
    if (tt < 0 || tt >= L) continue;
    const bf16* row = in + ((size_t)b * L + tt) * C;
    for (int c = lane; c < C; c += 32) s = fmaf(bf2f(row[c]), bf2f(w[k * C + c]), s);
  }
  s = warp_sum(s);
  if (lane == 0) wav[(size_t)b * L + t] = tanhf(rbf(s + bias[0]));   // conv out bf16 (autocast); tanh kept in fp32
}

}  // namespace

struct zb_conv_w { bf16* w = nullptr; float* bias = nullptr; int Cin = 0, Cout = 0, taps = 0, dil = 1, pad = 0, ups = 0; };
struct zb_dac {
  zb_ctx* ctx;
  zb_dac_desc d;
  float* table = nullptr;              // [Q][vocab][latent]
  zb_conv_w conv1, conv2;              // decoder.conv1 / decoder.conv2
  struct Block { float* snake1; zb_conv_w convT; float* ru_s1[3]; zb_conv_w ru_c1[3]; float* ru_s2[3]; zb_conv_w ru_c2[3]; int stride; };
  std::vector<Block> blocks;
  float* final_alpha = nullptr;
  std::vector<void*> owned;
};

namespace {
zb_status dev_alloc(zb_dac* d, void** p, size_t bytes) {
  ZB_CUDA(d->ctx, cudaMalloc(p, bytes));
  d->owned.push_back(*p);
  return ZB_OK;
}
zb_status copy_f32(zb_dac* d, const float* src, size_t n, float** out, cudaStream_t s) {
  if (zb_status st = dev_alloc(d, (void**)out, n * 4)) return st;
  ZB_CUDA(d->ctx, cudaMemcpyAsync(*out, src, n * 4, cudaMemcpyDeviceToDevice, s));
  return ZB_OK;
}
zb_status make_conv(zb_dac* d, const float* w, const float* b, int Cout, int Cin, int K, int dil, zb_conv_w* o, cudaStream_t s) {
  o->Cin = Cin; o->Cout = Cout; o->taps = K; o->dil = dil; o->pad = (K - 1) * dil / 2; o->ups = 0;
  if (zb_status st = dev_alloc(d, (void**)&o->w, (size_t)K * Cout * Cin * 2)) return st;
  relayout_conv_kernel<<<256, 256, 0, s>>>(w, Cout, Cin, K, o->w);
  ZB_CHECK_LAUNCH(d->ctx);
  return copy_f32(d, b, Cout, &o->bias, s);
}
zb_status make_convT(zb_dac* d, const float* w, const float* b, int Cin, int Cout, int stride, zb_conv_w* o, cudaStream_t s) {
  o->Cin = Cin; o->Cout = Cout; o->taps = 2; o->dil = 1; o->pad = (stride + 1) / 2; o->ups = stride;
  if (zb_status st = dev_alloc(d, (void**)&o->w, (size_t)2 * stride * Cout * Cin * 2)) return st;
  relayout_convT_kernel<<<256, 256, 0, s>>>(w, Cin, Cout, stride, o->w);
  ZB_CHECK_LAUNCH(d->ctx);
  return copy_f32(d, b, Cout, &o->bias, s);
}

zb_status run_conv(zb_ctx* ctx, const zb_conv_w& c, const bf16* in, int B, int Lin, const bf16* resid, bf16* out_raw, bf16* out_act,
                   const float* alpha, cudaStream_t s) {
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.B = B; a.Lin = Lin; a.Cin = c.Cin; a.Cout = c.Cout; a.taps = c.taps; a.dil = c.dil;
  a.pad = c.pad; a.ups = c.ups;
  if (c.ups) { a.N = c.ups * c.Cout; a.rows = Lin + 1; a.Lout = Lin * c.ups; }
  else { a.N = c.Cout; a.rows = Lin; a.Lout = Lin; }
  a.resid = resid; a.out_raw = out_raw; a.out_act = out_act; a.alpha = alpha;
  dim3 grid((a.rows + BM - 1) / BM, (a.N + BN - 1) / BN, B);
  conv_gemm_kernel<<<grid, 256, 0, s>>>(a);
  ZB_CHECK_LAUNCH(ctx);
  return ZB_OK;
}
}  // namespace

extern "C" zb_status zb_dac_create(zb_ctx* ctx, const zb_dac_desc* desc, zb_dac** out, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  ZB_REQUIRE(ctx, desc && out, "zb_dac_create: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int Q = desc->n_codebooks, nb = desc->n_blocks;
  const int expect = 3 * Q + 2 + nb * (3 + 3 * 6) + 1 + 2;
  ZB_REQUIRE(ctx, desc->n_tensors == expect, "zb_dac_create: expected %d tensors, got %d", expect, desc->n_tensors);
  ZB_REQUIRE(ctx, nb >= 1 && nb <= 8 && desc->channels % (1 << nb) == 0, "zb_dac_create: bad block structure");
  zb_dac* d = new zb_dac();
  d->ctx = ctx; d->d = *desc;
  const float* const* t = desc->tensors;
  int ti = 0;
  zb_status st = ZB_OK;
  auto fail = [&](zb_status e) { zb_dac_destroy(d); return e; };
  // quantizer: per codebook (codebook.weight [vocab,dim], out_proj.weight [latent,dim,1], out_proj.bias [latent])
  const size_t tab_per = (size_t)desc->codebook_size * desc->latent_dim;
  if ((st = dev_alloc(d, (void**)&d->table, (size_t)Q * tab_per * 4))) return fail(st);
  for (int k = 0; k < Q; ++k) {
    build_table_kernel<<<desc->codebook_size, 256, 0, s>>>(t[ti], t[ti + 1], t[ti + 2], desc->codebook_size, desc->codebook_dim,
                                                          desc->latent_dim, d->table + (size_t)k * tab_per);
    ti += 3;
  }
  int ch = desc->channels;
  if ((st = make_conv(d, t[ti], t[ti + 1], ch, desc->latent_dim, 7, 1, &d->conv1, s))) return fail(st);
  ti += 2;
  for (int i = 0; i < nb; ++i) {
    zb_dac::Block blk;
    blk.stride = desc->strides[i];
    if ((st = copy_f32(d, t[ti], ch, &blk.snake1, s))) return fail(st);
    if ((st = make_convT(d, t[ti + 1], t[ti + 2], ch, ch / 2, blk.stride, &blk.convT, s))) return fail(st);
    ti += 3;
    ch /= 2;
    const int dil[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) {
      if ((st = copy_f32(d, t[ti], ch, &blk.ru_s1[j], s))) return fail(st);
      if ((st = make_conv(d, t[ti + 1], t[ti + 2], ch, ch, 7, dil[j], &blk.ru_c1[j], s))) return fail(st);
      if ((st = copy_f32(d, t[ti + 3], ch, &blk.ru_s2[j], s))) return fail(st);
      if ((st = make_conv(d, t[ti + 4], t[ti + 5], ch, ch, 1, 1, &blk.ru_c2[j], s))) return fail(st);
      ti += 6;
    }
    d->blocks.push_back(blk);
  }
  if ((st = copy_f32(d, t[ti], ch, &d->final_alpha, s))) return fail(st);
  ti += 1;
  if ((st = make_conv(d, t[ti], t[ti + 1], 1, ch, 7, 1, &d->conv2, s))) return fail(st);
  *out = d;
  return ZB_OK;
}

extern "C" zb_status zb_dac_destroy(zb_dac* dac) {
  if (!dac) return ZB_OK;
  for (void* p : dac->owned) cudaFree(p);
  delete dac;
  return ZB_OK;
}

extern "C" zb_status zb_dac_decode(zb_ctx* ctx, const zb_dac* dac, const int64_t* codes, int32_t B, int32_t T, float* wav,
                                   zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  ZB_REQUIRE(ctx, dac && codes && wav && B >= 1 && T >= 1, "zb_dac_decode: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const zb_dac_desc& d = dac->d;
  // activation buffers: the widest layer in elements is max over stages of L*C; with strides (8,8,4,2) and channel
  // halving that is the last stage: L = 512 T, C = 96  (and the first: T x 1536 is far smaller)
  int up = 1;
  size_t max_elems = (size_t)T * d.channels;
  { int ch = d.channels; for (int i = 0; i < d.n_blocks; ++i) { up *= d.strides[i]; ch /= 2; max_elems = std::max(max_elems, (size_t)T * up * ch); } }
  max_elems = std::max(max_elems, (size_t)T * d.latent_dim);
  const size_t buf = (max_elems * B * 2 + 255) / 256 * 256;
  if (zb_status st = zb_dac_scratch_reserve(ctx, 4 * buf)) return st;
  bf16* act = (bf16*)ctx->dac_scratch;                 // snake-activated input of the next conv
  bf16* raw = (bf16*)((char*)ctx->dac_scratch + buf);  // residual stream
  bf16* tmp = (bf16*)((char*)ctx->dac_scratch + 2 * buf);
  bf16* act2 = (bf16*)((char*)ctx->dac_scratch + 3 * buf);
  from_codes_kernel<<<B * T, 256, 0, s>>>(codes, dac->table, B, d.n_codebooks, T, d.latent_dim, d.codebook_size, tmp);
  ZB_CHECK_LAUNCH(ctx);
  zb_status st;
  int L = T;
  // decoder.conv1, epilogue applies block0.snake1
  if ((st = run_conv(ctx, dac->conv1, tmp, B, L, nullptr, nullptr, act, dac->blocks[0].snake1, s))) return st;
  for (size_t i = 0; i < dac->blocks.size(); ++i) {
    const zb_dac::Block& blk = dac->blocks[i];
    // transposed conv -> raw x and snake(res_unit1.snake1)
    if ((st = run_conv(ctx, blk.convT, act, B, L, nullptr, raw, act2, blk.ru_s1[0], s))) return st;
    L *= blk.stride;
    bf16* cur_act = act2;
    bf16* other = act;
    for (int j = 0; j < 3; ++j) {
      // conv1 (k7, dilated) -> snake2
      if ((st = run_conv(ctx, blk.ru_c1[j], cur_act, B, L, nullptr, nullptr, tmp, blk.ru_s2[j], s))) return st;
      // conv2 (k1) + residual -> raw (in place) and the next snake
      const float* next_alpha = (j < 2) ? blk.ru_s1[j + 1] : (i + 1 < dac->blocks.size() ? dac->blocks[i + 1].snake1 : dac->final_alpha);
      if ((st = run_conv(ctx, blk.ru_c2[j], tmp, B, L, raw, raw, other, next_alpha, s))) return st;
      std::swap(cur_act, other);
    }
    if (cur_act != act) {   // keep the invariant: `act` holds the input of the next stage
      bf16* t2 = act; act = cur_act; act2 = t2;
    }
  }
  const int C = dac->conv2.Cin;
  const long long warps = (long long)B * L;
  final_conv_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(act, dac->conv2.w, dac->conv2.bias, B, L, C, wav);
  ZB_CHECK_LAUNCH(ctx);
  return ZB_OK;
}
