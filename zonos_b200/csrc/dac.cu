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
#include <stdlib.h>

#include <algorithm>

#include "tc.cuh"

namespace {

struct ConvArgs {
  const bf16* in;      // [B][Lin][Cin]
  const bf16* w;       // [taps][N][Cin]
  const float* bias;   // [Cout]
  int B, Lin, Cin, Cout, N, taps, dil, pad, ups, rows, Lout;
  int in_ld, out_ld;   // channel strides (channels padded to a multiple of 64 for the 128-byte-swizzled TMA boxes)
  const bf16* resid;   // [B][Lout][out_ld] or null
  bf16* out_raw;       // [B][Lout][out_ld] or null
  bf16* out_act;       // [B][Lout][out_ld] or null: snake(out, alpha)
  const float* alpha;  // [Cout] or null (identity)
};

// Snake1d (modeling_dac.py:85-99): x + sin(alpha x)^2 / (alpha + 1e-9), fp32.  sin^2 has period pi: two-term
// Cody-Waite reduction (exact product k * pi_hi for |k| < 2^11, error ~ k * 1e-15 beyond) + the SFU sine on [-pi/2, pi/2]
// (abs error 2^-21, far below the bf16 rounding that follows).  sinf()'s slow path (arguments > 1e5, local memory) made
// the epilogue the bottleneck of the decoder on saturated activations; the reciprocal is IEEE-rounded like torch's.
// inv = __frcp_rn(alpha + 1e-9f): a per-channel constant, computed once per CTA by the tensor-core kernel's epilogue
__device__ __forceinline__ float snake_inv(float x, float alpha, float inv) {
  const float y = alpha * x;
  const float k = rintf(y * 0.318309886183790672f);
  float r = fmaf(-k, 3.140625f, y);                            // pi = 3.140625 + 9.67653589793e-4 (hi has 8 significant bits)
  r = fmaf(-k, 9.67653589793e-4f, r);
  const float sn = __sinf(r);
  return x + inv * (sn * sn);
}
__device__ __forceinline__ float snake_f(float x, float alpha) { return snake_inv(x, alpha, __frcp_rn(alpha + 1e-9f)); }

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
  const bf16* inb = a.in + (size_t)b * a.Lin * a.in_ld;
  for (int tap = 0; tap < a.taps; ++tap) {
    const int shift = a.ups ? -tap : tap * a.dil - a.pad;
    const bf16* wt = a.w + (size_t)tap * a.N * a.in_ld;
    for (int c0 = 0; c0 < a.in_ld; c0 += BK) {
      // A tile: 64 rows x 32 channels, B tile: 64 cols x 32 channels (8 bf16 per thread each)
      {
        const int row = threadIdx.x / 4, c8 = (threadIdx.x % 4) * 8;
        const int jj = j0 + row + shift;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (jj >= 0 && jj < a.Lin && j0 + row < a.rows) v = *reinterpret_cast<const uint4*>(inb + (size_t)jj * a.in_ld + c0 + c8);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { As[c8 + 2 * q][row] = bf16lo(w4[q]); As[c8 + 2 * q + 1][row] = bf16hi(w4[q]); }
        const int n = n0 + row;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (n < a.N) u = *reinterpret_cast<const uint4*>(wt + (size_t)n * a.in_ld + c0 + c8);
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
      const size_t o = ((size_t)b * a.Lout + time) * a.out_ld + co;
      float v = rbf(acc[i][jj] + a.bias[co]);                       // conv output is bf16 under autocast
      if (a.resid) v = rbf(bf2f(a.resid[o]) + v);                   // bf16 residual add (modeling_dac.py:206)
      if (a.out_raw) a.out_raw[o] = f2bf(v);
      if (a.out_act) a.out_act[o] = f2bf(a.alpha ? snake_f(v, a.alpha[co]) : v);
    }
  }
}


// ------------------------------------------------------------------ tcgen05 implicit-GEMM convolution ---------
// D[128 time rows x BN output columns] (TMEM) = sum_tap sum_cin-chunk A_tap[128 x 64] * W_tap[BN x 64]^T.
// A is fetched by a 3-D TMA box {64 channels, 128 time steps, 1 batch} at time offset j0 + shift(tap): rows outside
// [0, Lin) are zero-filled by the TMA unit, which IS the convolution's zero padding (per batch element).
//   warp 4: TMA producer | warp 5: TMEM alloc + tcgen05.mma issuer | warps 0-3: epilogue (lane <-> time row)
// Halo mode (stride-1 convs with several taps): the taps of one 64-channel chunk read the SAME time rows shifted by
// tap * dilation, so the chunk's rows [j0 - pad, j0 - pad + 128 + (taps - 1) * dil) are fetched ONCE (box {64, halo_rows, 1},
// double-buffered) and every tap's A operand is that tile with the descriptor's start address moved down tap * dil rows: a
// K-major SWIZZLE_128B operand may start at any row, the swizzle being a function of the absolute shared-memory address
// (scripts/probes/desc_shift_probe.cu).  The ring then only carries the weight tiles: a k = 7 conv was bound by the
// L2 -> SM port with 7 x 16 KB of activations + 7 weight tiles per chunk (ncu, profiles/r2_ncu_full_conv_tc.txt).
constexpr int CV_THREADS = 192;
struct ConvTcArgs {
  CUtensorMap map_in;   // activations [B][Lin][in_ld], box {64, 128, 1}
  CUtensorMap map_w;    // weights [taps*N][in_ld], box {64, BN}
  ConvArgs c;
  int BN, stages;
  int halo_rows;        // > 0: "halo" mode, see conv_tc_kernel
};

__global__ void __launch_bounds__(CV_THREADS, 3) conv_tc_kernel(const __grid_constant__ ConvTcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem_cv[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar, hfull[2], hempty[2];
  __shared__ uint32_t tmem_base_smem;
  // per-column constants of this CTA's output tile (plain convs): bias, Snake alpha and its IEEE reciprocal - fetched and
  // computed once per CTA while the main loop runs instead of once per element (two global loads + a reciprocal each)
  __shared__ __align__(16) float s_bias[256], s_alpha[256], s_inv[256];
  const ConvArgs& a = p.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int a_bytes = 128 * 64 * 2, b_bytes = BN * 64 * 2;
  const int stage_bytes = a_bytes + ((b_bytes + 1023) / 1024) * 1024;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_cv) + 1023) & ~(uintptr_t)1023);
  const int j0 = blockIdx.x * 128, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int kchunks = a.in_ld / 64;
  const int nk = a.taps * kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&hfull[0], 1); mbar_init(&hfull[1], 1); mbar_init(&hempty[0], 1); mbar_init(&hempty[1], 1);
    mbar_fence_init();
  }
  // halo mode: two halo tiles, then the ring of weight tiles
  const int halo_bytes = ((p.halo_rows * 128 + 1023) / 1024) * 1024, w_bytes = ((b_bytes + 1023) / 1024) * 1024;
  unsigned char* wring = base + 2 * (size_t)halo_bytes;
  uint32_t ncols = 32;
  while ((int)ncols < BN) ncols <<= 1;
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 4 && p.halo_rows > 0) {
    if (lane == 0) {
      int wst = 0;
      for (int c = 0; c < kchunks; ++c) {
        const int hb = c & 1;
        if (c >= 2) mbar_wait(&hempty[hb], ((c >> 1) - 1) & 1);
        mbar_expect_tx(&hfull[hb], (uint32_t)(p.halo_rows * 128));      // rows outside [0, Lin) are zero-filled and still counted
        tma_load_3d(base + (size_t)hb * halo_bytes, &p.map_in, c * 64, j0 - a.pad, b, &hfull[hb]);
        for (int tap = 0; tap < a.taps; ++tap, ++wst) {
          const int s = wst % p.stages;
          if (wst >= p.stages) mbar_wait(&empty_bar[s], ((wst / p.stages) - 1) & 1);
          mbar_expect_tx(&full_bar[s], (uint32_t)b_bytes);
          tma_load_2d(wring + (size_t)s * w_bytes, &p.map_w, c * 64, tap * a.N + n0, &full_bar[s]);
        }
      }
    }
  } else if (p.halo_rows > 0 && warp_id_uniform() == 5) {
    const uint32_t idesc = make_idesc(128, BN);
    int wst = 0;
    for (int c = 0; c < kchunks; ++c) {
      const int hb = c & 1;
      mbar_wait(&hfull[hb], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t ha = smem_u32(base + (size_t)hb * halo_bytes);
      for (int tap = 0; tap < a.taps; ++tap, ++wst) {
        const int s = wst % p.stages;
        mbar_wait(&full_bar[s], (wst / p.stages) & 1);
        tc_fence_after();
        const uint64_t da = make_smem_desc(ha + (uint32_t)(tap * a.dil) * 128u);     // the tap's rows: tile start moved down tap * dil rows
        const uint64_t db = make_smem_desc(smem_u32(wring + (size_t)s * w_bytes));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (c | tap | kk) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
          if (tap == a.taps - 1) tc_commit(&hempty[hb]);     // every tap of the chunk has read the halo tile
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(&tmem_full_bar);
    __syncwarp();
  } else if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % p.stages;
        if (kb >= p.stages) mbar_wait(&empty_bar[s], ((kb / p.stages) - 1) & 1);
        const int tap = kb / kchunks, c0 = (kb % kchunks) * 64;
        const int shift = a.ups ? -tap : tap * a.dil - a.pad;
        unsigned char* sa = base + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + b_bytes));
        tma_load_3d(sa, &p.map_in, c0, j0 + shift, b, &full_bar[s]);
        tma_load_2d(sa + a_bytes, &p.map_w, c0, tap * a.N + n0, &full_bar[s]);
      }
    }
  } else if (warp_id_uniform() == 5) {
    // ===== MMA issuer: the whole warp walks the ring on uniform values, one elected lane issues (see elect_one) =====
    const uint32_t idesc = make_idesc(128, BN);
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % p.stages;
      mbar_wait(&full_bar[s], (kb / p.stages) & 1);
      tc_fence_after();
      const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
      const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + a_bytes);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)                    // UMMA K = 16 bf16 = 32 bytes: advance the start address
          tc_mma(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb | kk) ? 1u : 0u);
        tc_commit(&empty_bar[s]);                        // frees the smem slot when these MMAs retire
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&tmem_full_bar);          // accumulator complete
    __syncwarp();
  } else {
    // ===== epilogue: lane <-> GEMM row j (time), 16 output columns per TMEM read =====
    const int j = j0 + warp * 32 + lane;
    {
      for (int c = threadIdx.x; c < BN; c += 128) {
        const int n = n0 + c, co = a.ups ? n % a.Cout : n;        // transposed conv: GEMM column = phase * Cout + channel
        const bool ok = n < a.N;
        const float al = (a.alpha && ok) ? a.alpha[co] : 1.f;
        s_bias[c] = ok ? a.bias[co] : 0.f; s_alpha[c] = al; s_inv[c] = __frcp_rn(al + 1e-9f);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");          // the four epilogue warps
    }
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      const int nb = n0 + c0;
      if (j >= a.rows || nb >= a.N) continue;
      // the 16 columns share one output row when they do not straddle a phase boundary (always true for plain convs)
      int time = j, co0 = nb;
      if (a.ups) { const int ph = nb / a.Cout; co0 = nb % a.Cout; time = a.ups * j + ph - a.pad; }
      const bool same_row = !a.ups || (co0 + 16 <= a.Cout);
      if (same_row && co0 + 16 <= a.Cout) {
        if (time < 0 || time >= a.Lout) continue;
        const size_t o = ((size_t)b * a.Lout + time) * a.out_ld + co0;
        float r[16];
        if (a.resid) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(a.resid + o), r1 = *reinterpret_cast<const uint4*>(a.resid + o + 8);
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) { r[2 * q] = bf16lo(rw[q]); r[2 * q + 1] = bf16hi(rw[q]); }
        }
        uint32_t raw[8], act[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int cl = c0 + 2 * q;                              // column inside the tile: constants from shared memory (broadcast reads)
          float x0 = rbf(v[2 * q] + s_bias[cl]), x1 = rbf(v[2 * q + 1] + s_bias[cl + 1]);
          if (a.resid) { x0 = rbf(r[2 * q] + x0); x1 = rbf(r[2 * q + 1] + x1); }
          raw[q] = pack_bf16(x0, x1);
          act[q] = a.alpha ? pack_bf16(snake_inv(x0, s_alpha[cl], s_inv[cl]), snake_inv(x1, s_alpha[cl + 1], s_inv[cl + 1])) : raw[q];
        }
        if (a.out_raw) {
          *reinterpret_cast<uint4*>(a.out_raw + o) = make_uint4(raw[0], raw[1], raw[2], raw[3]);
          *reinterpret_cast<uint4*>(a.out_raw + o + 8) = make_uint4(raw[4], raw[5], raw[6], raw[7]);
        }
        if (a.out_act) {
          *reinterpret_cast<uint4*>(a.out_act + o) = make_uint4(act[0], act[1], act[2], act[3]);
          *reinterpret_cast<uint4*>(a.out_act + o + 8) = make_uint4(act[4], act[5], act[6], act[7]);
        }
        // zero the padding channels [Cout, out_ld) once per output row (the thread that owns the last real chunk)
        if (co0 + 16 == a.Cout && a.out_ld > a.Cout) {
          for (int cz = a.Cout; cz < a.out_ld; cz += 8) {
            const size_t oz = ((size_t)b * a.Lout + time) * a.out_ld + cz;
            if (a.out_raw) *reinterpret_cast<uint4*>(a.out_raw + oz) = make_uint4(0, 0, 0, 0);
            if (a.out_act) *reinterpret_cast<uint4*>(a.out_act + oz) = make_uint4(0, 0, 0, 0);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int n = nb + q;
          if (n >= a.N) continue;
          int t2 = j, co = n;
          if (a.ups) { const int ph = n / a.Cout; co = n % a.Cout; t2 = a.ups * j + ph - a.pad; }
          if (t2 < 0 || t2 >= a.Lout) continue;
          const size_t o = ((size_t)b * a.Lout + t2) * a.out_ld + co;
          float x0 = rbf(v[q] + s_bias[c0 + q]);
          if (a.resid) x0 = rbf(bf2f(a.resid[o]) + x0);
          if (a.out_raw) a.out_raw[o] = f2bf(x0);
          if (a.out_act) a.out_act[o] = f2bf(a.alpha ? snake_inv(x0, s_alpha[c0 + q], s_inv[c0 + q]) : x0);
          if (co == a.Cout - 1 && a.out_ld > a.Cout) {
            for (int cz = a.Cout; cz < a.out_ld; ++cz) {
              const size_t oz = ((size_t)b * a.Lout + t2) * a.out_ld + cz;
              if (a.out_raw) a.out_raw[oz] = f2bf(0.f);
              if (a.out_act) a.out_act[oz] = f2bf(0.f);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
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
__global__ void relayout_conv_kernel(const float* w, int Cout, int Cin, int Cp, int K, bf16* out) {
  const size_t total = (size_t)K * Cout * Cp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % Cp; const int co = (i / Cp) % Cout; const int k = i / ((size_t)Cp * Cout);
    out[i] = f2bf(ci < Cin ? w[((size_t)co * Cin + ci) * K + k] : 0.f);
  }
}
// w_out[tap][ph*Cout+co][ci] from ConvTranspose1d weight [Cin][Cout][2s]; tap 0 <-> k = ph, tap 1 <-> k = ph + s
__global__ void relayout_convT_kernel(const float* w, int Cin, int Cp, int Cout, int s, bf16* out) {
  const size_t total = (size_t)2 * s * Cout * Cp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % Cp; const size_t r = i / Cp; const int n = r % ((size_t)s * Cout); const int tap = r / ((size_t)s * Cout);
    const int ph = n / Cout, co = n % Cout;
    out[i] = f2bf(ci < Cin ? w[((size_t)ci * Cout + co) * (2 * s) + ph + tap * s] : 0.f);
  }
}

// final Conv1d(C -> 1, k7, pad 3) + tanh  (one warp per output sample)
__global__ void __launch_bounds__(256) final_conv_kernel(const bf16* in, const bf16* w /*[7][ld]*/, const float* bias, int B, int L,
                                                         int C, int ld, float* wav) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int b = warp / L, t = warp % L;
  float s = 0.f;
  for (int k = 0; k < 7; ++k) {
    const int tt = t + k - 3;
    if (tt < 0 || tt >= L) continue;
    const bf16* row = in + ((size_t)b * L + tt) * ld;
    for (int c = lane; c < C; c += 32) s = fmaf(bf2f(row[c]), bf2f(w[k * ld + c]), s);
  }
  s = warp_sum(s);
  if (lane == 0) wav[(size_t)b * L + t] = tanhf(rbf(s + bias[0]));   // conv out bf16 (autocast); tanh kept in fp32
}

}  // namespace

struct zb_conv_w { bf16* w = nullptr; float* bias = nullptr; int Cin = 0, Cinp = 0, Cout = 0, taps = 0, dil = 1, pad = 0, ups = 0; };
static inline int pad64(int c) { return (c + 63) / 64 * 64; }
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
int env_simt() { const char* v = getenv("ZB_DAC_SIMT"); return v && atoi(v); }

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
  o->Cin = Cin; o->Cinp = pad64(Cin); o->Cout = Cout; o->taps = K; o->dil = dil; o->pad = (K - 1) * dil / 2; o->ups = 0;
  if (zb_status st = dev_alloc(d, (void**)&o->w, (size_t)K * Cout * o->Cinp * 2)) return st;
  relayout_conv_kernel<<<256, 256, 0, s>>>(w, Cout, Cin, o->Cinp, K, o->w);
  ZB_CHECK_LAUNCH(d->ctx);
  return copy_f32(d, b, Cout, &o->bias, s);
}
zb_status make_convT(zb_dac* d, const float* w, const float* b, int Cin, int Cout, int stride, zb_conv_w* o, cudaStream_t s) {
  o->Cin = Cin; o->Cinp = pad64(Cin); o->Cout = Cout; o->taps = 2; o->dil = 1; o->pad = (stride + 1) / 2; o->ups = stride;
  if (zb_status st = dev_alloc(d, (void**)&o->w, (size_t)2 * stride * Cout * o->Cinp * 2)) return st;
  relayout_convT_kernel<<<256, 256, 0, s>>>(w, Cin, o->Cinp, Cout, stride, o->w);
  ZB_CHECK_LAUNCH(d->ctx);
  return copy_f32(d, b, Cout, &o->bias, s);
}

zb_status make_map(zb_ctx* ctx, CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  ZB_REQUIRE(ctx, enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ZB_REQUIRE(ctx, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return ZB_OK;
}

zb_status run_conv(zb_ctx* ctx, const zb_conv_w& c, const bf16* in, int B, int Lin, const bf16* resid, bf16* out_raw, bf16* out_act,
                   const float* alpha, cudaStream_t s) {
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.B = B; a.Lin = Lin; a.Cin = c.Cin; a.Cout = c.Cout; a.taps = c.taps; a.dil = c.dil;
  a.pad = c.pad; a.ups = c.ups; a.in_ld = c.Cinp; a.out_ld = pad64(c.Cout);
  if (c.ups) { a.N = c.ups * c.Cout; a.rows = Lin + 1; a.Lout = Lin * c.ups; }
  else { a.N = c.Cout; a.rows = Lin; a.Lout = Lin; }
  a.resid = resid; a.out_raw = out_raw; a.out_act = out_act; a.alpha = alpha;
  static const int simt = env_simt();
  if (simt) {
    dim3 grid((a.rows + BM - 1) / BM, (a.N + BN - 1) / BN, B);
    conv_gemm_kernel<<<grid, 256, 0, s>>>(a);
    ZB_CHECK_LAUNCH(ctx);
    return ZB_OK;
  }
  ConvTcArgs p;
  memset(&p, 0, sizeof(p));
  p.c = a;
  // output columns per CTA: a multiple of 16, at most 256, balanced over the tiles
  const int ntiles = (a.N + 255) / 256;
  int bn = ((a.N + ntiles - 1) / ntiles + 15) / 16 * 16;
  p.BN = bn;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.in_ld, (cuuint64_t)Lin, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)a.in_ld * 2, (cuuint64_t)Lin * a.in_ld * 2};
    // 64 x 861 frames: 225.8 -> 209.9 ms (no gain while the epilogue was the bound: 345.5 -> 344.7)
    static const int env_halo = getenv("ZB_DAC_HALO") ? atoi(getenv("ZB_DAC_HALO")) : 1;
    const int halo_rows = 128 + (c.taps - 1) * c.dil;
    p.halo_rows = (env_halo && !c.ups && c.taps > 1 && halo_rows <= 256) ? halo_rows : 0;
    cuuint32_t box[3] = {64, (cuuint32_t)(p.halo_rows ? p.halo_rows : 128), 1};
    if (zb_status st = make_map(ctx, &p.map_in, in, 3, dims, str, box)) return st;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.in_ld, (cuuint64_t)c.taps * a.N};
    cuuint64_t str[1] = {(cuuint64_t)a.in_ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    if (zb_status st = make_map(ctx, &p.map_w, c.w, 2, dims, str, box)) return st;
  }
  // Several CTAs per SM, so that the epilogue of one tile (Snake + stores on 128 threads) overlaps the TMA / MMA main
  // loops of the others: tensor memory allows 512 / columns CTAs, shared memory is split between them.
  const int a_bytes = 128 * 64 * 2, b_bytes = ((bn * 64 * 2 + 1023) / 1024) * 1024;
  static const int env_ctas = getenv("ZB_DAC_CTAS") ? atoi(getenv("ZB_DAC_CTAS")) : 0;
  int ctas = bn <= 128 ? 3 : 2;
  if (env_ctas > 0) ctas = std::min(env_ctas, bn <= 128 ? 4 : 2);
  const int nk_total = c.taps * (a.in_ld / 64);
  int stages = ((220 * 1024) / ctas - 1024) / (a_bytes + b_bytes);
  size_t halo_smem = 0;
  if (p.halo_rows) {                                          // two halo tiles + a ring of weight tiles only
    halo_smem = 2 * (size_t)(((p.halo_rows * 128 + 1023) / 1024) * 1024);
    stages = (int)(((220 * 1024) / ctas - 1024 - (long)halo_smem) / b_bytes);
  }
  if (stages > 8) stages = 8;
  if (stages > nk_total) stages = nk_total;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = p.halo_rows ? halo_smem + (size_t)stages * b_bytes + 1024 : (size_t)stages * (a_bytes + b_bytes) + 1024;
  ZB_CUDA(ctx, zb_ensure_smem(ctx, conv_tc_kernel, smem));
  dim3 grid((a.rows + 127) / 128, ntiles, B);
  conv_tc_kernel<<<grid, CV_THREADS, smem, s>>>(p);
  ZB_CHECK_LAUNCH(ctx);
  return ZB_OK;
}
}  // namespace

extern "C" zb_status zb_dac_create(zb_ctx* ctx, const zb_dac_desc* desc, zb_dac** out, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
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
  zb_device_guard dev_guard(dac->ctx);
  for (void* p : dac->owned) cudaFree(p);
  delete dac;
  return ZB_OK;
}

extern "C" zb_status zb_dac_decode(zb_ctx* ctx, const zb_dac* dac, const int64_t* codes, int32_t B, int32_t T, float* wav,
                                   zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, dac && codes && wav && B >= 1 && T >= 1, "zb_dac_decode: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const zb_dac_desc& d = dac->d;
  // activation buffers: the widest layer in elements is max over stages of L*C; with strides (8,8,4,2) and channel
  // halving that is the last stage: L = 512 T, C = 96  (and the first: T x 1536 is far smaller)
  int up = 1;
  size_t max_elems = (size_t)T * d.channels;
  { int ch = d.channels; for (int i = 0; i < d.n_blocks; ++i) { up *= d.strides[i]; ch /= 2; max_elems = std::max(max_elems, (size_t)T * up * pad64(ch)); } }
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
  final_conv_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(act, dac->conv2.w, dac->conv2.bias, B, L, C, dac->conv2.Cinp, wav);
  ZB_CHECK_LAUNCH(ctx);
  return ZB_OK;
}
