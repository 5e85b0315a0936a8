// Fused sampler + EOS state machine + delayed-frame write + loop counters.
//
// Replaces (reference paths): zonos/model.py:476 (logit bias add), zonos/sampling.py:130-231
// (repetition penalty, softmax, unified, top-p, top-k, min-p, exponential-race multinomial),
// zonos/model.py:483-497 + zonos/utilities/tensor_ops.py:155-211 (EOS state machine),
// tensor_ops.py:12-53 (frame write) and :56-105 (counters + early-exit test).
//
// One CTA per utterance, one warp per codebook row (Q <= 16, V <= 1056): a row of 1025 fp32 logits
// is 33 registers per lane, every reduction is a warp shuffle, and the whole chain for all 9 codebooks of
// an utterance is ONE launch instead of the reference's 50-70 aten launches.  fp32 op order follows the
// reference (each stage renormalises with a divide) so that, given the same Exp(1) draws q, the chosen
// token is the reference's.
#include "internal.h"

#define SAMP_NPER 33          // ceil(1056 / 32)
#define SAMP_MAXV (SAMP_NPER * 32)
#define SAMP_SORTN 2048

namespace {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
// Philox4x32-10: the four output words for counter (a,b,c,d)
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(a, b, c, d, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(a, b, c, d);
}
// p / s for p >= 0, s > 0 with the IEEE result of the plain division: +0 / s is +0, and skipping it matters - a zero
// operand sends the fp32 division down its slow path (~100 instructions), and after softmax / min-p most of a row IS zero
__device__ __forceinline__ float div_pos(float p, float s) { return p == 0.0f ? 0.0f : p / s; }
__device__ __forceinline__ float exp1_from_bits(uint32_t x) {
  float u = ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  return -__logf(u);
}
// Exp(1) draws for the columns col = j*32 + lane, j = 4*jb .. 4*jb+3, of row `row` in sample call `draw`: one Philox
// call (counter (draw, row, jb*32 + lane)) serves four columns of the lane
__device__ __forceinline__ uint4 exp1_block(uint64_t seed, uint64_t draw, uint32_t row, uint32_t jb, uint32_t lane) {
  return philox4(seed, (uint32_t)draw, (uint32_t)(draw >> 32), row, jb * 32u + lane);
}

struct SampleArgs {
  const float* logits;        // [B,Q,V]
  int B, Q, V;
  const int64_t* window;      // standalone mode: [B,Q,W]
  int64_t wsb, wsq;
  int W;
  const float* q;             // [B,Q,V] or null
  uint64_t seed, draw_index;
  zb_sampling sp;
  int apply_bias;
  int64_t* tokens;            // [B,Q] or null
  // loop mode
  zb_loop_state* st;          // null in standalone mode
  int64_t* delayed;           // [B,Q,T]
  int T;
  int ctx_len;                // min(max_new_tokens, 100)  (model.py:465)
  int32_t* lengths;           // [2B]
  const float* q_stream;      // [q_calls,B,Q,V]
  int q_calls;
  float* logits_trace;        // [trace_calls,B,Q,V]
  int trace_calls;
  int first;                  // 1: the post-prefill sample (model.py:423-431)
  int prefix_len;             // Lc + P + 1 (first only)
  int32_t* mirror;            // host-mapped progress words (offset, step_idx, done, steps)
  unsigned* reset_word;       // grid-barrier counter of the persistent decode kernel: zeroed after every step
  unsigned long long* steplog; // debug: [8200 + 2*step] start (after the dependency wait), +1 end (zb_debug_steplog)
};

// EOS state machine + frame write (lane k <-> codebook k), then counters (lane 0): run by ONE warp per utterance once all
// of its Q tokens are in s_tok
__device__ __forceinline__ void sample_close_row(const SampleArgs& a, zb_loop_state* st, int b, const volatile long long* s_tok, int offset_new,
                                                 int lane, int log_step) {
  const int V = a.V, Q = a.Q;
#define SLOG(k) do { if (log_step >= 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); a.steplog[16400 + 8 * log_step + (k)] = t_; } } while (0)
  const int eos = V - 1, mask_tok = V;           // 1024 / 1025
  long long rem = st->remaining[b];
  int stop = st->stopping[b];
  if (!a.first && s_tok[0] == eos) {             // model.py:483-488
    rem = min(rem, (long long)Q);
    stop = 1;
  }
  if (lane < Q) {
    long long tok = s_tok[lane];
    if (!a.first && stop) {                      // tensor_ops.py:193-211
      const long long eos_idx = min((long long)Q - rem, (long long)(Q - 1));   // model.py:490-491
      if (lane < eos_idx) tok = mask_tok;
      else if (lane == eos_idx) tok = eos;
    }
    int64_t* cell = a.delayed + ((size_t)b * Q + lane) * a.T + offset_new;    // tensor_ops.py:42-53 (only where == -1)
    if (*cell == -1) *cell = tok;
  }
  if (lane == 0) {
    const int adv = a.first ? a.prefix_len : 1;  // model.py:430-431 / tensor_ops.py:85-86
    a.lengths[b] += adv;
    a.lengths[a.B + b] += adv;
    if (!a.first) rem -= 1;                       // tensor_ops.py:87
    st->remaining[b] = rem;
    st->stopping[b] = stop;
    SLOG(5);
    __threadfence();
    const int arrived = atomicAdd(&st->arrive, 1);
    SLOG(6);
    if (arrived == a.B - 1) {                     // last utterance of this step closes the step
      st->arrive = 0;
      st->draw_idx += 1;
      if (!a.first) {
        const int step_idx = st->step_idx;
        int done = 0, steps = step_idx + 1;       // model.py:506
        bool check = (step_idx % 16 == 15);       // tensor_ops.py:90-103
        if (!check && (step_idx % 8 == 7)) {
          int est = a.B * 10 - (step_idx + 1);    // cpu_step_counter == step_idx + 1
          if (est < 0) est = 0;
          check = est < 5;
        }
        if (check) {
          bool all_done = true;
          for (int bb = 0; bb < a.B; ++bb) all_done = all_done && (((volatile long long*)st->remaining)[bb] <= 0);
          if (all_done) { done = 1; steps = step_idx; }   // `break` precedes `step = step_idx + 1`
        }
        st->offset = offset_new; st->steps = steps; st->step_idx = step_idx + 1; st->done = done;
        if (a.steplog) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.steplog[8200 + 2 * min(step_idx, 4000) + 1] = t; }
        // host-visible progress (host-mapped memory, PCIe write): only on the steps the host may look at
        if (a.mirror && (check || done)) { a.mirror[0] = offset_new; a.mirror[1] = step_idx + 1; a.mirror[3] = steps; a.mirror[2] = done; }
      }
    }
  }
}
#undef SLOG

// kThreads = 32 * (most codebooks the launch may have): 288 for the 9 codebooks of Zonos leaves 224 registers per thread
// (the row lives in registers: 33 values per lane plus temporaries - 128 registers spill)
template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1) sample_kernel(SampleArgs a) {
  extern __shared__ unsigned long long sort_keys[];   // [Q][SAMP_SORTN], only when top_p/top_k
  __shared__ long long s_tok[16];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = a.V, Q = a.Q;
  zb_loop_state* st = a.st;
  pdl_launch_dependents();
  pdl_wait();
  if (a.reset_word && blockIdx.x == 0 && threadIdx.x == 0) *a.reset_word = 0;
  const int log_step = (a.steplog && st && blockIdx.x == 0 && threadIdx.x == 0) ? min(st->steps, 4000) : -1;
  if (log_step >= 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.steplog[8200 + 2 * log_step] = t; }
#define SLOG(k) do { if (log_step >= 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); a.steplog[16400 + 8 * log_step + (k)] = t_; } } while (0)

  int offset_new = 0;
  uint64_t draw = a.draw_index;
  if (st) {
    // liveness: model.py:468-472
    if (st->done) return;
    offset_new = a.first ? st->offset : st->offset + 1;
    if (!a.first && offset_new >= a.T) {
      if (b == 0 && threadIdx.x == 0) {
        st->offset = offset_new; st->done = 1;
        if (a.mirror) { a.mirror[0] = offset_new; a.mirror[2] = 1; }
      }
      return;
    }
    draw = (uint64_t)st->draw_idx;
  }

  if (warp < Q) {
    const int qi = warp;
    const float* lrow = a.logits + ((size_t)b * Q + qi) * V;
    float x[SAMP_NPER];
#pragma unroll
    for (int j = 0; j < SAMP_NPER; ++j) {
      int i = j * 32 + lane;
      x[j] = (i < V) ? lrow[i] : -INFINITY;
    }
    // ---- logit bias (model.py:433-437,476): EOS = V-1 ----
    if (a.apply_bias) {
      const int eos = V - 1;
      const int j = eos >> 5;
      if ((eos & 31) == lane) {
#pragma unroll
        for (int jj = 0; jj < SAMP_NPER; ++jj)
          if (jj == j) x[jj] = (qi == 0) ? (x[jj] - 0.69314718055994530942f) : -INFINITY;
      }
    }
    if (st && a.logits_trace && draw < (uint64_t)a.trace_calls) {
      float* t = a.logits_trace + (((size_t)draw * a.B + b) * Q + qi) * V;
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) {
        int i = j * 32 + lane;
        if (i < V) t[i] = x[j];
      }
    }
    // ---- repetition penalty (sampling.py:159-163) ----
    const int64_t* wtok = nullptr;
    int wcount = 0;
    int64_t wstride = 1;
    if (st) {
      if (!a.first) {
        int avail = min(offset_new, a.ctx_len);          // columns in delayed[..., max(0,offset-ctx):offset]
        wcount = min(a.sp.repetition_penalty_window, avail);
        wtok = a.delayed + ((size_t)b * Q + qi) * a.T + (offset_new - wcount);
      }
    } else if (a.window) {
      wcount = min(a.sp.repetition_penalty_window, a.W);
      wtok = a.window + b * a.wsb + qi * a.wsq + (a.W - wcount);
    }
    if (wtok && a.sp.repetition_penalty != 1.0f && wcount > 0) {
      // factors = prod of penalty over occurrences; then where(l<=0, l*f, l/f)
      float f[SAMP_NPER];
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) f[j] = 1.0f;
      for (int w = 0; w < wcount; ++w) {
        long long t = wtok[w * wstride];
        int idx = (int)min(t, (long long)(V - 1));
        if (idx < 0) idx += V;                              // torch index wrap; never hit on the path
        if ((idx & 31) == lane) {
          int j = idx >> 5;
#pragma unroll
          for (int jj = 0; jj < SAMP_NPER; ++jj)
            if (jj == j) f[jj] *= a.sp.repetition_penalty;
        }
      }
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) x[j] = (f[j] == 1.0f) ? x[j] : ((x[j] <= 0.0f) ? x[j] * f[j] : x[j] / f[j]);   // x * 1 == x / 1 == x
    }

    SLOG(0);
    int best = 0;
    if (a.sp.temperature > 0.0f) {
      // ---- softmax(logits / T) (sampling.py:217) ----
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) {
        if (a.sp.temperature != 1.0f) x[j] = x[j] / a.sp.temperature;           // x / 1 == x
        m = fmaxf(m, x[j]);
      }
      m = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) {
        x[j] = expf(x[j] - m);
        s += x[j];
      }
      s = warp_sum(s);
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) x[j] = div_pos(x[j], s);

      SLOG(1);
      // ---- NovelAI unified (sampling.py:60-63) ----
      if (a.sp.linear > 0.0f) {
        float ent = 0.f;
        float lp[SAMP_NPER];
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) {
          lp[j] = logf(fmaxf(x[j], 1e-20f));
          ent += x[j] * lp[j];
        }
        ent = -warp_sum(ent);
        const float lin = __fadd_rn(a.sp.linear, __fmul_rn(ent, a.sp.conf));
        float m2 = -INFINITY;
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) {
          int i = j * 32 + lane;
          float raw = __fsub_rn(__fmul_rn(lp[j], lin), __fmul_rn(__fmul_rn(lp[j], lp[j]), a.sp.quad));
          x[j] = (i < V) ? raw : -INFINITY;
          m2 = fmaxf(m2, x[j]);
        }
        m2 = warp_max(m2);
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) {
          x[j] = expf(x[j] - m2);
          s2 += x[j];
        }
        s2 = warp_sum(s2);
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) x[j] = div_pos(x[j], s2);
      }

      // ---- top-p / top-k need the row sorted (sampling.py:77-80,93-98) ----
      if (a.sp.top_p > 0.0f || a.sp.top_k > 0) {
        unsigned long long* keys = sort_keys + (size_t)qi * SAMP_SORTN;
        // key: descending p, ties by ascending index (== torch's stable descending sort)
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) {
          int i = j * 32 + lane;
          if (i < V) keys[i] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(x[j])) << 32) | (unsigned)i;
        }
        for (int i = V + lane; i < SAMP_SORTN; i += 32) keys[i] = ~0ull;
        __syncwarp();
        for (int k = 2; k <= SAMP_SORTN; k <<= 1) {
          for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = lane; i < SAMP_SORTN; i += 32) {
              int ixj = i ^ jj;
              if (ixj > i) {
                unsigned long long A = keys[i], Bv = keys[ixj];
                bool up = ((i & k) == 0);
                if ((A > Bv) == up) { keys[i] = Bv; keys[ixj] = A; }
              }
            }
            __syncwarp();
          }
        }
        if (a.sp.top_p > 0.0f) {
          // sequential cumsum like torch.cumsum on the CPU; entry i is dropped where (cumsum_i - p_i) > top_p.
          // lane 0 walks the sorted row once, marks dropped entries (bit 31 of the index word) and, for a
          // following top-k, remembers the k-th largest surviving value.
          float pv_sorted = 0.f;          // k-th largest kept p (before renormalisation); 0 if fewer than k kept
          if (lane == 0) {
            float cs = 0.f;
            int kept = 0;
            const int kk = a.sp.top_k > 0 ? min(a.sp.top_k, V) : 0;
            for (int i = 0; i < V; ++i) {
              const float p = __uint_as_float(0xFFFFFFFFu - (unsigned)(keys[i] >> 32));
              cs += p;
              if (cs - p > a.sp.top_p) keys[i] |= 0x80000000ull;
              else if (++kept == kk) pv_sorted = p;
            }
          }
          pv_sorted = __shfl_sync(0xffffffffu, pv_sorted, 0);
          __syncwarp();
          // scatter the keep flags back to vocabulary order through the (now unused) padding tail of the keys
          float* flags = reinterpret_cast<float*>(keys + V);     // 1023 u64 = 2046 floats >= V floats
          for (int i = lane; i < V; i += 32) {
            const unsigned long long kv = keys[i];
            flags[(unsigned)(kv & 0x7FFFFFFFu)] = (kv & 0x80000000ull) ? 0.f : 1.f;
          }
          __syncwarp();
          float s3 = 0.f;
#pragma unroll
          for (int j = 0; j < SAMP_NPER; ++j) {
            int i = j * 32 + lane;
            if (i < V) x[j] = x[j] * flags[i];
            s3 += x[j];
          }
          s3 = warp_sum(s3);
#pragma unroll
          for (int j = 0; j < SAMP_NPER; ++j) x[j] = div_pos(x[j], s3);
          __syncwarp();
          if (a.sp.top_k > 0) {
            // kept entries were all divided by s3 (order preserved), dropped ones are 0: the k-th largest of the
            // new row is pv_sorted / s3 when at least k entries survived, else 0 (nothing below 0 to drop)
            const float pv = pv_sorted / s3;
            float s4 = 0.f;
#pragma unroll
            for (int j = 0; j < SAMP_NPER; ++j) {
              if (x[j] < pv) x[j] = 0.f;
              s4 += x[j];
            }
            s4 = warp_sum(s4);
#pragma unroll
            for (int j = 0; j < SAMP_NPER; ++j) x[j] = div_pos(x[j], s4);
          }
        } else {
          int k = min(a.sp.top_k, V);
          float pv = __uint_as_float(0xFFFFFFFFu - (unsigned)(keys[k - 1] >> 32));
          float s4 = 0.f;
#pragma unroll
          for (int j = 0; j < SAMP_NPER; ++j) {
            if (x[j] < pv) x[j] = 0.f;
            s4 += x[j];
          }
          s4 = warp_sum(s4);
#pragma unroll
          for (int j = 0; j < SAMP_NPER; ++j) x[j] = div_pos(x[j], s4);
        }
      }

      // ---- min-p (sampling.py:123-126) ----
      if (a.sp.min_p > 0.0f) {
        float pm = 0.f;
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) pm = fmaxf(pm, x[j]);
        pm = warp_max(pm);
        const float thr = a.sp.min_p * pm;
        float s5 = 0.f;
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) {
          if (x[j] < thr) x[j] = 0.f;
          s5 += x[j];
        }
        s5 = warp_sum(s5);
#pragma unroll
        for (int j = 0; j < SAMP_NPER; ++j) x[j] = div_pos(x[j], s5);
      }

      SLOG(2);
      // ---- exponential race: argmax(p / q) (sampling.py:28-30) ----
      const float* qrow = nullptr;
      if (st && a.q_stream && draw < (uint64_t)a.q_calls)
        qrow = a.q_stream + (((size_t)draw * a.B + b) * Q + qi) * V;
      else if (!st && a.q)
        qrow = a.q + ((size_t)b * Q + qi) * V;
      float bs = -INFINITY;
      int bi = 0x7fffffff;
      uint4 rnd = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) {
        int i = j * 32 + lane;
        if (!qrow && (j & 3) == 0) rnd = exp1_block(a.seed, draw, (uint32_t)(b * Q + qi), (uint32_t)(j >> 2), (uint32_t)lane);
        if (i < V) {
          const uint32_t bits = (j & 3) == 0 ? rnd.x : ((j & 3) == 1 ? rnd.y : ((j & 3) == 2 ? rnd.z : rnd.w));
          float qq = qrow ? qrow[i] : exp1_from_bits(bits);
          float sc = div_pos(x[j], qq);
          if (sc > bs) { bs = sc; bi = i; }     // ascending i per lane: first max kept
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float os = __shfl_xor_sync(0xffffffffu, bs, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
      }
      best = bi;
    } else {
      // greedy (sampling.py:229): argmax of the penalised logits, lowest index on ties
      float bs = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < SAMP_NPER; ++j) {
        int i = j * 32 + lane;
        if (i < V && (x[j] > bs || bi == 0x7fffffff)) { bs = x[j]; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float os = __shfl_xor_sync(0xffffffffu, bs, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
      }
      best = bi;
    }
    SLOG(3);
    if (lane == 0) {
      s_tok[qi] = best;
      if (a.tokens) a.tokens[(size_t)b * Q + qi] = best;
    }
  }
  if (!st) return;
  __syncthreads();
  SLOG(4);

  // ---- EOS state machine + frame write (lane k <-> codebook k of warp 0), then counters (lane 0) ----
  if (warp != 0) return;
  sample_close_row(a, st, b, s_tok, offset_new, lane, log_step);
}


// ---- wide variant: grid (B, Q), 128 threads per codebook row ------------------------------------------------------
// One warp per row is a serial chain of ~12 k instructions (IEEE divisions, precise exp/log in the reference's op order):
// 21 us per step whether the 9 warps share an SM or not.  Four warps per row and one CTA per row cut the chain to 9
// elements per thread.  Same stages and op order per element; sums are reduced thread -> warp (butterfly) -> the 4
// warps in order.  Used when no sort is needed (top_p = top_k = 0); the Exp(1) draw of an element is the same Philox
// word as in sample_kernel.
#define SAMPW_THREADS 128
#define SAMPW_NPER 9          // ceil(1056 / 128)
__device__ __forceinline__ float blockw_sum(float v, float* red, int warp, int lane) {
  v = warp_sum(v);
  __syncthreads();                                  // red is free again
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return ((red[0] + red[1]) + red[2]) + red[3];
}
__device__ __forceinline__ float blockw_max(float v, float* red, int warp, int lane) {
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}

__global__ void __launch_bounds__(SAMPW_THREADS) sample_wide_kernel(SampleArgs a) {
  __shared__ float red[4];
  __shared__ float red_s[4];
  __shared__ int red_i[4];
  __shared__ int s_last;
  const int b = blockIdx.x, qi = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int V = a.V, Q = a.Q;
  zb_loop_state* st = a.st;
  pdl_launch_dependents();
  pdl_wait();
  const int log_step = (a.steplog && st && b == 0 && qi == 0 && t == 0) ? min(st->steps, 4000) : -1;
  if (log_step >= 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); a.steplog[8200 + 2 * log_step] = t_; }

  int offset_new = 0;
  uint64_t draw = a.draw_index;
  if (st) {
    if (st->done) return;                                                   // liveness: model.py:468-472
    offset_new = a.first ? st->offset : st->offset + 1;
    if (!a.first && offset_new >= a.T) {
      if (b == 0 && qi == 0 && t == 0) {
        st->offset = offset_new; st->done = 1;
        if (a.mirror) { a.mirror[0] = offset_new; a.mirror[2] = 1; }
      }
      return;
    }
    draw = (uint64_t)st->draw_idx;
  }

  const float* lrow = a.logits + ((size_t)b * Q + qi) * V;
  float x[SAMPW_NPER];
#pragma unroll
  for (int j = 0; j < SAMPW_NPER; ++j) {
    const int i = j * SAMPW_THREADS + t;
    x[j] = (i < V) ? lrow[i] : -INFINITY;
  }
  if (a.apply_bias) {                                                       // model.py:433-437,476: EOS = V-1
    const int eos = V - 1;
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j)
      if (j * SAMPW_THREADS + t == eos) x[j] = (qi == 0) ? (x[j] - 0.69314718055994530942f) : -INFINITY;
  }
  if (st && a.logits_trace && draw < (uint64_t)a.trace_calls) {
    float* tr = a.logits_trace + (((size_t)draw * a.B + b) * Q + qi) * V;
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) {
      const int i = j * SAMPW_THREADS + t;
      if (i < V) tr[i] = x[j];
    }
  }
  // ---- repetition penalty (sampling.py:159-163) ----
  const int64_t* wtok = nullptr;
  int wcount = 0;
  if (st) {
    if (!a.first) {
      const int avail = min(offset_new, a.ctx_len);
      wcount = min(a.sp.repetition_penalty_window, avail);
      wtok = a.delayed + ((size_t)b * Q + qi) * a.T + (offset_new - wcount);
    }
  } else if (a.window) {
    wcount = min(a.sp.repetition_penalty_window, a.W);
    wtok = a.window + b * a.wsb + qi * a.wsq + (a.W - wcount);
  }
  if (wtok && a.sp.repetition_penalty != 1.0f && wcount > 0) {
    float f[SAMPW_NPER];
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) f[j] = 1.0f;
    for (int w = 0; w < wcount; ++w) {
      const long long tk = wtok[w];
      int idx = (int)min(tk, (long long)(V - 1));
      if (idx < 0) idx += V;
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j)
        if (j * SAMPW_THREADS + t == idx) f[j] *= a.sp.repetition_penalty;
    }
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) x[j] = (f[j] == 1.0f) ? x[j] : ((x[j] <= 0.0f) ? x[j] * f[j] : x[j] / f[j]);
  }

  int best = 0;
  if (a.sp.temperature > 0.0f) {
    // ---- softmax(logits / T) (sampling.py:217) ----
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) {
      if (a.sp.temperature != 1.0f) x[j] = x[j] / a.sp.temperature;
      m = fmaxf(m, x[j]);
    }
    m = blockw_max(m, red, warp, lane);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) {
      x[j] = expf(x[j] - m);
      sum += x[j];
    }
    sum = blockw_sum(sum, red, warp, lane);
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) x[j] = div_pos(x[j], sum);

    // ---- NovelAI unified (sampling.py:60-63) ----
    if (a.sp.linear > 0.0f) {
      float ent = 0.f;
      float lp[SAMPW_NPER];
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) {
        lp[j] = logf(fmaxf(x[j], 1e-20f));
        ent += x[j] * lp[j];
      }
      ent = -blockw_sum(ent, red, warp, lane);
      const float lin = __fadd_rn(a.sp.linear, __fmul_rn(ent, a.sp.conf));
      float m2 = -INFINITY;
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) {
        const int i = j * SAMPW_THREADS + t;
        const float raw = __fsub_rn(__fmul_rn(lp[j], lin), __fmul_rn(__fmul_rn(lp[j], lp[j]), a.sp.quad));
        x[j] = (i < V) ? raw : -INFINITY;
        m2 = fmaxf(m2, x[j]);
      }
      m2 = blockw_max(m2, red, warp, lane);
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) {
        x[j] = expf(x[j] - m2);
        s2 += x[j];
      }
      s2 = blockw_sum(s2, red, warp, lane);
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) x[j] = div_pos(x[j], s2);
    }

    // ---- min-p (sampling.py:123-126) ----
    if (a.sp.min_p > 0.0f) {
      float pm = 0.f;
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) pm = fmaxf(pm, x[j]);
      pm = blockw_max(pm, red, warp, lane);
      const float thr = a.sp.min_p * pm;
      float s5 = 0.f;
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) {
        if (x[j] < thr) x[j] = 0.f;
        s5 += x[j];
      }
      s5 = blockw_sum(s5, red, warp, lane);
#pragma unroll
      for (int j = 0; j < SAMPW_NPER; ++j) x[j] = div_pos(x[j], s5);
    }

    // ---- exponential race: argmax(p / q) (sampling.py:28-30) ----
    const float* qrow = nullptr;
    if (st && a.q_stream && draw < (uint64_t)a.q_calls)
      qrow = a.q_stream + (((size_t)draw * a.B + b) * Q + qi) * V;
    else if (!st && a.q)
      qrow = a.q + ((size_t)b * Q + qi) * V;
    float bs = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) {
      const int i = j * SAMPW_THREADS + t;
      if (i < V) {
        float qq;
        if (qrow) qq = qrow[i];
        else {
          // element i of sample_kernel is (j32 = i / 32, lane): word j32 % 4 of the Philox block (j32 / 4, lane) = (j, lane), word = warp
          const uint4 rnd = exp1_block(a.seed, draw, (uint32_t)(b * Q + qi), (uint32_t)j, (uint32_t)lane);
          qq = exp1_from_bits(warp == 0 ? rnd.x : (warp == 1 ? rnd.y : (warp == 2 ? rnd.z : rnd.w)));
        }
        const float sc = div_pos(x[j], qq);
        if (sc > bs) { bs = sc; bi = i; }       // ascending i per thread: first max kept
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    if (lane == 0) { red_s[warp] = bs; red_i[warp] = bi; }
    __syncthreads();
    bs = red_s[0]; bi = red_i[0];
#pragma unroll
    for (int w = 1; w < 4; ++w)
      if (red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi)) { bs = red_s[w]; bi = red_i[w]; }
    best = bi;
  } else {
    // greedy (sampling.py:229): argmax of the penalised logits, lowest index on ties
    float bs = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < SAMPW_NPER; ++j) {
      const int i = j * SAMPW_THREADS + t;
      if (i < V && (x[j] > bs || bi == 0x7fffffff)) { bs = x[j]; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    if (lane == 0) { red_s[warp] = bs; red_i[warp] = bi; }
    __syncthreads();
    bs = red_s[0]; bi = red_i[0];
#pragma unroll
    for (int w = 1; w < 4; ++w)
      if (red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi)) { bs = red_s[w]; bi = red_i[w]; }
    best = bi;
  }
  if (t == 0 && a.tokens) a.tokens[(size_t)b * Q + qi] = best;
  if (!st) return;
  // publish the token; the last row of the utterance to arrive carries on with the utterance's bookkeeping
  if (t == 0) {
    st->tok[b * 16 + qi] = best;
    __threadfence();
    const int last = atomicAdd(&st->row_arrive[b], 1) == Q - 1;
    if (last) { st->row_arrive[b] = 0; __threadfence(); }
    s_last = last;
  }
  __syncthreads();
  if (!s_last || warp != 0) return;
  sample_close_row(a, st, b, st->tok + b * 16, offset_new, lane, log_step);
}

}  // namespace

zb_status zb_launch_sample(zb_ctx* ctx, const zb_sample_launch& L, cudaStream_t stream) {
  SampleArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = L.logits; a.B = L.B; a.Q = L.Q; a.V = L.V;
  a.window = L.window; a.wsb = L.wsb; a.wsq = L.wsq; a.W = L.W;
  a.q = L.q; a.seed = L.seed; a.draw_index = L.draw_index; a.sp = L.sp; a.apply_bias = L.apply_bias;
  a.tokens = L.tokens; a.st = L.st; a.delayed = L.delayed; a.T = L.T; a.ctx_len = L.ctx_len;
  a.lengths = L.lengths; a.q_stream = L.q_stream; a.q_calls = L.q_calls; a.logits_trace = L.logits_trace;
  a.trace_calls = L.trace_calls; a.first = L.first; a.prefix_len = L.prefix_len; a.mirror = L.mirror; a.reset_word = L.reset_word; a.steplog = zb_debug_steplog_ptr();
  ZB_REQUIRE(ctx, L.Q >= 1 && L.Q <= 16, "sampler: Q=%d unsupported (1..16)", L.Q);
  ZB_REQUIRE(ctx, L.V >= 2 && L.V <= SAMP_MAXV, "sampler: V=%d unsupported (<= %d)", L.V, SAMP_MAXV);
  ZB_REQUIRE(ctx, L.B >= 1, "sampler: B=%d", L.B);
  size_t smem = 0;
  if (L.sp.temperature > 0.f && (L.sp.top_p > 0.f || L.sp.top_k > 0)) smem = (size_t)L.Q * SAMP_SORTN * 8;
  ZB_REQUIRE(ctx, smem <= 227 * 1024, "sampler: top-p/top-k with Q=%d codebooks needs %zu bytes of shared memory", L.Q, smem);
  const int variant = L.Q <= 9 ? 0 : 1;
  auto kernel = variant == 0 ? sample_kernel<288> : sample_kernel<512>;
  if (smem > 48 * 1024) ZB_CUDA(ctx, zb_ensure_smem(ctx, kernel, smem));
  static const int wide = getenv("ZB_SAMPLER_WIDE") ? atoi(getenv("ZB_SAMPLER_WIDE")) : 1;
  if (wide && smem == 0 && L.V <= SAMPW_NPER * SAMPW_THREADS && L.B <= ZB_MAX_B) {     // no sort needed: one CTA of 4 warps per codebook row
    ZB_CUDA(ctx, zb_launch_pdl(sample_wide_kernel, dim3(L.B, L.Q), dim3(SAMPW_THREADS), 0, stream, a));
    ctx->launches++;
    return ZB_OK;
  }
  ZB_CUDA(ctx, zb_launch_pdl(kernel, dim3(L.B), dim3(32 * L.Q), smem, stream, a));
  ctx->launches++;
  return ZB_OK;
}
