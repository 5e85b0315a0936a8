// Internal launcher interfaces between the translation units of libzonos_b200.so.
#pragma once
#include "common.cuh"

// Device-resident loop state of one generate() session (see sampler.cu, zonos/model.py:439-509).
#define ZB_MAX_B 256
struct zb_loop_state {
  int32_t offset;      // model.py `offset`: column of `delayed` written last
  int32_t step_idx;    // loop iteration about to run
  int32_t done;        // loop finished
  int32_t steps;       // model.py `step`
  int32_t draw_idx;    // sample calls made so far (index into q_stream / logits_trace)
  int32_t arrive;      // last-CTA counter
  int32_t max_steps;
  int32_t _pad;
  long long remaining[ZB_MAX_B];
  int32_t stopping[ZB_MAX_B];
  int32_t row_arrive[ZB_MAX_B];        // wide sampler: codebook rows of an utterance that have their token (one CTA per row)
  long long tok[ZB_MAX_B * 16];        // wide sampler: the tokens of the current step, [utterance][codebook]
};

struct zb_model {
  zb_ctx* ctx;
  zb_model_desc d;
  std::vector<zb_layer> layers;
  std::vector<const void*> emb;
  int n_attn = 0, n_mamba = 0;
  std::vector<int> attn_index;    // layer -> index among attention layers (or -1)
  std::vector<int> mamba_index;
  // Weight copy in the tile order of the persistent kernel's tcgen05 consumer (decode.cu, MegaTcGeo): built by the first
  // generate session that takes that path, rebuilt after zb_model_weights_changed(); owned by the model.
  mutable void* tcw = nullptr; mutable size_t tcw_bytes = 0; mutable int tcw_grid = 0; mutable bool tcw_valid = false;
  // FP8 (e4m3, one power-of-two scale per weight row) copy of the decode matrices for the persistent kernel's opt-in FP8 mode
  // (decode.cu, ZB_FP8=1): [layer][in_proj | out_proj | fc1 | fc2 bytes] [heads bytes] [row scales, fp32].
  // Built by the first generate session that asks for it, rebuilt after zb_model_weights_changed(); owned by the model.
  mutable void* f8w = nullptr; mutable size_t f8w_bytes = 0; mutable bool f8w_valid = false;
  ~zb_model() {
    if (tcw || f8w) cudaSetDevice(ctx->device);
    if (tcw) cudaFree(tcw);
    if (f8w) cudaFree(f8w);
  }
};

struct zb_sample_launch {
  const float* logits = nullptr; int B = 0, Q = 0, V = 0;
  const int64_t* window = nullptr; int64_t wsb = 0, wsq = 0; int W = 0;
  const float* q = nullptr; uint64_t seed = 0, draw_index = 0;
  zb_sampling sp{}; int apply_bias = 0; int64_t* tokens = nullptr;
  zb_loop_state* st = nullptr; int64_t* delayed = nullptr; int T = 0; int ctx_len = 0;
  int32_t* lengths = nullptr; const float* q_stream = nullptr; int q_calls = 0;
  float* logits_trace = nullptr; int trace_calls = 0; int first = 0; int prefix_len = 0;
  int32_t* mirror = nullptr;   // host-mapped copy of the first 8 words of zb_loop_state (device pointer)
  unsigned* reset_word = nullptr;   // zeroed by the kernel (grid-barrier counter of the persistent decode step)
};
zb_status zb_launch_sample(zb_ctx* ctx, const zb_sample_launch& L, cudaStream_t stream);

// ---- backbone pieces (decode.cu) ----
struct zb_embed_launch {
  const zb_model* model; const int64_t* codes; int64_t sb, sq, st; int B, T, repeat; bf16* out;
  int64_t out_rs;             // elements between consecutive output rows (>= T*D)
  const zb_loop_state* loop = nullptr; int T_delayed = 0;   // loop mode: column = loop->offset
};
zb_status zb_launch_embed(zb_ctx* ctx, const zb_embed_launch& L, cudaStream_t stream);

// Runs all layers on the residual stream x[M=R*T, D] in place (scratch-resident), appending K/V.
// loop != null: every kernel exits early once the loop is done (device-driven generate).
// max_kv_len: host upper bound of lengths[r] + T (sizes the split-KV grid).
zb_status zb_run_layers(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, bf16* x, int R, int T, int max_kv_len,
                        const zb_loop_state* loop, int T_delayed, cudaStream_t stream);
// y[R or R*T rows] = final norm(x)
zb_status zb_launch_final_norm(zb_ctx* ctx, const zb_model* model, const bf16* x, int R, int T, int last_only, bf16* y,
                               cudaStream_t stream);
// logits = heads(norm_f?(hidden)) with the CFG mix.  hidden row r at hidden + r*row_stride.
zb_status zb_launch_heads(zb_ctx* ctx, const zb_model* model, const bf16* hidden, int64_t row_stride, int R,
                          int apply_norm, float cfg_scale, float* logits, const zb_loop_state* loop, int T_delayed,
                          cudaStream_t stream);
size_t zb_backbone_scratch_bytes(const zb_model* model, int R, int T, int max_kv_len);
// persistent single-launch decode step (decode.cu)
bool zb_mega_supported(const zb_model* model, int R);
unsigned long long* zb_debug_steplog_ptr();                  // debug (zb_debug_steplog), nullptr normally
size_t zb_mega_layers_bytes(const zb_model* model);
size_t zb_mega_arena_bytes(const zb_model* model, int R);     // tagged activation words of one generate session (zeroed by the caller)
zb_status zb_mega_layers_build(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, void* host_buf, cudaStream_t stream);
zb_status zb_launch_decode_step(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* mega_layers_dev, unsigned* sync,
                                uint32_t* arena, int R, int max_kv_len, float cfg_scale, float* logits, const int64_t* delayed, int T_delayed,
                                const zb_loop_state* loop, cudaStream_t stream);

// persistent decode step with a tcgen05 consumer, R = 2..128 rows (decode_tc.cu)
bool zb_tc_supported(const zb_model* model, int R);
size_t zb_tc_table_bytes(const zb_model* model);              // per-layer tensor maps + pointers (host-built, uploaded once per session)
size_t zb_tc_arena_bytes(const zb_model* model, int R);       // activation / partial buffers of one generate session
zb_status zb_tc_table_build(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, int R, void* arena_dev, void* host_buf);
zb_status zb_launch_decode_tc(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* table_dev, unsigned* bar, void* arena, int R,
                              float cfg_scale, float* logits, const int64_t* delayed, int T_delayed, const zb_loop_state* loop, cudaStream_t stream);

// causal attention for T > 1 tokens per row on mma.sync tiles (attn_prefill.cu)
bool zb_attn_prefill_supported(const zb_model_desc& d);
zb_status zb_launch_attn_prefill(zb_ctx* ctx, const zb_model_desc& d, const zb_cache* cache, const bf16* q, const bf16* kv_layer, bf16* y, int R, int T,
                                 cudaStream_t stream);

// ---- tcgen05 GEMM (gemm_tc.cu): Y[M,N] = X[M,K] W[N,K]^T with fused epilogue ----
struct zb_gemm_tc {
  const bf16* W = nullptr; const bf16* x = nullptr; long long ldx = 0; int M = 0, N = 0, K = 0;
  int epi = 0;                 // 0 store, 1 +residual, 2 QKV (RoPE + KV append), 3 SiLU gate, 4 heads (fp32 logits, CFG mix)
  int F = 0;
  bf16* y = nullptr; long long ldy = 0; const bf16* resid = nullptr; long long ldr = 0;
  int T = 1, Hq = 0, Hkv = 0, hd = 0, rope_interleaved = 1, rope_len = 0, max_pages = 0;
  const float* rope = nullptr; const int32_t* lengths = nullptr; const int32_t* page_table = nullptr; bf16* kv_layer = nullptr; bf16* q_out = nullptr;
  int B = 0; float cfg_scale = 1.0f; float* logits = nullptr; int QV = 0;
  bool decode = false;         // one token per row: split-K allowed (never in prefill, whose results must not depend on the batch size)
};
zb_status zb_launch_gemm_tc(zb_ctx* ctx, const zb_gemm_tc& g, cudaStream_t stream);

// ---- DAC (dac.cu) ----
