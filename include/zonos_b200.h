/* zonos_b200.h - C ABI of libzonos_b200.so: the B200 (sm_100a) implementation of the
 * Zonos-v0.1 inference hot path (reference: langfod/Zonos).
 *
 * Conventions:
 *   - extern "C", opaque handles, plain pointers and sizes; no torch / C++ types.
 *   - Every data pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *     the library owns only its handles and their scratch memory.
 *   - Every call enqueues work on the cudaStream_t passed as `stream` and returns without
 *     synchronising (zb_generate_poll is the one documented exception).
 *   - Return value: ZB_OK (0) or an error code; the message is in zb_last_error().
 *     Nothing throws or aborts.  One ctx per device; calls on one ctx are serialised by the caller.
 *   - bf16 = CUDA __nv_bfloat16 bits; row-major, innermost dimension contiguous.
 *   - Rows: R = 2*B; rows [0,B) are the conditional half, rows [B,2B) the unconditional half
 *     of classifier-free guidance (zonos/utilities/conditioning_cache.py:176, zonos/model.py:231).
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference repository root).  INTEGRATION.md shows the ctypes binding used on the reference side.
 */
#ifndef ZONOS_B200_H
#define ZONOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZB_ABI_VERSION 1

#if defined(__GNUC__)
#define ZB_API __attribute__((visibility("default")))
#else
#define ZB_API
#endif

typedef int32_t zb_status;
enum {
  ZB_OK = 0,
  ZB_ERR_INVALID = 1,     /* bad argument / unsupported shape */
  ZB_ERR_CUDA = 2,        /* a CUDA runtime call failed */
  ZB_ERR_NOMEM = 3
};

typedef struct zb_ctx zb_ctx;       /* per-device context: scratch buffers, CUDA graphs */
typedef struct zb_model zb_model;   /* backbone + embeddings + heads (borrowed weight pointers) */
typedef struct zb_dac zb_dac;       /* DAC decoder (library-owned re-laid-out weights) */
typedef struct zb_gen zb_gen;       /* one device-driven generate() session */
typedef void* zb_stream;            /* cudaStream_t */

ZB_API int32_t zb_abi_version(void);
ZB_API zb_status zb_ctx_create(int32_t device, zb_ctx** out);
ZB_API zb_status zb_ctx_destroy(zb_ctx* ctx);
/* Message of the last failed call on ctx (ctx == NULL: last failed zb_ctx_create). Never NULL. */
ZB_API const char* zb_last_error(const zb_ctx* ctx);
/* How many kernels this ctx has launched so far (graph replays count their nodes). */
ZB_API int64_t zb_launch_count(const zb_ctx* ctx);

/* ------------------------------------------------------------------ model ------------------ */
enum { ZB_LAYER_ATTENTION = 0, ZB_LAYER_MAMBA2 = 1 };
enum { ZB_NORM_LAYERNORM = 0, ZB_NORM_RMSNORM = 1 };

/* One backbone layer.  Names follow the reference state_dict:
 * backbone.layers.{i}.norm / mixer.in_proj / mixer.out_proj / norm2 / mlp.fc1 / mlp.fc2
 * (zonos/backbone/_torch.py:278-281,369-370,470-471).  Unused pointers are NULL. */
typedef struct {
  int32_t kind;                 /* ZB_LAYER_* */
  int32_t _pad;
  const void* norm_w;           /* bf16 [D] */
  const void* norm_b;           /* bf16 [D]; NULL with RMSNorm */
  const void* in_proj;          /* attention: bf16 [(H+2*Hkv)*hd, D] rows q|k|v (_torch.py:401)
                                   mamba2:    bf16 [2*d_inner + 2*ngroups*d_state + nheads, D] rows z|xBC|dt */
  const void* out_proj;         /* attention: bf16 [D, H*hd]; mamba2: bf16 [D, d_inner] */
  const void* norm2_w;          /* bf16 [D] or NULL when the layer has no MLP */
  const void* norm2_b;
  const void* fc1;              /* bf16 [2*F, D] rows value|gate (_torch.py:473) */
  const void* fc2;              /* bf16 [D, F] */
  /* Mamba2 only (mamba_ssm 2.2.5 Mamba2 parameter names) */
  const void* conv_w;           /* bf16 [conv_dim, d_conv] */
  const void* conv_b;           /* bf16 [conv_dim] */
  const void* dt_bias;          /* bf16 [nheads] (the module is cast to bf16, zonos/model.py:158) */
  const void* A_log;            /* bf16 [nheads] */
  const void* D;                /* bf16 [nheads] */
  const void* mnorm_w;          /* bf16 [d_inner] gated RMSNorm weight */
} zb_layer;

typedef struct {
  int32_t d_model, n_layer, n_heads, n_heads_kv, head_dim, d_ff;
  int32_t n_codebooks;          /* 9 */
  int32_t head_vocab;           /* 1025 logits per codebook (zonos/model.py:82) */
  int32_t emb_vocab;            /* 1032 embedding rows (zonos/model.py:80-81) */
  int32_t norm_kind;            /* ZB_NORM_*; _torch.py hard-codes LayerNorm (:155,278,280) */
  int32_t rope_interleaved;     /* 1: pairs (2i,2i+1) (_torch.py:57-68); 0: rotate-half (flash-attn rotary) */
  int32_t out_proj_repeats;     /* 2 reproduces _torch.py:419-420; 1 for the mamba_ssm MHA */
  float norm_eps;
  int32_t rope_len;             /* rows of rope_table (16384, _torch.py:206) */
  const float* rope_table;      /* fp32 [rope_len, head_dim/2, 2] (cos, sin) (_torch.py:29-34) */
  const zb_layer* layers;       /* HOST array [n_layer]; copied by zb_model_create */
  const void* norm_f_w;         /* bf16 [D] backbone.norm_f */
  const void* norm_f_b;
  const void* const* embeddings;/* HOST array [n_codebooks] of device pointers bf16 [emb_vocab, D] */
  const void* heads;            /* bf16 [n_codebooks*head_vocab, D] fused_heads.weight (zonos/model.py:82,208-223) */
  /* Mamba2 dims (0 for the transformer variant) */
  int32_t d_inner, d_state, d_conv, m_headdim, m_ngroups, _pad;
} zb_model_desc;

/* Replaces: Zonos.__init__/from_local weight ownership (zonos/model.py:68-86,150-176).  Weight
 * pointers are BORROWED (they stay owned by the caller's tensors and must outlive the model). */
ZB_API zb_status zb_model_create(zb_ctx* ctx, const zb_model_desc* desc, zb_model** out);
ZB_API zb_status zb_model_destroy(zb_model* model);
/* The caller has modified weights in place (same pointers, e.g. nn.Module.load_state_dict after the first call): copies
 * the library derived from them (the tile-ordered decode weights of ZB_MEGA_TC=1, the e4m3 copy of the opt-in FP8 mode
 * ZB_FP8=1 - both environment switches read at zb_generate_begin) are rebuilt by the next zb_generate_begin.  Replaces
 * nothing in the reference - its torch modules read the live parameters (zonos/model.py:160-175). */
ZB_API zb_status zb_model_weights_changed(zb_model* model);

/* ------------------------------------------------------------------ paged KV cache ---------- */
#define ZB_PAGE_TOKENS 64
/* Replaces InferenceParams.key_value_memory_dict + lengths_per_sample (zonos/config.py:8-52) and the
 * contiguous [R,S,2,Hkv,hd] cache of _torch.py:305.  All memory is caller-owned.
 * Token s of row r, attention layer a (a-th attention layer in order), lives in page
 *   p = page_table[r*max_pages_per_row + s/ZB_PAGE_TOKENS] at
 *   kv_pages[a][p][kv(0=K,1=V)][h][s % ZB_PAGE_TOKENS][hd]   (bf16). */
typedef struct {
  int32_t rows;                 /* R */
  int32_t num_pages;            /* pages per attention layer */
  int32_t max_pages_per_row;
  int32_t _pad;
  void* kv_pages;
  const int32_t* page_table;    /* int32 [rows, max_pages_per_row] */
  int32_t* lengths;             /* int32 [rows]: tokens cached so far == lengths_per_sample; position of
                                   the next token for RoPE (_torch.py:233-234) and for the cache write (:105) */
  void* conv_state;             /* mamba2: bf16 [n_mamba, rows, conv_dim, d_conv] or NULL */
  void* ssm_state;              /* mamba2: bf16 [n_mamba, rows, nheads, headdim, d_state] or NULL */
} zb_cache;

/* ------------------------------------------------------------------ backbone plugin --------- */
/* Replaces embed_codes_static (zonos/utilities/codec_utils.py:15-37): out[b*repeat.., t, :] =
 * sum_k E_k[codes[b,k,t]] with sequential bf16 adds; each result row is written `repeat` times
 * (rows b, B+b, ... : the CFG duplication of generation_utils.py:192 / :237-238).
 * codes: int64, element strides given; out: bf16 [B*repeat, T, D]. */
ZB_API zb_status zb_embed_codes(zb_ctx* ctx, const zb_model* model, const int64_t* codes, int64_t stride_b,
                         int64_t stride_q, int64_t stride_t, int32_t B, int32_t T, int32_t repeat,
                         void* out, zb_stream stream);

/* Replaces TorchZonosBackbone.forward (zonos/backbone/_torch.py:213-238) / MambaSSMZonosBackbone.forward
 * (zonos/backbone/_mamba_ssm.py:90-119): x bf16 [R,T,D] -> y.  Token t of row r sits at position
 * cache->lengths[r] + t; its K/V are appended to the cache; attention is causal.  lengths is NOT advanced
 * (the caller does, as zonos/model.py:430-431 and utilities/tensor_ops.py:85-87 do).
 * last_only != 0: y is bf16 [R,1,D] holding only the last position (what model.py:228 keeps). */
ZB_API zb_status zb_backbone_forward(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* x,
                              int32_t T, int32_t last_only, void* y, zb_stream stream);

/* Replaces apply_heads_static + the CFG mix of Zonos._compute_logits (codec_utils.py:68-79,
 * zonos/model.py:229-233): hidden bf16 [R, D] (row stride given, elements) -> logits fp32 [B,Q,V];
 * cfg_scale == 1: logits fp32 [R,Q,V], no mix. */
ZB_API zb_status zb_heads_cfg(zb_ctx* ctx, const zb_model* model, const void* hidden, int64_t row_stride, int32_t R,
                       float cfg_scale, float* logits, zb_stream stream);

/* ------------------------------------------------------------------ sampler ----------------- */
/* Keyword arguments of sample_from_logits (zonos/sampling.py:166-177). */
typedef struct {
  float temperature, top_p, min_p, linear, conf, quad, repetition_penalty;
  int32_t top_k, repetition_penalty_window;
} zb_sampling;

/* Replaces sample_from_logits (zonos/sampling.py:166-231).  logits fp32 [B,Q,V] (not modified);
 * window: int64 [B,Q,W] recent tokens (element strides given) or NULL = no repetition penalty;
 * q: fp32 [B,Q,V] Exp(1) draws (sampling.py:29) or NULL = counter-based Philox4x32-10 draws keyed by
 * (seed, draw_index, row, column); apply_logit_bias != 0 adds the fork's EOS bias first
 * (zonos/model.py:433-437,476).  tokens: int64 [B,Q]. */
ZB_API zb_status zb_sample_from_logits(zb_ctx* ctx, const zb_sampling* params, const float* logits, int32_t B,
                                int32_t Q, int32_t V, const int64_t* window, int64_t win_stride_b,
                                int64_t win_stride_q, int32_t W, const float* q, uint64_t seed,
                                uint64_t draw_index, int32_t apply_logit_bias, int64_t* tokens,
                                zb_stream stream);

/* ------------------------------------------------------------------ device-driven generate -- */
/* Replaces the hot loop of Zonos.generate (zonos/model.py:404-509): prefill, first sample, and per step
 * embed -> backbone -> heads/CFG -> logit bias -> repetition penalty -> sampler -> EOS state machine
 * (model.py:483-497, utilities/tensor_ops.py:155-211) -> frame write (tensor_ops.py:12-53) -> counters
 * (tensor_ops.py:56-105), without host round trips.  The early-exit test of tensor_ops.py:90-103 is
 * evaluated ON THE DEVICE on exactly the steps the reference evaluates it, so the final offset (hence the
 * returned length) is the reference's. */
typedef struct {
  int32_t B;                    /* utterances */
  int32_t Q;                    /* codebooks (9) */
  int32_t T_delayed;            /* columns of delayed = P + N + Q */
  int32_t prefix_audio_len;     /* P */
  int32_t cond_len;             /* Lc */
  int32_t max_new_tokens;       /* N */
  int64_t* delayed;             /* int64 [B,Q,T_delayed], delay pattern applied (codebook_pattern.py:31-32), -1 = unknown */
  const void* prefix_conditioning; /* bf16 [2B, Lc, D] */
  float cfg_scale;
  zb_sampling sampling;
  const float* q_stream;        /* optional fp32 [q_calls,B,Q,V] explicit Exp(1) draws (tests) */
  int32_t q_calls;
  int32_t _pad;
  uint64_t seed;                /* Philox key when q_stream == NULL */
  float* logits_trace;          /* optional fp32 [trace_calls,B,Q,V]: logits handed to each sample call */
  int32_t trace_calls;
  int32_t _pad2;
} zb_gen_desc;

typedef struct {                /* host-visible progress, filled by zb_generate_poll */
  int32_t done;                 /* loop finished (early exit fired or offset reached T_delayed) */
  int32_t offset;               /* model.py's `offset` after the last executed step */
  int32_t steps;                /* executed loop iterations (model.py's `step`) */
  int32_t max_steps;
} zb_gen_progress;

ZB_API zb_status zb_generate_begin(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const zb_gen_desc* desc,
                            zb_gen** out, zb_stream stream);         /* prefill + first sample (model.py:408-431) */
ZB_API zb_status zb_generate_steps(zb_gen* gen, int32_t n_steps, zb_stream stream);   /* enqueue up to n loop iterations */
ZB_API zb_status zb_generate_poll(zb_gen* gen, zb_gen_progress* out, zb_stream stream); /* SYNCHRONISES the stream */
/* Non-blocking: progress as last published by the device into host-mapped memory (the closing CTA of every step
 * writes it); pair it with an event recorded after an earlier chunk to bound how far ahead the host enqueues. */
ZB_API zb_status zb_generate_peek(zb_gen* gen, zb_gen_progress* out);
ZB_API zb_status zb_generate_end(zb_gen* gen);

/* ------------------------------------------------------------------ DAC decoder -------------- */
/* Weights of transformers.DacModel (decode side), fp32 device pointers, borrowed only during create.
 * Arrays are in module order; conv weights [Cout,Cin,K], transposed conv [Cin,Cout,K] as torch stores them. */
typedef struct {
  int32_t n_codebooks, codebook_size, codebook_dim, latent_dim;    /* 9, 1024, 8, 1024 */
  int32_t channels;             /* 1536 */
  int32_t n_blocks;             /* 4 */
  int32_t strides[8];           /* 8,8,4,2 */
  const float* const* tensors;  /* HOST array of device pointers, order documented in zonos_b200/autoencoder.py */
  int32_t n_tensors;
  int32_t _pad;
} zb_dac_desc;

/* Replaces DACAutoencoder.decode (zonos/autoencoder.py:119-140 -> transformers modeling_dac.py:608-638):
 * codes int64 [B,Q,T] -> wav fp32 [B, 1, 512*T]. */
ZB_API zb_status zb_dac_create(zb_ctx* ctx, const zb_dac_desc* desc, zb_dac** out, zb_stream stream);
ZB_API zb_status zb_dac_destroy(zb_dac* dac);
ZB_API zb_status zb_dac_decode(zb_ctx* ctx, const zb_dac* dac, const int64_t* codes, int32_t B, int32_t T, float* wav,
                        zb_stream stream);

/* ------------------------------------------------------------------ DAC encoder -------------- */
/* Weights of transformers.DacModel (encode side), fp32 device pointers in module order, BORROWED for the call only:
 * encoder.conv1 {weight, bias}; per block: 3 x res_unit {snake1.alpha, conv1.weight, conv1.bias, snake2.alpha, conv2.weight,
 * conv2.bias}, snake1.alpha, conv1 {weight, bias}; encoder.snake1.alpha, encoder.conv2 {weight, bias}; per codebook:
 * in_proj {weight, bias}, codebook.weight, out_proj {weight, bias}  (zonos_b200/autoencoder.py:dac_encoder_tensor_order). */
typedef struct {
  int32_t n_codebooks, codebook_size, codebook_dim, latent_dim;    /* 9, 1024, 8, 1024 */
  int32_t hidden;               /* 64 */
  int32_t n_blocks;             /* 4 */
  int32_t strides[8];           /* 2,4,8,8 */
  const float* const* tensors;  /* HOST array of device pointers */
  int32_t n_tensors;
  int32_t _pad;
} zb_dac_enc_desc;

/* Replaces DACAutoencoder.encode (zonos/autoencoder.py:104-117 -> transformers modeling_dac.py:581-640, :442-473, :281-343):
 * wav fp32 [B, 1, L] at 44.1 kHz, L a multiple of 512 (what `preprocess`, autoencoder.py:80-100, returns) -> codes int64
 * [B, Q, L/512].  fp32 throughout like the reference (no autocast on this path).  `workspace` is caller-owned device
 * memory of at least zb_dac_encode_workspace_bytes(B, L) bytes. */
ZB_API size_t zb_dac_encode_workspace_bytes(int32_t B, int64_t L);
ZB_API zb_status zb_dac_encode(zb_ctx* ctx, const zb_dac_enc_desc* desc, const float* wav, int32_t B, int64_t L, int64_t* codes,
                        void* workspace, size_t workspace_bytes, zb_stream stream);

/* ------------------------------------------------------------------ diagnostics -------------- */
/* Launches ONE production decode kernel `iters` times back to back (layer, layer+1, ... cyclically, so every launch
 * streams its weights from HBM) on scratch activations with `rows` activation rows (1..8), so a benchmark can time it
 * alone with CUDA events: which = 1 out_proj GEMV, 2 norm2+fc1+SiLU GEMV, 3 fc2+residual GEMV. */
ZB_API zb_status zb_bench_kernel(zb_ctx* ctx, const zb_model* model, int32_t layer, int32_t which, int32_t rows,
                                 int32_t iters, zb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* ZONOS_B200_H */
