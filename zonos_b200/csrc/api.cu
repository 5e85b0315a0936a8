// C-ABI entry points of libzonos_b200.so (see include/zonos_b200.h).
#include <stdarg.h>

#include <algorithm>

#include "internal.h"

thread_local std::string g_zb_create_error;

zb_status zb_fail(zb_ctx* ctx, zb_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_zb_create_error = buf;
  return code;
}

static zb_status arena_reserve(zb_ctx* ctx, void** p, size_t* have, size_t bytes) {
  if (bytes <= *have) return ZB_OK;
  if (*p) { ZB_CUDA(ctx, cudaDeviceSynchronize()); ZB_CUDA(ctx, cudaFree(*p)); *p = nullptr; *have = 0; }
  const size_t want = (bytes + (size_t)(1 << 20) - 1) / (1 << 20) * (1 << 20);
  ZB_CUDA(ctx, cudaMalloc(p, want));
  *have = want;
  return ZB_OK;
}

zb_status zb_scratch_reserve(zb_ctx* ctx, size_t bytes) {
  if (bytes > ctx->scratch_bytes && ctx->scratch_pins > 0)
    return zb_fail(ctx, ZB_ERR_INVALID, "backbone scratch would have to grow (%zu > %zu bytes) while a generate session is live",
                   bytes, ctx->scratch_bytes);
  return arena_reserve(ctx, &ctx->scratch, &ctx->scratch_bytes, bytes);
}
zb_status zb_dac_scratch_reserve(zb_ctx* ctx, size_t bytes) {
  return arena_reserve(ctx, &ctx->dac_scratch, &ctx->dac_scratch_bytes, bytes);
}
zb_status zb_tc_workspace_reserve(zb_ctx* ctx, size_t bytes) {
  // fixed 64 MB, allocated once: [0, 48 MB) split-K partials (at most 296 tiles x 128 x 256 fp32 = 39 MB), [48, 64 MB) row staging;
  // it never moves, so CUDA graphs may bake its address
  if (bytes > ((size_t)48 << 20)) return zb_fail(ctx, ZB_ERR_INVALID, "split-K workspace request of %zu bytes exceeds 48 MB", bytes);
  return arena_reserve(ctx, &ctx->tc_ws, &ctx->tc_ws_bytes, (size_t)64 << 20);
}

extern "C" {

int32_t zb_abi_version(void) { return ZB_ABI_VERSION; }

zb_status zb_ctx_create(int32_t device, zb_ctx** out) {
  if (!out) return zb_fail(nullptr, ZB_ERR_INVALID, "zb_ctx_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return zb_fail(nullptr, ZB_ERR_CUDA, "zb_ctx_create: no CUDA device (%s); this library has no CPU path",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return zb_fail(nullptr, ZB_ERR_INVALID, "zb_ctx_create: device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return zb_fail(nullptr, ZB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return zb_fail(nullptr, ZB_ERR_INVALID, "zb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                   prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return zb_fail(nullptr, ZB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  zb_ctx* ctx = new zb_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if ((e = cudaMalloc(&ctx->counters, ZB_NUM_COUNTERS * sizeof(int32_t))) != cudaSuccess ||
      (e = cudaMemset(ctx->counters, 0, ZB_NUM_COUNTERS * sizeof(int32_t))) != cudaSuccess) {
    delete ctx;
    return zb_fail(nullptr, ZB_ERR_NOMEM, "zb_ctx_create: %s", cudaGetErrorString(e));
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    cudaFree(ctx->counters);
    delete ctx;
    return zb_fail(nullptr, ZB_ERR_CUDA, "zb_ctx_create: %s", cudaGetErrorString(e));
  }
  *out = ctx;
  return ZB_OK;
}

zb_status zb_ctx_destroy(zb_ctx* ctx) {
  if (!ctx) return ZB_OK;
  zb_device_guard dev_guard(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->dac_scratch) cudaFree(ctx->dac_scratch);
  if (ctx->tc_ws) cudaFree(ctx->tc_ws);
  if (ctx->counters) cudaFree(ctx->counters);
  if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
  for (zb_gen_slab& sl : ctx->gen_slabs) {
    if (sl.dev) cudaFree(sl.dev);
    if (sl.host) cudaFreeHost(sl.host);
    if (sl.stage) cudaFreeHost(sl.stage);
    if (sl.stage_ev) cudaEventDestroy(sl.stage_ev);
  }
  delete ctx;
  return ZB_OK;
}

const char* zb_last_error(const zb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_zb_create_error.c_str(); }
int64_t zb_launch_count(const zb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------
zb_status zb_model_create(zb_ctx* ctx, const zb_model_desc* desc, zb_model** out) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, desc && out, "zb_model_create: null argument");
  ZB_REQUIRE(ctx, desc->n_layer >= 1 && desc->layers, "zb_model_create: no layers");
  ZB_REQUIRE(ctx, desc->d_model % 256 == 0, "zb_model_create: d_model %d must be a multiple of 256", desc->d_model);
  ZB_REQUIRE(ctx, desc->head_dim == 128, "zb_model_create: head_dim %d unsupported (128)", desc->head_dim);
  ZB_REQUIRE(ctx, desc->n_codebooks >= 1 && desc->n_codebooks <= 16, "zb_model_create: n_codebooks %d", desc->n_codebooks);
  ZB_REQUIRE(ctx, desc->rope_table && desc->rope_len > 0, "zb_model_create: rope_table missing");
  zb_model* m = new zb_model();
  m->ctx = ctx;
  m->d = *desc;
  m->layers.assign(desc->layers, desc->layers + desc->n_layer);
  m->d.layers = m->layers.data();
  if (desc->embeddings) m->emb.assign(desc->embeddings, desc->embeddings + desc->n_codebooks);
  m->d.embeddings = m->emb.data();
  for (int i = 0; i < desc->n_layer; ++i) {
    const zb_layer& L = m->layers[i];
    m->attn_index.push_back(L.kind == ZB_LAYER_ATTENTION ? m->n_attn++ : -1);
    m->mamba_index.push_back(L.kind == ZB_LAYER_MAMBA2 ? m->n_mamba++ : -1);
    if (L.kind == ZB_LAYER_ATTENTION && !(L.norm_w && L.in_proj && L.out_proj && L.norm2_w && L.fc1 && L.fc2)) {
      delete m;
      return zb_fail(ctx, ZB_ERR_INVALID, "zb_model_create: layer %d misses a weight pointer", i);
    }
  }
  *out = m;
  return ZB_OK;
}

zb_status zb_model_destroy(zb_model* model) { delete model; return ZB_OK; }
zb_status zb_model_weights_changed(zb_model* model) {
  if (model) model->tcw_valid = model->f8w_valid = false;   // derived copies are rebuilt by the next generate session
  return ZB_OK;
}

// ---------------------------------------------------------------------------------------------
zb_status zb_embed_codes(zb_ctx* ctx, const zb_model* model, const int64_t* codes, int64_t stride_b, int64_t stride_q,
                         int64_t stride_t, int32_t B, int32_t T, int32_t repeat, void* out, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, model && codes && out && B >= 1 && T >= 1 && repeat >= 1, "zb_embed_codes: bad arguments");
  ZB_REQUIRE(ctx, (int)model->emb.size() == model->d.n_codebooks, "zb_embed_codes: model has no embedding tables");
  zb_embed_launch L;
  L.model = model; L.codes = codes; L.sb = stride_b; L.sq = stride_q; L.st = stride_t; L.B = B; L.T = T; L.repeat = repeat;
  L.out = (bf16*)out; L.out_rs = (int64_t)T * model->d.d_model;
  return zb_launch_embed(ctx, L, (cudaStream_t)stream);
}

zb_status zb_backbone_forward(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const void* x, int32_t T, int32_t last_only,
                              void* y, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, model && cache && x && y && T >= 1, "zb_backbone_forward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const zb_model_desc& d = model->d;
  const int R = cache->rows;
  // the host does not know lengths[]; size the split-KV grid for the whole page table
  const int max_kv = cache->max_pages_per_row * ZB_PAGE_TOKENS;
  const size_t xbytes = ((size_t)R * T * d.d_model * 2 + 255) / 256 * 256;
  const size_t need = xbytes + zb_backbone_scratch_bytes(model, R, T, max_kv);
  if (zb_status st = zb_scratch_reserve(ctx, need)) return st;
  // residual stream lives at the END of the arena (the layer scratch is carved from the front)
  bf16* xbuf = (bf16*)((char*)ctx->scratch + (ctx->scratch_bytes - xbytes));
  ZB_CUDA(ctx, cudaMemcpyAsync(xbuf, x, (size_t)R * T * d.d_model * 2, cudaMemcpyDeviceToDevice, s));
  if (zb_status st = zb_run_layers(ctx, model, cache, xbuf, R, T, max_kv, nullptr, 0, s)) return st;
  return zb_launch_final_norm(ctx, model, xbuf, R, T, last_only, (bf16*)y, s);
}

zb_status zb_heads_cfg(zb_ctx* ctx, const zb_model* model, const void* hidden, int64_t row_stride, int32_t R, float cfg_scale,
                       float* logits, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, model && hidden && logits && R >= 1 && model->d.heads, "zb_heads_cfg: bad arguments");
  return zb_launch_heads(ctx, model, (const bf16*)hidden, row_stride, R, 0, cfg_scale, logits, nullptr, 0, (cudaStream_t)stream);
}

zb_status zb_sample_from_logits(zb_ctx* ctx, const zb_sampling* params, const float* logits, int32_t B, int32_t Q, int32_t V,
                                const int64_t* window, int64_t win_stride_b, int64_t win_stride_q, int32_t W, const float* q,
                                uint64_t seed, uint64_t draw_index, int32_t apply_logit_bias, int64_t* tokens, zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, params && logits && tokens, "zb_sample_from_logits: null argument");
  zb_sample_launch L;
  L.logits = logits; L.B = B; L.Q = Q; L.V = V; L.window = window; L.wsb = win_stride_b; L.wsq = win_stride_q; L.W = W;
  L.q = q; L.seed = seed; L.draw_index = draw_index; L.sp = *params; L.apply_bias = apply_logit_bias; L.tokens = tokens;
  return zb_launch_sample(ctx, L, (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// device-driven generate
// ---------------------------------------------------------------------------------------------
struct zb_gen {
  zb_ctx* ctx;
  const zb_model* model;
  zb_cache cache;
  zb_gen_desc d;
  zb_loop_state* st = nullptr;      // device
  int32_t* st_host = nullptr;       // host-mapped progress words (first 8 words of zb_loop_state)
  int32_t* st_host_dev = nullptr;   // its device alias
  float* logits = nullptr;          // [B,Q,V]
  bf16* xdec = nullptr;             // [2B, D] residual stream of one decode step
  void* mega_layers = nullptr;      // device array of per-layer pointers for the persistent decode kernel
  unsigned* mega_bar = nullptr;     // sync words of the persistent kernel ([1] = epoch of the tagged arena)
  uint32_t* mega_arena = nullptr;   // tagged activation words exchanged between its CTAs
  bool mega = false;
  bool tc = false;                  // persistent tcgen05 decode step (decode_tc.cu); mega_layers / mega_bar / mega_arena then hold its table / barrier / arena
  int slab = -1;                    // index into ctx->gen_slabs
  cudaGraphExec_t graph = nullptr;
  int64_t launches_per_step = 0;
  int max_steps = 0, steps_enqueued = 0, max_kv = 0;
  bool pinned = false;
};

namespace {
__global__ void init_state_kernel(zb_loop_state* st, int B, int offset, int max_steps) {
  if (threadIdx.x == 0) {
    st->offset = offset; st->step_idx = 0; st->done = 0; st->steps = 0; st->draw_idx = 0; st->arrive = 0; st->max_steps = max_steps;
  }
  for (int b = threadIdx.x; b < B; b += blockDim.x) { st->remaining[b] = max_steps; st->stopping[b] = 0; st->row_arrive[b] = 0; }
}

zb_status enqueue_step(zb_gen* g, cudaStream_t s) {
  zb_ctx* ctx = g->ctx;
  const zb_model_desc& md = g->model->d;
  const int B = g->d.B, R = 2 * B;
  zb_embed_launch E;
  E.model = g->model; E.codes = g->d.delayed; E.sb = (int64_t)g->d.Q * g->d.T_delayed; E.sq = g->d.T_delayed; E.st = 1;
  E.B = B; E.T = 1; E.repeat = 2; E.out = g->xdec; E.out_rs = md.d_model; E.loop = g->st; E.T_delayed = g->d.T_delayed;
  if (g->tc) {
    // one cooperative launch: embed + all layers + heads, tcgen05 consumer
    if (zb_status st = zb_launch_decode_tc(ctx, g->model, &g->cache, g->mega_layers, g->mega_bar, g->mega_arena, R, g->d.cfg_scale, g->logits, g->d.delayed,
                                           g->d.T_delayed, g->st, s)) return st;
  } else if (g->mega) {
    // one cooperative launch: embed + all layers + heads
    if (zb_status st = zb_launch_decode_step(ctx, g->model, &g->cache, g->mega_layers, g->mega_bar, g->mega_arena, R, g->max_kv, g->d.cfg_scale, g->logits,
                                             g->d.delayed, g->d.T_delayed, g->st, s)) return st;
  } else {
    if (zb_status st = zb_launch_embed(ctx, E, s)) return st;
    if (zb_status st = zb_run_layers(ctx, g->model, &g->cache, g->xdec, R, 1, g->max_kv, g->st, g->d.T_delayed, s)) return st;
    if (zb_status st = zb_launch_heads(ctx, g->model, g->xdec, md.d_model, R, 1, g->d.cfg_scale, g->logits, g->st, g->d.T_delayed, s)) return st;
  }
  zb_sample_launch L;
  L.logits = g->logits; L.B = B; L.Q = g->d.Q; L.V = md.head_vocab; L.sp = g->d.sampling; L.apply_bias = 1; L.seed = g->d.seed;
  L.st = g->st; L.delayed = g->d.delayed; L.T = g->d.T_delayed; L.ctx_len = g->d.max_new_tokens < 100 ? g->d.max_new_tokens : 100;
  L.lengths = g->cache.lengths; L.q_stream = g->d.q_stream; L.q_calls = g->d.q_calls; L.logits_trace = g->d.logits_trace;
  L.trace_calls = g->d.trace_calls; L.first = 0; L.mirror = g->st_host_dev; L.reset_word = nullptr;
  return zb_launch_sample(ctx, L, s);
}
}  // namespace

extern "C" {

zb_status zb_generate_begin(zb_ctx* ctx, const zb_model* model, const zb_cache* cache, const zb_gen_desc* desc, zb_gen** out,
                            zb_stream stream) {
  if (!ctx) return ZB_ERR_INVALID;
  zb_device_guard dev_guard(ctx);
  ZB_REQUIRE(ctx, model && cache && desc && out, "zb_generate_begin: null argument");
  const zb_model_desc& md = model->d;
  const int B = desc->B, R = 2 * B, Q = desc->Q;
  ZB_REQUIRE(ctx, B >= 1 && B <= ZB_MAX_B, "zb_generate_begin: batch %d unsupported (1..%d)", B, ZB_MAX_B);
  ZB_REQUIRE(ctx, Q == md.n_codebooks, "zb_generate_begin: Q=%d but the model has %d codebooks", Q, md.n_codebooks);
  ZB_REQUIRE(ctx, desc->cfg_scale != 1.0f, "TODO: add support for cfg_scale=1");   // zonos/model.py:399
  ZB_REQUIRE(ctx, cache->rows == R, "zb_generate_begin: cache has %d rows, need 2*B=%d", cache->rows, R);
  ZB_REQUIRE(ctx, desc->T_delayed == desc->prefix_audio_len + desc->max_new_tokens + Q, "zb_generate_begin: T_delayed mismatch");
  ZB_REQUIRE(ctx, desc->delayed && desc->prefix_conditioning && desc->cond_len >= 1, "zb_generate_begin: missing tensors");
  const int P = desc->prefix_audio_len, Lc = desc->cond_len;
  const int Tp = Lc + P + 1;                               // prefill tokens per row (model.py:428)
  const int total = Lc + desc->T_delayed;                  // model.py:409 seq_len
  ZB_REQUIRE(ctx, total <= cache->max_pages_per_row * ZB_PAGE_TOKENS, "zb_generate_begin: cache too small for %d tokens", total);
  ZB_REQUIRE(ctx, total <= md.rope_len, "zb_generate_begin: %d tokens exceed the rotary table", total);
  cudaStream_t s = (cudaStream_t)stream;

  zb_gen* g = new zb_gen();
  g->ctx = ctx; g->model = model; g->cache = *cache; g->d = *desc;
  g->max_kv = (total + ZB_PAGE_TOKENS - 1) / ZB_PAGE_TOKENS * ZB_PAGE_TOKENS;
  auto fail = [&](zb_status e) { zb_generate_end(g); return e; };
#define G_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { zb_fail(ctx, ZB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); return fail(ZB_ERR_CUDA); } } while (0)
  // one slab per session (reused across sessions, see zb_gen_slab): loop state | logits | decode residual | layer
  // table | sync words | tagged activation arena
  g->tc = zb_tc_supported(model, R);
  g->mega = g->tc || zb_mega_supported(model, R);           // both persistent paths: no CUDA graph, session table + arena in the slab
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t table_bytes = !g->mega ? 0 : g->tc ? zb_tc_table_bytes(model) : zb_mega_layers_bytes(model);
  const size_t arena_bytes = !g->mega ? 0 : g->tc ? zb_tc_arena_bytes(model, R) : zb_mega_arena_bytes(model, R);
  const size_t o_st = 0, o_logits = o_st + up(sizeof(zb_loop_state)), o_x = o_logits + up((size_t)B * Q * md.head_vocab * 4);
  const size_t o_layers = o_x + up((size_t)R * md.d_model * 2), o_bar = o_layers + up(table_bytes);
  const size_t o_arena = o_bar + 256, slab_need = o_arena + up(arena_bytes);
  {
    int pick = -1;
    for (size_t i = 0; i < ctx->gen_slabs.size(); ++i)
      if (!ctx->gen_slabs[i].in_use && (pick < 0 || ctx->gen_slabs[i].dev_bytes >= slab_need)) { pick = (int)i; if (ctx->gen_slabs[i].dev_bytes >= slab_need) break; }
    if (pick < 0) { ctx->gen_slabs.push_back(zb_gen_slab()); pick = (int)ctx->gen_slabs.size() - 1; }
    zb_gen_slab& sl = ctx->gen_slabs[pick];
    sl.in_use = true;
    g->slab = pick;
    if (sl.dev_bytes < slab_need) {
      if (sl.dev) { G_CUDA(cudaStreamSynchronize(s)); G_CUDA(cudaFree(sl.dev)); sl.dev = nullptr; sl.dev_bytes = 0; }
      G_CUDA(cudaMalloc(&sl.dev, slab_need));
      sl.dev_bytes = slab_need;
      sl.layers.clear();
    }
    if (!sl.host) {
      G_CUDA(cudaHostAlloc(&sl.host, 16 * sizeof(int32_t), cudaHostAllocMapped));
      G_CUDA(cudaHostGetDevicePointer((void**)&sl.host_dev, sl.host, 0));
    }
    char* base = (char*)sl.dev;
    g->st = (zb_loop_state*)(base + o_st); g->logits = (float*)(base + o_logits); g->xdec = (bf16*)(base + o_x);
    g->st_host = sl.host; g->st_host_dev = sl.host_dev;
    memset(g->st_host, 0, 16 * sizeof(int32_t));
    g->st_host[0] = desc->prefix_audio_len + 1;
    if (g->mega) {
      g->mega_layers = base + o_layers; g->mega_bar = (unsigned*)(base + o_bar); g->mega_arena = (uint32_t*)(base + o_arena);
      std::vector<unsigned char> hb(table_bytes);
      if (zb_status st = g->tc ? zb_tc_table_build(ctx, model, cache, R, g->mega_arena, hb.data()) : zb_mega_layers_build(ctx, model, cache, hb.data(), s)) return fail(st);
      if (sl.layers != hb || sl.layers_off != o_layers) {  // same model and cache as the slab's last session: already there
        sl.layers = hb; sl.layers_off = o_layers;
        // a fresh KV allocation per generate() changes the table every call: stage it in pinned memory so the upload
        // needs no stream synchronisation (the event only guards the staging buffer against being rewritten too early)
        if (sl.stage_bytes < hb.size()) {
          if (sl.stage) { G_CUDA(cudaEventSynchronize(sl.stage_ev)); G_CUDA(cudaFreeHost(sl.stage)); sl.stage = nullptr; sl.stage_bytes = 0; }
          G_CUDA(cudaHostAlloc(&sl.stage, hb.size(), cudaHostAllocDefault));
          sl.stage_bytes = hb.size();
          if (!sl.stage_ev) G_CUDA(cudaEventCreateWithFlags(&sl.stage_ev, cudaEventDisableTiming));
        } else {
          G_CUDA(cudaEventSynchronize(sl.stage_ev));
        }
        memcpy(sl.stage, hb.data(), hb.size());
        G_CUDA(cudaMemcpyAsync(g->mega_layers, sl.stage, hb.size(), cudaMemcpyHostToDevice, s));
        G_CUDA(cudaEventRecord(sl.stage_ev, s));
      }
      G_CUDA(cudaMemsetAsync(g->mega_bar, 0, 256, s));
      G_CUDA(cudaMemsetAsync(g->mega_arena, 0, arena_bytes, s));   // tag 0 = never written (decode.cu) / arrival counters (decode_tc.cu)
    }
  }
  const int offset0 = P + 1;
  g->max_steps = desc->T_delayed - offset0;               // model.py:440
  init_state_kernel<<<1, 256, 0, s>>>(g->st, B, offset0, g->max_steps);
  ctx->launches++;

  // ---- prefill (generation_utils.py:236-244): [prefix_conditioning ; embed(delayed[..., :P+1])] ----
  const size_t xbytes = ((size_t)R * Tp * md.d_model * 2 + 255) / 256 * 256;
  const size_t need = xbytes + std::max(zb_backbone_scratch_bytes(model, R, Tp, g->max_kv), zb_backbone_scratch_bytes(model, R, 1, g->max_kv));
  if (zb_status st = zb_scratch_reserve(ctx, need)) return fail(st);
  ctx->scratch_pins++;
  g->pinned = true;
  bf16* xbuf = (bf16*)((char*)ctx->scratch + (ctx->scratch_bytes - xbytes));
  G_CUDA(cudaMemcpy2DAsync(xbuf, (size_t)Tp * md.d_model * 2, desc->prefix_conditioning, (size_t)Lc * md.d_model * 2,
                           (size_t)Lc * md.d_model * 2, R, cudaMemcpyDeviceToDevice, s));
  zb_embed_launch E;
  E.model = model; E.codes = desc->delayed; E.sb = (int64_t)Q * desc->T_delayed; E.sq = desc->T_delayed; E.st = 1;
  E.B = B; E.T = P + 1; E.repeat = 2; E.out = xbuf + (size_t)Lc * md.d_model; E.out_rs = (int64_t)Tp * md.d_model;
  if (zb_status st = zb_launch_embed(ctx, E, s)) return fail(st);
  if (zb_status st = zb_run_layers(ctx, model, cache, xbuf, R, Tp, g->max_kv, nullptr, 0, s)) return fail(st);
  if (zb_status st = zb_launch_heads(ctx, model, xbuf + (size_t)(Tp - 1) * md.d_model, (int64_t)Tp * md.d_model, R, 1, desc->cfg_scale,
                                     g->logits, nullptr, 0, s)) return fail(st);
  // ---- first sample: no repetition penalty, no logit bias (model.py:423-431) ----
  zb_sample_launch L;
  L.logits = g->logits; L.B = B; L.Q = Q; L.V = md.head_vocab; L.sp = desc->sampling; L.apply_bias = 0; L.seed = desc->seed;
  L.st = g->st; L.delayed = desc->delayed; L.T = desc->T_delayed; L.ctx_len = desc->max_new_tokens < 100 ? desc->max_new_tokens : 100;
  L.lengths = cache->lengths; L.q_stream = desc->q_stream; L.q_calls = desc->q_calls; L.logits_trace = desc->logits_trace;
  L.trace_calls = desc->trace_calls; L.first = 1; L.prefix_len = Tp; L.mirror = g->st_host_dev;
  if (zb_status st = zb_launch_sample(ctx, L, s)) return fail(st);

  // ---- capture one loop iteration as a CUDA graph (all positions come from device state).  The persistent path is
  // two launches per step (cooperative step kernel + sampler): no graph needed ----
  if (!g->mega) {
    const int64_t before = ctx->launches;
    cudaStream_t cs = ctx->capture_stream;
    G_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    zb_status st = enqueue_step(g, cs);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    if (st != ZB_OK) { if (graph) cudaGraphDestroy(graph); return fail(st); }
    if (e != cudaSuccess) { zb_fail(ctx, ZB_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e)); return fail(ZB_ERR_CUDA); }
    g->launches_per_step = ctx->launches - before;
    ctx->launches = before;
    e = cudaGraphInstantiate(&g->graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { zb_fail(ctx, ZB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); return fail(ZB_ERR_CUDA); }
  }
#undef G_CUDA
  *out = g;
  return ZB_OK;
}

zb_status zb_generate_steps(zb_gen* gen, int32_t n_steps, zb_stream stream) {
  if (!gen) return ZB_ERR_INVALID;
  zb_ctx* ctx = gen->ctx;
  zb_device_guard dev_guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < n_steps && gen->steps_enqueued < gen->max_steps; ++i) {
    if (gen->mega) {
      if (zb_status st = enqueue_step(gen, s)) return st;
    } else {
      ZB_CUDA(ctx, cudaGraphLaunch(gen->graph, s));
      ctx->launches += gen->launches_per_step;
    }
    gen->steps_enqueued++;
  }
  return ZB_OK;
}

zb_status zb_generate_poll(zb_gen* gen, zb_gen_progress* out, zb_stream stream) {
  if (!gen || !out) return ZB_ERR_INVALID;
  zb_ctx* ctx = gen->ctx;
  zb_device_guard dev_guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* tmp = gen->st_host + 8;                 // second half of the pinned block: staging for an exact read-back
  ZB_CUDA(ctx, cudaMemcpyAsync(tmp, gen->st, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  ZB_CUDA(ctx, cudaStreamSynchronize(s));
  out->offset = tmp[0];
  out->steps = tmp[3];
  out->max_steps = gen->max_steps;
  out->done = tmp[2] || gen->steps_enqueued >= gen->max_steps;
  return ZB_OK;
}

zb_status zb_generate_peek(zb_gen* gen, zb_gen_progress* out) {
  if (!gen || !out) return ZB_ERR_INVALID;
  volatile int32_t* m = gen->st_host;
  out->offset = m[0];
  out->steps = m[3];
  out->done = m[2];
  out->max_steps = gen->max_steps;
  return ZB_OK;
}

zb_status zb_generate_end(zb_gen* gen) {
  if (!gen) return ZB_OK;
  zb_device_guard dev_guard(gen->ctx);
  if (gen->pinned) gen->ctx->scratch_pins--;
  if (gen->graph) cudaGraphExecDestroy(gen->graph);
  if (gen->slab >= 0) gen->ctx->gen_slabs[gen->slab].in_use = false;   // the memory stays with the context for the next session
  delete gen;
  return ZB_OK;
}

}  // extern "C"
