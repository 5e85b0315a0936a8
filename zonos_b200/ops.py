"""`torch.library` custom ops `zonos_b200::*`: the thin PyTorch layer over the C ABI (include/zonos_b200.h).

Every op is a tensor-in / tensor-out wrapper of exactly one `zb_*` entry point of libzonos_b200.so; native handles
(`zb_model*`, `zb_dac*`, `zb_gen*`) travel as Python ints.  The Python classes of this package (`Zonos`,
`B200ZonosBackbone`, `DACAutoencoder`, `sample_from_logits`) call these ops and nothing else on the hot path, so the
same ops are what a maintainer of the reference binds when wiring the B200 path into `zonos/backbone/__init__.py:24-36`
and `zonos/model.py:179-234`.  There is no CPU implementation: a CPU tensor raises.

    torch.ops.zonos_b200.embed_codes        zonos/model.py:179-192          -> zb_embed_codes
    torch.ops.zonos_b200.backbone_forward   zonos/backbone/_torch.py:213-238 -> zb_backbone_forward
    torch.ops.zonos_b200.heads_cfg          zonos/model.py:194-206,225-234  -> zb_heads_cfg
    torch.ops.zonos_b200.sample_update      zonos/sampling.py:166-231       -> zb_sample_from_logits
    torch.ops.zonos_b200.decode_step        one or more whole decode steps of a generate session (embed + backbone +
                                            heads/CFG + sampler + EOS/delay bookkeeping, zonos/model.py:439-509)
                                                                            -> zb_generate_steps
    torch.ops.zonos_b200.dac_decode         zonos/autoencoder.py:119-140    -> zb_dac_decode
    torch.ops.zonos_b200.dac_encode         zonos/autoencoder.py:104-117    -> zb_dac_encode
"""
import ctypes as C
from typing import List, Optional

import torch
from torch.library import custom_op

from . import _lib


def _cuda(t: torch.Tensor, what: str):
    if t.device.type != "cuda":
        raise RuntimeError(f"zonos_b200::{what} runs on CUDA (B200) only, got a tensor on '{t.device}'; there is no CPU path")


@custom_op("zonos_b200::embed_codes", mutates_args=())
def embed_codes(model: int, codes: torch.Tensor, repeat: int, d_model: int) -> torch.Tensor:
    """int64 [B,Q,T] -> bf16 [B*repeat,T,D]: sequential bf16 sum of the Q codebook embeddings."""
    _cuda(codes, "embed_codes")
    B, Q, T = codes.shape
    out = torch.empty((B * repeat, T, d_model), dtype=torch.bfloat16, device=codes.device)
    ctx = _lib.context(codes.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_embed_codes(ctx.handle, C.c_void_p(model), _lib.ptr(codes), codes.stride(0), codes.stride(1),
                                         codes.stride(2), B, T, repeat, _lib.ptr(out), _lib.stream_ptr(codes.device)))
    return out


@embed_codes.register_fake
def _(model, codes, repeat, d_model):
    B, Q, T = codes.shape
    return codes.new_empty((B * repeat, T, d_model), dtype=torch.bfloat16)


@custom_op("zonos_b200::backbone_forward", mutates_args=("kv_pages", "conv_state", "ssm_state"))
def backbone_forward(model: int, x: torch.Tensor, kv_pages: torch.Tensor, page_table: torch.Tensor, lengths: torch.Tensor,
                     conv_state: Optional[torch.Tensor], ssm_state: Optional[torch.Tensor], last_only: bool) -> torch.Tensor:
    """bf16 [R,T,D] -> [R,T or 1,D]; appends K/V (and the Mamba2 states) of the T tokens at positions lengths[r]...
    `lengths` itself is advanced by the caller like `InferenceParams.lengths_per_sample` in the reference."""
    _cuda(x, "backbone_forward")
    R, T, D = x.shape
    y = torch.empty((R, 1 if last_only else T, D), dtype=x.dtype, device=x.device)
    cache = _lib.zb_cache()
    cache.rows, cache.num_pages, cache.max_pages_per_row = R, kv_pages.shape[1], page_table.shape[1]
    cache.kv_pages, cache.page_table, cache.lengths = kv_pages.data_ptr(), page_table.data_ptr(), lengths.data_ptr()
    cache.conv_state = conv_state.data_ptr() if conv_state is not None else None
    cache.ssm_state = ssm_state.data_ptr() if ssm_state is not None else None
    ctx = _lib.context(x.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_backbone_forward(ctx.handle, C.c_void_p(model), C.byref(cache), _lib.ptr(x), T, int(last_only),
                                              _lib.ptr(y), _lib.stream_ptr(x.device)))
    return y


@backbone_forward.register_fake
def _(model, x, kv_pages, page_table, lengths, conv_state, ssm_state, last_only):
    R, T, D = x.shape
    return x.new_empty((R, 1 if last_only else T, D))


@custom_op("zonos_b200::heads_cfg", mutates_args=())
def heads_cfg(model: int, hidden: torch.Tensor, cfg_scale: float, n_codebooks: int, head_vocab: int) -> torch.Tensor:
    """bf16 [R,D] -> fp32 logits [R or R/2, Q, V]: fused heads, fp32 cast, u + (c - u) * cfg_scale when cfg_scale != 1."""
    _cuda(hidden, "heads_cfg")
    R, D = hidden.shape
    rows = R // 2 if cfg_scale != 1.0 else R
    logits = torch.empty((rows, n_codebooks, head_vocab), dtype=torch.float32, device=hidden.device)
    ctx = _lib.context(hidden.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_heads_cfg(ctx.handle, C.c_void_p(model), _lib.ptr(hidden), hidden.stride(0), R, float(cfg_scale),
                                       _lib.ptr(logits), _lib.stream_ptr(hidden.device)))
    return logits


@heads_cfg.register_fake
def _(model, hidden, cfg_scale, n_codebooks, head_vocab):
    R = hidden.shape[0]
    return hidden.new_empty((R // 2 if cfg_scale != 1.0 else R, n_codebooks, head_vocab), dtype=torch.float32)


@custom_op("zonos_b200::sample_update", mutates_args=())
def sample_update(logits: torch.Tensor, window: Optional[torch.Tensor], q: Optional[torch.Tensor], params: List[float], top_k: int,
                  penalty_window: int, seed: int, draw_index: int, apply_logit_bias: bool) -> torch.Tensor:
    """fp32 [B,Q,V] -> int64 [B,Q]: logit bias, repetition penalty over `window`, temperature, top-p/top-k/min-p or the
    unified sampler, Gumbel-max race against `q` (explicit Exp(1) draws) or a Philox stream keyed by (seed, draw_index).
    params = [temperature, top_p, min_p, linear, conf, quad, repetition_penalty]."""
    _cuda(logits, "sample_update")
    B, Q, V = logits.shape
    sp = _lib.zb_sampling()
    sp.temperature, sp.top_p, sp.min_p, sp.linear, sp.conf, sp.quad, sp.repetition_penalty = params
    sp.top_k, sp.repetition_penalty_window = int(top_k), int(penalty_window)
    wsb = wsq = W = 0
    if window is not None:
        wsb, wsq, W = window.stride(0), window.stride(1), window.shape[2]
    tokens = torch.empty((B, Q), dtype=torch.int64, device=logits.device)
    ctx = _lib.context(logits.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_sample_from_logits(ctx.handle, C.byref(sp), _lib.ptr(logits), B, Q, V, _lib.ptr(window), wsb, wsq, W,
                                                _lib.ptr(q), seed, draw_index, int(apply_logit_bias), _lib.ptr(tokens),
                                                _lib.stream_ptr(logits.device)))
    return tokens


@sample_update.register_fake
def _(logits, window, q, params, top_k, penalty_window, seed, draw_index, apply_logit_bias):
    return logits.new_empty(logits.shape[:2], dtype=torch.int64)


@custom_op("zonos_b200::decode_step", mutates_args=("delayed", "lengths"))
def decode_step(session: int, n_steps: int, delayed: torch.Tensor, lengths: torch.Tensor) -> None:
    """Enqueue `n_steps` whole decode steps of a live generate session (zb_generate_begin): each one embeds the last
    frame, runs the backbone over the paged cache, applies heads + CFG, samples and writes the next column of `delayed`.
    Steps enqueued after the device-side stop flag was raised are no-ops."""
    _cuda(delayed, "decode_step")
    ctx = _lib.context(delayed.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_generate_steps(C.c_void_p(session), n_steps, _lib.stream_ptr(delayed.device)))


@custom_op("zonos_b200::dac_decode", mutates_args=())
def dac_decode(dac: int, codes: torch.Tensor, upsample: int) -> torch.Tensor:
    """int64 [B,Q,T] -> fp32 [B,1,upsample*T] (DAC 44.1 kHz decoder: RVQ lookup, conv / Snake / transposed-conv stack)."""
    _cuda(codes, "dac_decode")
    B, _, T = codes.shape
    wav = torch.empty((B, 1, upsample * T), dtype=torch.float32, device=codes.device)
    if T == 0:
        return wav
    ctx = _lib.context(codes.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_dac_decode(ctx.handle, C.c_void_p(dac), _lib.ptr(codes), B, T, _lib.ptr(wav), _lib.stream_ptr(codes.device)))
    return wav


@dac_decode.register_fake
def _(dac, codes, upsample):
    B, _, T = codes.shape
    return codes.new_empty((B, 1, upsample * T), dtype=torch.float32)


@custom_op("zonos_b200::dac_encode", mutates_args=())
def dac_encode(wav: torch.Tensor, weights: List[torch.Tensor], n_codebooks: int) -> torch.Tensor:
    """fp32 [B,1,L] at 44.1 kHz (L a multiple of 512) -> int64 [B,Q,L/512]: DAC encoder convs + residual vector quantiser in
    fp32 (zonos/autoencoder.py:104-117).  `weights`: the fp32 tensors in the order of `autoencoder.dac_encoder_tensor_order`."""
    _cuda(wav, "dac_encode")
    B, _, L = wav.shape
    wav = wav.to(torch.float32).contiguous()
    codes = torch.empty((B, n_codebooks, L // 512), dtype=torch.int64, device=wav.device)
    ctx = _lib.context(wav.device)
    arr = (C.c_void_p * len(weights))(*[t.data_ptr() for t in weights])
    d = _lib.zb_dac_enc_desc()
    d.n_codebooks, d.codebook_size, d.codebook_dim, d.latent_dim, d.hidden, d.n_blocks = n_codebooks, 1024, 8, 1024, 64, 4
    for i, s_ in enumerate((2, 4, 8, 8)):
        d.strides[i] = s_
    d.tensors, d.n_tensors = arr, len(weights)
    nbytes = int(ctx.lib.zb_dac_encode_workspace_bytes(B, L))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=wav.device)
    with ctx.lock:
        ctx.check(ctx.lib.zb_dac_encode(ctx.handle, C.byref(d), _lib.ptr(wav), B, L, _lib.ptr(codes), _lib.ptr(ws), nbytes, _lib.stream_ptr(wav.device)))
    return codes


@dac_encode.register_fake
def _(wav, weights, n_codebooks):
    B, _, L = wav.shape
    return wav.new_empty((B, n_codebooks, L // 512), dtype=torch.int64)


OPS = ("embed_codes", "backbone_forward", "heads_cfg", "sample_update", "decode_step", "dac_decode", "dac_encode")
