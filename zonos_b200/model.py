"""`Zonos` with the reference's public surface (`zonos/model.py:43-548`), hot path on B200.

Drop-in methods: `from_pretrained`, `from_local`, `prepare_conditioning`, `generate`, `autoencoder.decode`,
plus `embed_codes`, `apply_heads`, `setup_cache`, `device`.  `generate` runs prefill, the whole autoregressive
loop (embed -> backbone -> heads/CFG -> sampler -> EOS/delay bookkeeping) and the early-exit test on the device
through `zb_generate_*`; the host only enqueues CUDA-graph replays and polls a flag.
"""
import os
import ctypes as C
import json
from typing import Callable

import torch
import torch.nn as nn

from . import _lib, ops  # noqa: F401  (ops registers torch.ops.zonos_b200.*)
from .autoencoder import DACAutoencoder
from .backbone import BACKBONES, B200ZonosBackbone
from .codebook_pattern import apply_delay_pattern, revert_delay_pattern
from .config import InferenceParams, ZonosConfig
from .sampling import sampling_struct

DEFAULT_BACKBONE_CLS = B200ZonosBackbone
POLL_EVERY = 16          # the reference looks at the stop flag every 16 (or 8) steps, utilities/tensor_ops.py:90-103


def find_multiple(n: int, k: int) -> int:
    """zonos/utilities/utils.py:6-29."""
    return n if k == 0 or n % k == 0 else n + k - (n % k)


class Zonos(nn.Module):
    def __init__(self, config: ZonosConfig, backbone_cls=DEFAULT_BACKBONE_CLS, autoencoder: DACAutoencoder | None = None,
                 num_codebooks: int | None = None):
        """`autoencoder=None` defers building the DAC front-end until first use (`model.autoencoder`), which then loads
        `descript/dac_44khz` like zonos/autoencoder.py:74; pass a ready `DACAutoencoder` to stay offline."""
        super().__init__()
        self.config = config
        dim = config.backbone.d_model
        self.eos_token_id = config.eos_token_id
        self.masked_token_id = config.masked_token_id
        self._autoencoder = autoencoder
        self.backbone = backbone_cls(config.backbone)
        if config.prefix_conditioner.conditioners:
            from .conditioning import PrefixConditioner
            self.prefix_conditioner = PrefixConditioner(config.prefix_conditioner, dim)
        else:
            self.prefix_conditioner = None
        n_q = num_codebooks or (autoencoder.num_codebooks if autoencoder is not None else config.codebook_dimension)
        self.num_codebooks = n_q
        vocab = find_multiple(1026, 8)                        # 1024 codes + EOS + MASK, padded to 1032 (model.py:79)
        self.embeddings = nn.ModuleList([nn.Embedding(vocab, dim) for _ in range(n_q)])
        self.fused_heads = nn.Linear(dim, n_q * 1025, bias=False)
        self._conditioning_cache = {}

    # ---------------------------------------------------------------- loading ------------------
    @property
    def autoencoder(self) -> DACAutoencoder:
        if self._autoencoder is None:
            self._autoencoder = DACAutoencoder(device=self.device)
        return self._autoencoder

    @autoencoder.setter
    def autoencoder(self, value):
        self._autoencoder = value

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @classmethod
    def from_pretrained(cls, repo_id: str, revision: str | None = None, device: str = "cuda", **kwargs) -> "Zonos":
        """zonos/model.py:103-126 (needs network or a local HF cache)."""
        from huggingface_hub import hf_hub_download
        config_path = hf_hub_download(repo_id=repo_id, filename="config.json", revision=revision)
        model_path = hf_hub_download(repo_id=repo_id, filename="model.safetensors", revision=revision)
        return cls.from_local(config_path, model_path, device, **kwargs)

    @classmethod
    def from_local(cls, config_path: str, model_path: str, device: str = "cuda", backbone: str | None = None,
                   autoencoder: DACAutoencoder | None = None) -> "Zonos":
        """zonos/model.py:128-176: config.json + model.safetensors -> bf16 model on `device`."""
        import safetensors
        config = ZonosConfig.from_dict(json.load(open(config_path)))
        backbone_cls = BACKBONES[backbone] if backbone else DEFAULT_BACKBONE_CLS
        model = cls(config, backbone_cls, autoencoder=autoencoder).to(device, torch.bfloat16)
        sd = model.state_dict()
        with safetensors.safe_open(model_path, framework="pt") as f:
            for k in f.keys():
                t = f.get_tensor(k)
                if k.startswith("embeddings.") and k.endswith(".weight") and k in sd and sd[k].shape[0] != t.shape[0] \
                        and sd[k].shape[1] == t.shape[1]:
                    padded = torch.zeros(sd[k].shape, dtype=t.dtype)          # 1026 -> 1032 rows (model.py:164-172)
                    padded[: t.shape[0]] = t
                    t = padded
                sd[k] = t
        model.load_state_dict(sd)
        return model

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """Checkpoints store `heads.{i}.weight`; fuse them row-wise into `fused_heads.weight` (model.py:208-223)."""
        if f"{prefix}heads.0.weight" in state_dict:
            ws, i = [], 0
            while f"{prefix}heads.{i}.weight" in state_dict:
                ws.append(state_dict.pop(f"{prefix}heads.{i}.weight"))
                i += 1
            state_dict[f"{prefix}fused_heads.weight"] = torch.cat(ws, dim=0)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    # ---------------------------------------------------------------- native handles -----------
    def _native_model(self):
        self.backbone.register_tables([e.weight for e in self.embeddings], self.fused_heads.weight, self.num_codebooks, 1025)
        return self.backbone.native_model()

    def _ctx(self):
        return _lib.context(self.device)

    # ---------------------------------------------------------------- plugin-level pieces ------
    def embed_codes(self, codes: torch.Tensor, repeat: int = 1) -> torch.Tensor:
        """zonos/model.py:179-192: int64 [B,Q,T] -> bf16 [B*repeat,T,D] (sum of the Q codebook embeddings)."""
        B, Q, T = codes.shape
        assert Q == self.num_codebooks
        codes = codes.to(self.device, torch.int64)
        return torch.ops.zonos_b200.embed_codes(self._native_model().value, codes, repeat, self.config.backbone.d_model)

    def apply_heads(self, hidden_states: torch.Tensor) -> torch.Tensor:
        """zonos/model.py:194-206: bf16 [R,T,D] -> [R,Q,T,1025] (returned in fp32; the reference casts right after)."""
        R, T, D = hidden_states.shape
        h = hidden_states.contiguous().view(R * T, D)
        logits = torch.ops.zonos_b200.heads_cfg(self._native_model().value, h, 1.0, self.num_codebooks, 1025)
        return logits.view(R, T, self.num_codebooks, 1025).transpose(1, 2)

    def _compute_logits(self, hidden_states: torch.Tensor, inference_params: InferenceParams, cfg_scale: float) -> torch.Tensor:
        """zonos/model.py:225-234: backbone -> last token -> heads -> fp32 -> CFG mix; [R,T,D] -> [B,Q,1025]."""
        last = self.backbone(hidden_states, inference_params, last_only=True)[:, 0]
        return torch.ops.zonos_b200.heads_cfg(self._native_model().value, last, float(cfg_scale), self.num_codebooks, 1025)

    def setup_cache(self, batch_size: int, max_seqlen: int, dtype: torch.dtype = torch.bfloat16) -> InferenceParams:
        """zonos/model.py:305-338."""
        max_seqlen = find_multiple(max_seqlen, 8)
        kv = self.backbone.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype)
        lengths = torch.zeros(batch_size, dtype=torch.int32, device=self.device)
        return InferenceParams(max_seqlen, batch_size, 0, 0, kv, lengths)

    # ---------------------------------------------------------------- conditioning --------------
    def prepare_conditioning(self, cond_dict: dict, uncond_dict: dict | None = None, use_cache: bool = False,
                             cfg_scale: float = 1.0) -> torch.Tensor:
        """zonos/model.py:237-265 -> bf16 [B or 2B, Lc, D]; torch code, runs once per utterance."""
        if self.prefix_conditioner is None:
            raise RuntimeError("this model was built without prefix conditioners")
        from .conditioning import prepare_conditioning_with_cache
        return prepare_conditioning_with_cache(self.prefix_conditioner, cond_dict, uncond_dict, use_cache, cfg_scale,
                                               self._conditioning_cache if use_cache else None)

    # ---------------------------------------------------------------- generate ------------------
    @torch.inference_mode()
    def generate(self, prefix_conditioning: torch.Tensor, audio_prefix_codes: torch.Tensor | None = None,
                 max_new_tokens: int = 86 * 30, cfg_scale: float = 2.0, batch_size: int = 1,
                 sampling_params: dict = dict(min_p=0.1), disable_torch_compile: bool = False,
                 callback: Callable[[torch.Tensor, int, int], bool] | None = None, *,
                 q_stream: torch.Tensor | None = None, seed: int | None = None, trace: dict | None = None) -> torch.Tensor:
        """Same contract as zonos/model.py:354-548: returns int64 [B, 9, valid_len] in [0, 1023].

        Keyword-only extras (not in the reference): `q_stream` fp32 [n_calls,B,Q,1025] explicit Exp(1) draws (parity
        tests), `seed` for the device Philox stream (default: drawn from torch's global generator, so
        `torch.manual_seed` reproduces runs), `trace` dict that receives delayed codes / offset / per-call logits.
        `disable_torch_compile` is accepted and ignored (nothing here is traced).
        """
        assert cfg_scale != 1, "TODO: add support for cfg_scale=1"          # model.py:399
        device = self.device
        Q, B = self.num_codebooks, batch_size
        P = 0 if audio_prefix_codes is None else audio_prefix_codes.shape[2]
        cond = prefix_conditioning.to(device, torch.bfloat16).contiguous()
        assert cond.shape[0] == 2 * B, "prefix_conditioning must hold the cond and uncond halves ([2*batch, Lc, D])"
        Lc = cond.shape[1]
        audio_len = P + max_new_tokens
        seq_len = Lc + audio_len + Q                                         # model.py:409
        _dbg = os.environ.get("ZB_HOST_TIMES")
        if _dbg:
            import time as _time
            torch.cuda.synchronize(device); _t = [_time.perf_counter()]
            def _mark():
                torch.cuda.synchronize(device); _t.append(_time.perf_counter())
        params = self.setup_cache(batch_size=2 * B, max_seqlen=seq_len)
        codes = torch.full((B, Q, audio_len), -1, dtype=torch.int64, device=device)
        if audio_prefix_codes is not None:
            codes[..., :P] = audio_prefix_codes.to(device)
        delayed = apply_delay_pattern(codes, self.masked_token_id).contiguous()
        T_delayed = delayed.shape[2]
        if _dbg: _mark()

        n_calls = T_delayed - P                                              # sample calls at most (1 + max_steps - 1)
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,)).item())
        logits_trace = None
        if trace is not None:
            logits_trace = torch.zeros((n_calls, B, Q, 1025), dtype=torch.float32, device=device)
        if q_stream is not None:
            q_stream = q_stream.to(device, torch.float32).contiguous()
            assert q_stream.shape[1:] == (B, Q, 1025)

        desc = _lib.zb_gen_desc()
        desc.B, desc.Q, desc.T_delayed, desc.prefix_audio_len, desc.cond_len, desc.max_new_tokens = B, Q, T_delayed, P, Lc, max_new_tokens
        desc.delayed, desc.prefix_conditioning, desc.cfg_scale = delayed.data_ptr(), cond.data_ptr(), float(cfg_scale)
        desc.sampling = sampling_struct(**sampling_params)
        desc.q_stream = q_stream.data_ptr() if q_stream is not None else None
        desc.q_calls = q_stream.shape[0] if q_stream is not None else 0
        desc.seed = seed
        desc.logits_trace = logits_trace.data_ptr() if logits_trace is not None else None
        desc.trace_calls = n_calls if logits_trace is not None else 0

        ctx = self._ctx()
        lib = ctx.lib
        cache = self.backbone._cache.desc(params.lengths_per_sample)
        stream = _lib.stream_ptr(device)
        gen = C.c_void_p()
        prog = _lib.zb_gen_progress()
        with ctx.lock:
            ctx.check(lib.zb_generate_begin(ctx.handle, self._native_model(), C.byref(cache), C.byref(desc), C.byref(gen), stream))
            if _dbg: _mark()
            try:
                max_steps = T_delayed - (P + 1)
                first_frame = delayed[..., P + 1:P + 2]
                if callback is None:
                    # Enqueue one chunk ahead of the chunk whose stop flag is inspected, so the GPU never waits for the
                    # host; graph replays issued after the device set its stop flag are no-ops.
                    events, enq = [], 0
                    cur = torch.cuda.current_stream(device)
                    while enq < max_steps:
                        n = min(POLL_EVERY, max_steps - enq)
                        torch.ops.zonos_b200.decode_step(gen.value, n, delayed, params.lengths_per_sample)
                        enq += n
                        ev = torch.cuda.Event()
                        ev.record(cur)
                        events.append(ev)
                        if len(events) >= 2:
                            events[-2].synchronize()
                            ctx.check(lib.zb_generate_peek(gen, C.byref(prog)))
                            if prog.done:
                                break
                else:
                    for step in range(max_steps):                           # reference cadence: one host visit per step
                        torch.ops.zonos_b200.decode_step(gen.value, 1, delayed, params.lengths_per_sample)
                        ctx.check(lib.zb_generate_poll(gen, C.byref(prog), stream))
                        if prog.done:
                            break
                        if not callback(first_frame, prog.steps, max_steps):
                            break
                ctx.check(lib.zb_generate_poll(gen, C.byref(prog), stream))
                if _dbg: _mark()
            finally:
                lib.zb_generate_end(gen)
        if _dbg: _mark()
        offset = int(prog.offset)
        if trace is not None:
            trace.update(delayed=delayed.clone(), offset=offset, steps=int(prog.steps), logits=logits_trace, seed=seed)
        out = self._finalize(delayed, offset)
        if _dbg:
            _mark()
            print("host times ms: setup %.1f begin+prefill %.1f steps %.1f end %.1f finalize %.1f" % tuple((b - a) * 1e3 for a, b in zip(_t[:-1], _t[1:])))
        return out


    # ---------------------------------------------------------------- streaming -----------
    @torch.inference_mode()
    def generate_stream(self, prefix_conditioning: torch.Tensor, audio_prefix_codes: torch.Tensor | None = None,
                        max_new_tokens: int = 86 * 30, cfg_scale: float = 2.0, batch_size: int = 1,
                        sampling_params: dict = dict(min_p=0.1), *, chunk_frames: int = 43, holdback_frames: int = 25,
                        seed: int | None = None):
        """Generator of (wav fp32 [B,1,512*n], codes int64 [B,9,n]) chunks: the same loop as `generate`, but finished frames
        are DAC-decoded while the loop keeps running (the reference can only decode after `generate` returns).

        A frame is final once the loop front is `holdback_frames` past it: 25 >= the decoder's receptive field (16-frame halo,
        so every emitted sample equals the one a full decode produces) and = the span the reference's EOS-boundary scan can
        still cut (model.py:513-528: the EOS frame is at most 9 + 16 frames behind the front when the loop stops -
        remaining_steps = 9, stop flag read every 16 / 8 steps).  Time to first audio = prefill +
        (chunk_frames + holdback_frames + 9) decode steps + one chunk decode
        (tests: test_generate_stream_with_eos_never_emits_a_cut_frame)."""
        assert cfg_scale != 1, "TODO: add support for cfg_scale=1"
        device = self.device
        Q, B = self.num_codebooks, batch_size
        P = 0 if audio_prefix_codes is None else audio_prefix_codes.shape[2]
        cond = prefix_conditioning.to(device, torch.bfloat16).contiguous()
        assert cond.shape[0] == 2 * B
        Lc = cond.shape[1]
        params = self.setup_cache(batch_size=2 * B, max_seqlen=Lc + P + max_new_tokens + Q)
        codes = torch.full((B, Q, P + max_new_tokens), -1, dtype=torch.int64, device=device)
        if audio_prefix_codes is not None:
            codes[..., :P] = audio_prefix_codes.to(device)
        delayed = apply_delay_pattern(codes, self.masked_token_id).contiguous()
        T_delayed = delayed.shape[2]
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,)).item())
        desc = _lib.zb_gen_desc()
        desc.B, desc.Q, desc.T_delayed, desc.prefix_audio_len, desc.cond_len, desc.max_new_tokens = B, Q, T_delayed, P, Lc, max_new_tokens
        desc.delayed, desc.prefix_conditioning, desc.cfg_scale = delayed.data_ptr(), cond.data_ptr(), float(cfg_scale)
        desc.sampling = sampling_struct(**sampling_params)
        desc.seed = seed
        ctx = self._ctx()
        lib = ctx.lib
        cache = self.backbone._cache.desc(params.lengths_per_sample)
        stream = _lib.stream_ptr(device)
        gen = C.c_void_p()
        prog = _lib.zb_gen_progress()
        halo = 16                                                   # frames of decoder context on each side of a chunk
        emitted = P                                                 # frames [0, emitted) are done (the prefix is not re-emitted)

        def emit(lo: int, hi: int, last: bool):
            ctx_lo = max(0, lo - halo)
            ctx_hi = hi if last else hi + halo
            part = revert_delay_pattern(delayed[..., :ctx_hi + Q + 1])[..., ctx_lo:ctx_hi]
            part = torch.where(part > 1024, torch.full_like(part, 512), part)
            part = torch.where(part == 1024, torch.zeros_like(part), part).clamp(0, 1023)
            wav = self.autoencoder.decode(part)
            return wav[..., 512 * (lo - ctx_lo):512 * (hi - ctx_lo)], part[..., lo - ctx_lo:hi - ctx_lo]

        # The context lock is held only around the native calls, never across a `yield`: a slow or abandoned consumer
        # must not block other threads of the device.  The session (its slab, the pinned scratch) lives until the
        # generator is exhausted or closed - use `contextlib.closing(model.generate_stream(...))` or exhaust it.
        with ctx.lock:
            ctx.check(lib.zb_generate_begin(ctx.handle, self._native_model(), C.byref(cache), C.byref(desc), C.byref(gen), stream))
        try:
            max_steps = T_delayed - (P + 1)
            enq = 0
            while True:
                piece = None
                with ctx.lock:
                    # run until the next chunk is final: front (= offset - Q complete frames) >= emitted + chunk + holdback
                    need_front = emitted + chunk_frames + holdback_frames
                    need_steps = min(max_steps, need_front + Q - (P + 1))
                    if need_steps > enq:
                        torch.ops.zonos_b200.decode_step(gen.value, need_steps - enq, delayed, params.lengths_per_sample)
                        enq = need_steps
                    ctx.check(lib.zb_generate_poll(gen, C.byref(prog), stream))
                    front = int(prog.offset) - Q
                    finished = bool(prog.done) or enq >= max_steps
                    if not finished:
                        hi = min(front - holdback_frames, emitted + chunk_frames)
                        if hi > emitted:
                            piece = emit(emitted, hi, last=False)
                            emitted = hi
                if finished:
                    break
                if piece is not None:
                    yield piece
            piece = None
            with ctx.lock:
                final = self._finalize(delayed, int(prog.offset))
                valid = final.shape[-1]
                if valid > emitted:
                    ctx_lo = max(0, emitted - halo)
                    wav = self.autoencoder.decode(final[..., ctx_lo:valid])
                    piece = (wav[..., 512 * (emitted - ctx_lo):], final[..., emitted:valid])
            if piece is not None:
                yield piece
        finally:
            with ctx.lock:
                lib.zb_generate_end(gen)

    def _finalize(self, delayed: torch.Tensor, offset: int) -> torch.Tensor:
        """zonos/model.py:511-539: revert the delay pattern, batch-global EOS-boundary scan over the last <=50
        positions, sanitise (mask -> 512, EOS -> 0), slice, clamp."""
        Q = self.num_codebooks
        out = revert_delay_pattern(delayed)
        valid = offset - Q
        window = min(50, valid // 4)
        start = max(0, valid - window)
        if start < valid:
            counts = (out[:, :, start:valid] == self.eos_token_id).sum(dim=(0, 1))      # one sync instead of <=50
            hit = torch.nonzero(counts >= Q // 2)
            if hit.numel() > 0:
                valid = start + int(hit[0])
        out = torch.where(out > 1024, torch.full_like(out, 512), out)
        out = torch.where(out == 1024, torch.zeros_like(out), out)
        return out[..., :valid].clamp(0, 1023)
