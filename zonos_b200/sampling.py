"""Sampler front-end with the reference's signature (`zonos/sampling.py:166-231`), executed by one CUDA launch."""
import ctypes as C

import torch

from . import _lib, ops  # noqa: F401


def sampling_struct(temperature=1.0, top_p=0.0, top_k=0, min_p=0.0, linear=0.0, conf=0.0, quad=0.0,
                    repetition_penalty=3.0, repetition_penalty_window=2) -> _lib.zb_sampling:
    s = _lib.zb_sampling()
    s.temperature, s.top_p, s.min_p, s.linear, s.conf, s.quad = temperature, top_p, min_p, linear, conf, quad
    s.repetition_penalty, s.top_k, s.repetition_penalty_window = repetition_penalty, int(top_k), int(repetition_penalty_window)
    return s


def sample_from_logits(logits: torch.Tensor, temperature: float = 1.0, top_p: float = 0.0, top_k: int = 0, min_p: float = 0.0,
                       linear: float = 0.0, conf: float = 0.0, quad: float = 0.0, generated_tokens: torch.Tensor | None = None,
                       repetition_penalty: float = 3.0, repetition_penalty_window: int = 2, *, q: torch.Tensor | None = None,
                       seed: int | None = None, draw_index: int = 0, apply_logit_bias: bool = False) -> torch.Tensor:
    """logits fp32 [B,Q,V] -> int64 [B,Q,1].  Keyword-only extras: `q` = explicit Exp(1) draws [B,Q,V] (what
    `torch.empty_like(p).exponential_(1)` is in zonos/sampling.py:29); otherwise a Philox stream keyed by `seed`
    (default: one draw from torch's global generator, so `torch.manual_seed` makes runs reproducible)."""
    if logits.device.type != "cuda":
        raise RuntimeError("zonos_b200.sample_from_logits runs on CUDA only")
    assert logits.dim() == 3
    B, Q, V = logits.shape
    lg = logits.contiguous().float()
    win = None
    if generated_tokens is not None:
        win = generated_tokens.to(torch.int64)
        if win.stride(-1) != 1:
            win = win.contiguous()
    if q is not None:
        q = q.contiguous().float()
        assert q.shape == lg.shape and q.is_cuda
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,)).item()) if q is None else 0
    tokens = torch.ops.zonos_b200.sample_update(lg, win, q, [float(temperature), float(top_p), float(min_p), float(linear), float(conf),
                                                              float(quad), float(repetition_penalty)], int(top_k),
                                                int(repetition_penalty_window), int(seed), int(draw_index), bool(apply_logit_bias))
    return tokens.unsqueeze(-1)
