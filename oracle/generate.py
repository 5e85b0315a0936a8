"""Oracle: the autoregressive generate loop (zonos/model.py:354-548), CPU.

Composes the other oracle pieces exactly in the reference's order:
prefill (generation_utils.py:236-244) -> first sample without penalty/bias
(model.py:422-423) -> hot loop (model.py:467-509) -> finalize (model.py:511-539).
The Exp(1) draws come from torch's global CPU generator in the same order as
the reference (one [B,Q,V] tensor per sample call), or from `q_stream`
(fp32 [n_calls,B,Q,V]) when given - that is what the CUDA parity tests feed to
both sides.
"""
import numpy as np
import torch

from .codebook import apply_delay_pattern, revert_delay_pattern
from .sampling import eos_state_update, make_logit_bias, sample_from_logits, should_sync_check
from .transformer import TransformerOracle


def finalize_codes(delayed: torch.Tensor, offset: int, Q: int = 9, eos: int = 1024):
    """model.py:511-539: revert delay, batch-global EOS boundary scan, sanitise, slice, clamp."""
    out = torch.from_numpy(revert_delay_pattern(delayed.numpy()))
    valid = offset - Q
    window = min(50, valid // 4)
    start = max(0, valid - window)
    for pos in range(start, valid):
        if int((out[:, :, pos] == eos).sum()) >= Q // 2:
            valid = pos
            break
    out = torch.where(out > 1024, torch.full_like(out, 512), out)   # invalid -> 512 (mask token included)
    out = torch.where(out == eos, torch.zeros_like(out), out)       # eos -> 0
    return out[..., :valid].clamp(0, 1023)


@torch.inference_mode()
def generate(model: TransformerOracle, prefix_conditioning: torch.Tensor, audio_prefix_codes=None,
             max_new_tokens: int = 86 * 30, cfg_scale: float = 2.0, batch_size: int = 1,
             sampling_params: dict | None = None, q_stream: torch.Tensor | None = None,
             trace: dict | None = None, callback=None):
    """Returns int64 [B, Q, valid_len].  `trace`, when a dict, receives
    delayed (final delayed codes), offset, logits (list per sample call) and steps."""
    sampling_params = dict(min_p=0.1) if sampling_params is None else sampling_params
    assert cfg_scale != 1                                            # model.py:399
    d = model.d
    Q, B = d.n_codebooks, batch_size
    eos, mask, unknown = 1024, 1025, -1
    P = 0 if audio_prefix_codes is None else audio_prefix_codes.shape[2]
    Lc = prefix_conditioning.shape[1]
    audio_len = P + max_new_tokens
    st = model.allocate(2 * B, Lc + audio_len + Q)                   # model.py:410-413
    codes = torch.full((B, Q, audio_len), unknown, dtype=torch.int64)
    if audio_prefix_codes is not None:
        codes[..., :P] = audio_prefix_codes
    delayed = torch.from_numpy(apply_delay_pattern(codes.numpy(), mask))
    ncall = [0]

    def draw():
        if q_stream is None:
            return None
        ncall[0] += 1
        return q_stream[ncall[0] - 1]

    # ---- prefill (generation_utils.py:236-244; `repeat` where the reference's
    # `expand` only works for B == 1 - SURVEY.md 2.3 quirk 12) -----------------
    ids = delayed[..., :P + 1].repeat(2, 1, 1)
    hidden = torch.cat([prefix_conditioning.to(model.dtype), model.embed(ids)], dim=1)
    logits = model.logits(hidden, st, cfg_scale)
    if trace is not None:
        trace["logits"] = [logits.clone()]
    tok = sample_from_logits(logits, q=draw(), **sampling_params)    # no penalty, no bias (model.py:423)
    offset = P + 1
    frame = delayed[..., offset]
    delayed[..., offset] = torch.where(frame == unknown, tok, frame)
    first_frame = delayed[..., offset:offset + 1]                    # what model.py:508 hands the callback
    st.seqlen_offset += Lc + P + 1
    st.lengths += Lc + P + 1

    bias = make_logit_bias(B, Q, d.head_vocab, eos)
    stopping = torch.zeros(B, dtype=torch.bool)
    max_steps = delayed.shape[2] - offset
    remaining = torch.full((B,), max_steps, dtype=torch.int64)
    ctx_len = min(max_new_tokens, 100)
    steps_done = 0
    for step_idx in range(max_steps):
        offset += 1
        if offset >= delayed.shape[2]:
            break
        ids = delayed[..., offset - 1:offset].repeat(2, 1, 1)        # generation_utils.py:191-193
        logits = model.logits(model.embed(ids), st, cfg_scale) + bias
        if trace is not None:
            trace["logits"].append(logits.clone())
        window = delayed[..., max(0, offset - ctx_len):offset]
        tok = sample_from_logits(logits, q=draw(), generated_tokens=window, **sampling_params)
        tok, remaining, stopping = eos_state_update(tok, remaining, stopping, Q, eos, mask)
        frame = delayed[..., offset]
        delayed[..., offset] = torch.where(frame == unknown, tok, frame)
        st.seqlen_offset += 1                                        # tensor_ops.py:85-87
        st.lengths += 1
        remaining = remaining - 1
        steps_done = step_idx + 1
        if should_sync_check(step_idx, step_idx + 1, B) and bool((remaining <= 0).all()):
            break
        if callback is not None and not callback(first_frame, steps_done, max_steps):
            break
    if trace is not None:
        trace.update(delayed=delayed.clone(), offset=offset, steps=steps_done)
    return finalize_codes(delayed, offset, Q, eos)
