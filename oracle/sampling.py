"""Oracle: sampler chain (fp32) + EOS state machine (integer).  CPU torch.

Restates zonos/sampling.py:130-231 and zonos/model.py:433-437,476-500 with the
Exp(1) draws `q` passed in explicitly, so a CUDA run with the same `q` must
pick the same token ("bit-exact given identical uniform draws").
Floating point stays in torch CPU fp32 with the reference's op order (each
stage renormalises with a divide); integer work is exact.
"""
import math
import torch


def repetition_penalty(logits, window_tokens, penalty: float, window: int):
    """sampling.py:159-163.  logits fp32 [B,Q,V]; window_tokens int64 [B,Q,W'].

    Only the last `window` columns count; ids clamp to V-1; a token that occurs
    n times gets penalty**n (scatter_reduce prod); non-positive logits are
    multiplied, positive ones divided.
    """
    V = logits.shape[-1]
    toks = window_tokens[..., -window:].clamp(max=V - 1).to(torch.int64)
    factors = torch.ones_like(logits)
    for j in range(toks.shape[-1]):  # sequential products == reduce="prod"
        idx = toks[..., j:j + 1]
        factors.scatter_(2, idx, factors.gather(2, idx) * penalty)
    return torch.where(logits <= 0, logits * factors, logits / factors)


def unified(probs, linear: float, conf: float, quad: float):
    """sampling.py:60-63 (NovelAI unified sampler)."""
    logp = torch.log(probs.clamp_min(1e-20))
    entropy = -(probs * logp).sum(-1, keepdim=True)
    raw = logp * (linear + entropy * conf) - logp ** 2 * quad
    return torch.softmax(raw, dim=-1)


def top_p(probs, p: float):
    """sampling.py:93-98: stable descending sort, exclusive-cumsum > p dropped."""
    srt, idx = torch.sort(probs, dim=-1, descending=True, stable=True)
    csum = torch.cumsum(srt, dim=-1)
    keep = ~((csum - srt) > p)
    srt = srt * keep.float()
    out = torch.zeros_like(probs).scatter(-1, idx, srt)
    return out / out.sum(-1, keepdim=True)


def top_k(probs, k: int):
    """sampling.py:77-80: keep everything >= the k-th largest value."""
    v, _ = torch.topk(probs, min(k, probs.shape[-1]))
    pivot = v[..., -1:]
    out = torch.where(probs < pivot, torch.zeros_like(probs), probs)
    return out / out.sum(-1, keepdim=True)


def min_p(probs, m: float):
    """sampling.py:123-126: drop p < m * max(p)."""
    top = probs.max(-1, keepdim=True).values
    out = probs.masked_fill(probs < (m * top), 0.0)
    return out / out.sum(-1, keepdim=True)


def final_probs(logits, temperature=1.0, top_p_=0.0, top_k_=0, min_p_=0.0,
                linear=0.0, conf=0.0, quad=0.0, window_tokens=None,
                repetition_penalty_=3.0, repetition_penalty_window=2):
    """Probabilities right before the multinomial draw (sampling.py:213-227)."""
    if repetition_penalty_ != 1.0 and window_tokens is not None:
        logits = repetition_penalty(logits, window_tokens, repetition_penalty_, repetition_penalty_window)
    probs = torch.softmax(logits / temperature, dim=-1)
    if linear > 0.0:
        probs = unified(probs, linear, conf, quad)
    if top_p_ > 0:
        probs = top_p(probs, top_p_)
    if top_k_ > 0:
        probs = top_k(probs, top_k_)
    if min_p_ > 0:
        probs = min_p(probs, min_p_)
    return probs


def sample_from_logits(logits, q=None, temperature=1.0, top_p=0.0, top_k=0, min_p=0.0,
                       linear=0.0, conf=0.0, quad=0.0, generated_tokens=None,
                       repetition_penalty=3.0, repetition_penalty_window=2,
                       return_margin=False):
    """sampling.py:166-231.  Returns int64 [B,Q] (the reference returns [B,Q,1]).

    `q` fp32 [B,Q,V] are the Exp(1) draws of sampling.py:29; when None they are
    drawn from torch's global CPU generator exactly like the reference does
    (`empty_like(p).exponential_(1)`), so seeding reproduces its stream.
    temperature <= 0 -> argmax of the (penalised) logits (:229).
    `return_margin` adds the ratio second-best / best score, which tests use to
    tell real mismatches from floating-point near-ties.
    """
    if repetition_penalty != 1.0 and generated_tokens is not None:
        logits = globals()["repetition_penalty"](logits, generated_tokens, repetition_penalty,
                                                  repetition_penalty_window)
    if temperature > 0:
        probs = final_probs(logits, temperature, top_p, top_k, min_p, linear, conf, quad,
                            None, 1.0, repetition_penalty_window)
        if q is None:
            q = torch.empty_like(probs).exponential_(1)
        score = probs / q
    else:
        score = logits
    tok = torch.argmax(score, dim=-1)
    if return_margin:
        top2 = torch.topk(score, 2, dim=-1).values
        margin = torch.where(top2[..., 0] > 0, top2[..., 1] / top2[..., 0], torch.ones_like(top2[..., 0]))
        return tok, margin
    return tok


def make_logit_bias(B: int, Q: int, V: int, eos: int = 1024):
    """model.py:433-437: EOS forbidden on codebooks 1.., minus ln2 on codebook 0."""
    bias = torch.zeros(B, Q, V)
    bias[:, 1:, eos] = -math.inf
    bias[:, 0, eos] -= torch.log(torch.tensor(2.0))
    return bias


def eos_state_update(tok, remaining, stopping, Q: int = 9, eos: int = 1024, mask: int = 1025):
    """model.py:483-497 + utilities/tensor_ops.py:193-211, on int64 CPU tensors.

    tok [B,Q] sampled tokens; remaining int64 [B]; stopping bool [B].
    Returns (masked tokens, remaining, stopping) - `remaining` is NOT yet
    decremented (that is fused_parameter_updates, tensor_ops.py:87).
    """
    eos0 = tok[:, 0] == eos
    remaining = torch.where(eos0, torch.minimum(remaining, torch.full_like(remaining, Q)), remaining)
    stopping = stopping | eos0
    eos_idx = (Q - remaining).clamp(max=Q - 1)  # model.py:490-491 clamp_(max=8)
    cb = torch.arange(Q).unsqueeze(0)
    before = stopping.unsqueeze(1) & (cb < eos_idx.unsqueeze(1))
    at = stopping.unsqueeze(1) & (cb == eos_idx.unsqueeze(1))
    out = torch.where(before, torch.full_like(tok, mask), torch.where(at, torch.full_like(tok, eos), tok))
    return out, remaining, stopping


def should_sync_check(step_idx: int, cpu_step_counter: int, batch: int) -> bool:
    """utilities/tensor_ops.py:90-103: on which steps the reference looks at
    `(remaining_steps <= 0).all()`."""
    if step_idx % 16 == 15:
        return True
    if step_idx % 8 == 7:
        return max(0, batch * 10 - cpu_step_counter) < 5
    return False
