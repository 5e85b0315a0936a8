"""Codebook delay pattern (integer).  Same contract as zonos/codebook_pattern.py:5-61; runs once per
generate() call on whatever device the codes live on."""
import torch


def apply_delay_pattern(codes: torch.Tensor, mask_token: int) -> torch.Tensor:
    """[B,Q,T] -> [B,Q,T+Q]: codebook k is shifted right by k+1, gaps hold `mask_token`."""
    B, Q, T = codes.shape
    out = codes.new_full((B, Q, T + Q), mask_token)
    for k in range(Q):
        out[:, k, k + 1:k + 1 + T] = codes[:, k]
    return out


def revert_delay_pattern(codes: torch.Tensor) -> torch.Tensor:
    """[B,Q,T+Q] -> [B,Q,T]."""
    _, Q, L = codes.shape
    return torch.stack([codes[:, k, k + 1:L - Q + k + 1] for k in range(Q)], dim=1)
