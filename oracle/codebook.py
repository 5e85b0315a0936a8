"""Oracle: codebook delay pattern (integer, bit-exact).  numpy only.

Follows zonos/codebook_pattern.py:5-32 (apply) and :35-61 (revert).
"""
import numpy as np


def apply_delay_pattern(codes: np.ndarray, mask_token: int) -> np.ndarray:
    """codes int64 [B, Q, T] -> [B, Q, T+Q].

    Reference (codebook_pattern.py:31-32): right-pad every row by Q mask tokens,
    then roll codebook k right by k+1.  Because the pad is Q >= k+1 long, the
    roll only wraps mask tokens, so row k is  (k+1 masks) ++ codes ++ (Q-k-1 masks).
    """
    codes = np.asarray(codes)
    B, Q, T = codes.shape
    out = np.full((B, Q, T + Q), mask_token, dtype=codes.dtype)
    for k in range(Q):
        out[:, k, k + 1:k + 1 + T] = codes[:, k]
    return out


def revert_delay_pattern(delayed: np.ndarray) -> np.ndarray:
    """delayed int64 [B, Q, T+Q] -> [B, Q, T]  (codebook_pattern.py:60-61)."""
    delayed = np.asarray(delayed)
    B, Q, L = delayed.shape
    return np.stack([delayed[:, k, k + 1:L - Q + k + 1] for k in range(Q)], axis=1)
