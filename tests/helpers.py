"""Shared test helpers (deterministic inputs, model builders).  The oracle is only ever the CHECKER here."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SAMPLER_CASES = [          # must stay in sync with tests/golden/make_golden.py
    dict(min_p=0.1),
    dict(min_p=0.1, temperature=0.7),
    dict(linear=0.5, conf=0.4, quad=0.0),
    dict(linear=0.8, conf=0.2, quad=0.3, min_p=0.05),
    dict(top_p=0.9),
    dict(top_k=50),
    dict(top_p=0.8, top_k=20, min_p=0.02, temperature=1.3),
    dict(temperature=0.0),
    dict(min_p=0.1, repetition_penalty=1.0),
    dict(min_p=0.1, repetition_penalty=2.0, repetition_penalty_window=5),
]


def sampler_inputs(case_idx: int, B=2, Q=9, V=1025, W=7):
    g = torch.Generator().manual_seed(1000 + case_idx)
    scale = (0.5, 2.0, 6.0)[case_idx % 3]
    logits = torch.randn(B, Q, V, generator=g) * scale
    logits[:, 1:, 1024] = -math.inf
    if case_idx % 4 == 1:
        logits[0, 0, 100:400] = -math.inf
    window = torch.randint(0, 1026, (B, Q, W), generator=g)
    window[0, 0, -1] = window[0, 0, -2]
    window[1, 2, -1] = 1025
    q = torch.empty(B, Q, V).exponential_(1, generator=g)
    return logits, window, q


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def q_stream_from_seed(seed: int, n_calls: int, B: int, Q: int = 9, V: int = 1025) -> torch.Tensor:
    """The Exp(1) draws the reference makes under torch.manual_seed(seed): one [B,Q,V] tensor per sample call."""
    torch.manual_seed(seed)
    return torch.stack([torch.empty(B, Q, V).exponential_(1) for _ in range(n_calls)])


def oracle_dims(dims: dict):
    from oracle.transformer import BackboneDims
    return BackboneDims(d_model=dims["d_model"], n_layer=dims["n_layer"], n_heads=dims["n_heads"],
                        n_heads_kv=dims["n_heads_kv"], d_ff=dims["d_ff"])


def build_b200_model(dims: dict, weights: dict, device, dac_weights=None):
    from zonos_b200 import DACAutoencoder, Zonos, ZonosConfig, transformer_config_dict
    ae = DACAutoencoder(dac_weights, device=device) if dac_weights is not None else None
    m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**dims)), autoencoder=ae).to(device, torch.bfloat16)
    m.load_state_dict(weights)
    return m.eval().requires_grad_(False)


def eos_boosted(weights: dict, boost: float) -> dict:
    """Same tweak as make_golden.py: make codebook-0 EOS likely so the EOS state machine is exercised."""
    w = dict(weights)
    hw = w["fused_heads.weight"].clone()
    hw[1024] = (boost * w["backbone.norm_f.bias"].float()).to(hw.dtype)
    w["fused_heads.weight"] = hw
    return w


FP8_KEYS = ("mixer.in_proj.weight", "mixer.out_proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")


def fp8_dequantised(weights: dict) -> dict:
    """The decode matrices after the FP8 mode's quantiser (zonos_b200/csrc/decode.cu: quant_e4m3_kernel): e4m3 with one
    power-of-two scale per weight row, the row's largest |w| / scale in [1, 2).  The dequantised value q * scale is exactly
    a bf16 number, so the FP8 kernel can be compared with the bf16 kernel run on THESE weights."""
    out = dict(weights)
    for k, v in weights.items():
        if k.endswith(FP8_KEYS) or k == "fused_heads.weight":
            f = v.float()
            amax = f.abs().amax(dim=1, keepdim=True)
            _, e = torch.frexp(amax)                                        # amax = m * 2^e, m in [0.5, 1)
            sc = torch.where(amax > 0, torch.ldexp(torch.ones_like(amax), e - 1), torch.ones_like(amax))
            out[k] = ((f / sc).clamp(-1.875, 1.875).to(torch.float8_e4m3fn).float() * sc).to(v.dtype)   # never rounds up to 2.0: idempotent
    return out
