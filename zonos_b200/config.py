"""Configuration schema of the path, field-compatible with the reference's `zonos/config.py:8-149`
(same names and defaults, so a reference `config.json` loads unchanged)."""
from dataclasses import dataclass, field
from typing import Literal

import torch


@dataclass
class InferenceParams:
    """Decode state handed to a backbone (zonos/config.py:8-52).  For the B200 backbone
    `key_value_memory_dict[layer] = (kv_pages_of_that_layer, page_table)` (paged cache) and the device
    tensor `lengths_per_sample` (int32 [R]) is the source of truth for positions."""
    max_seqlen: int
    max_batch_size: int
    seqlen_offset: int = 0
    batch_size_offset: int = 0
    key_value_memory_dict: dict = field(default_factory=dict)
    lengths_per_sample: torch.Tensor | None = None

    def reset(self, max_seqlen, max_batch_size):
        self.max_seqlen = max_seqlen
        self.max_batch_size = max_batch_size
        self.seqlen_offset = 0
        if self.lengths_per_sample is not None:
            self.lengths_per_sample.zero_()


@dataclass
class BackboneConfig:
    """zonos/config.py:55-84."""
    d_model: int = 1024
    d_intermediate: int = 0
    attn_mlp_d_intermediate: int = 0
    n_layer: int = 16
    ssm_cfg: dict = field(default_factory=dict)
    attn_layer_idx: list = field(default_factory=list)
    attn_cfg: dict = field(default_factory=dict)
    rms_norm: bool = False
    residual_in_fp32: bool = False
    norm_epsilon: float = 1e-5


@dataclass
class PrefixConditionerConfig:
    """zonos/config.py:87-102."""
    conditioners: list[dict]
    projection: Literal["none", "linear", "mlp"]


@dataclass
class ZonosConfig:
    """zonos/config.py:105-149."""
    backbone: BackboneConfig
    prefix_conditioner: PrefixConditionerConfig
    eos_token_id: int = 1024
    masked_token_id: int = 1025
    pad_vocab_to_multiple_of: int = 8
    codebook_dimension: int = 9

    @classmethod
    def from_dict(cls, d: dict) -> "ZonosConfig":
        d = {k: v for k, v in d.items() if not k.startswith("_")}          # "_comment" keys of the shipped configs/*.json
        backbone = BackboneConfig(**d.pop("backbone"))
        prefix = PrefixConditionerConfig(**d.pop("prefix_conditioner"))
        return cls(backbone, prefix, **d)


def transformer_config_dict(d_model=2048, n_layer=26, n_heads=16, n_heads_kv=4, d_ff=8192, conditioners=None) -> dict:
    """The Zonos-v0.1-transformer shape (D=2048, L=26, H=16, Hkv=4, F=8192) as a config dict."""
    return dict(
        backbone=dict(d_model=d_model, d_intermediate=0, attn_mlp_d_intermediate=d_ff, n_layer=n_layer, ssm_cfg={},
                      attn_layer_idx=list(range(n_layer)),
                      attn_cfg=dict(causal=True, num_heads=n_heads, num_heads_kv=n_heads_kv, rotary_emb_dim=128,
                                    qkv_proj_bias=False, out_proj_bias=False),
                      rms_norm=False, residual_in_fp32=False, norm_epsilon=1e-5),
        prefix_conditioner=dict(projection="linear", conditioners=conditioners if conditioners is not None else []),
        eos_token_id=1024, masked_token_id=1025)


def hybrid_config_dict(d_model=2048, n_layer=46, attn_layer_idx=(9, 18, 27, 36, 45), n_heads=16, n_heads_kv=4, d_ff=8192,
                       rms_norm=False, conditioners=None) -> dict:
    """Zonos-v0.1-hybrid shape.  The hybrid config.json is NOT in the reference tree; n_layer / attn_layer_idx here are the
    assumption stated in SURVEY.md 8(c) - everything stays parameterised by BackboneConfig (zonos/config.py:75-84)."""
    return dict(
        backbone=dict(d_model=d_model, d_intermediate=0, attn_mlp_d_intermediate=d_ff, n_layer=n_layer, ssm_cfg={"layer": "Mamba2"},
                      attn_layer_idx=list(attn_layer_idx),
                      attn_cfg=dict(causal=True, num_heads=n_heads, num_heads_kv=n_heads_kv, rotary_emb_dim=128,
                                    qkv_proj_bias=False, out_proj_bias=False),
                      rms_norm=rms_norm, residual_in_fp32=False, norm_epsilon=1e-5),
        prefix_conditioner=dict(projection="linear", conditioners=conditioners if conditioners is not None else []),
        eos_token_id=1024, masked_token_id=1025)
