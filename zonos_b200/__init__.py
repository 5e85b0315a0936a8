"""zonos_b200: the Zonos-v0.1 inference hot path on B200 (sm_100a), drop-in behind the reference's API.

    from zonos_b200 import Zonos
    model = Zonos.from_pretrained("Zyphra/Zonos-v0.1-transformer", device="cuda")
    cond = model.prepare_conditioning(cond_dict, cfg_scale=2.0)
    codes = model.generate(cond)
    wav = model.autoencoder.decode(codes)

All arithmetic of that path runs in `libzonos_b200.so` (hand-written CUDA behind the C ABI of
`include/zonos_b200.h`); importing this package never falls back to a CPU or library implementation.
"""
from .autoencoder import DACAutoencoder
from .backbone import BACKBONES, B200ZonosBackbone
from .codebook_pattern import apply_delay_pattern, revert_delay_pattern
from .config import (BackboneConfig, InferenceParams, PrefixConditionerConfig, ZonosConfig, hybrid_config_dict,
                     transformer_config_dict)
from .model import Zonos
from .sampling import sample_from_logits

__all__ = ["Zonos", "ZonosConfig", "BackboneConfig", "PrefixConditionerConfig", "InferenceParams", "BACKBONES",
           "B200ZonosBackbone", "DACAutoencoder", "sample_from_logits", "apply_delay_pattern", "revert_delay_pattern",
           "transformer_config_dict", "hybrid_config_dict"]
