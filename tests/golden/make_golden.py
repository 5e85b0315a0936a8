"""Generate the golden fixtures in this directory by running the REFERENCE itself.

Run in the build container only (needs /root/reference, CPU is enough):
    python tests/golden/make_golden.py
It imports langfod/Zonos from /root/reference (text front-end deps stubbed, DAC
built from `transformers` with seeded random weights - SURVEY.md Appendix A),
feeds it the seeded synthetic weights of `zonos_b200.synthetic`, records its
outputs into small .npz files, and asserts on the spot that the CPU oracle
(`oracle/`) reproduces every one of them.  The fixtures travel to the GPU box;
/root/reference does not.
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_stub("phonemizer"); _stub("phonemizer.backend", EspeakBackend=object)
_stub("inflect", engine=lambda: types.SimpleNamespace(number_to_words=lambda *a, **k: ""))
_stub("kanjize", number2kanji=str)


class _D:
    def __init__(self, dict=None): pass
    def create(self): return self


_stub("sudachipy", Dictionary=_D, SplitMode=types.SimpleNamespace(A=0))

from transformers.models.dac import DacConfig, DacModel  # noqa: E402
import zonos.autoencoder as ref_ae  # noqa: E402

from zonos_b200.synthetic import TINY_DIMS, make_backbone_weights, make_conditioning, make_dac_weights  # noqa: E402


def _dac_init(self):
    self.dac = DacModel(DacConfig(sampling_rate=44100)).eval().requires_grad_(False)
    self.dac.load_state_dict(make_dac_weights(seed=1), strict=False)
    self.codebook_size, self.num_codebooks = self.dac.config.codebook_size, self.dac.quantizer.n_codebooks
    self.sampling_rate = self.dac.config.sampling_rate


ref_ae.DACAutoencoder.__init__ = _dac_init

import zonos.model as ref_model  # noqa: E402
import zonos.sampling as ref_sampling  # noqa: E402
from zonos.backbone import BACKBONES  # noqa: E402
from zonos.codebook_pattern import apply_delay_pattern as ref_apply, revert_delay_pattern as ref_revert  # noqa: E402
from zonos.config import ZonosConfig  # noqa: E402

from oracle import codebook as o_cb, dac as o_dac, generate as o_gen, sampling as o_samp  # noqa: E402
from oracle.transformer import BackboneDims, TransformerOracle  # noqa: E402

torch.set_grad_enabled(False)


def ref_config(dims):
    return dict(backbone=dict(d_model=dims["d_model"], d_intermediate=0, attn_mlp_d_intermediate=dims["d_ff"],
                              n_layer=dims["n_layer"], ssm_cfg={}, attn_layer_idx=list(range(dims["n_layer"])),
                              attn_cfg=dict(causal=True, num_heads=dims["n_heads"], num_heads_kv=dims["n_heads_kv"],
                                            rotary_emb_dim=128, qkv_proj_bias=False, out_proj_bias=False),
                              rms_norm=False, residual_in_fp32=False, norm_epsilon=1e-5),
                prefix_conditioner=dict(projection="linear", conditioners=[
                    dict(type="PassthroughConditioner", name="speaker", cond_dim=128, uncond_type="learned",
                         projection="linear")]),
                eos_token_id=1024, masked_token_id=1025)


def build_reference(dims, weights):
    m = ref_model.Zonos(ZonosConfig.from_dict(ref_config(dims)), BACKBONES["torch"]).to("cpu", torch.bfloat16)
    sd = m.state_dict()
    sd.update(weights)
    m.load_state_dict(sd)
    return m.eval().requires_grad_(False)


# ---------------------------------------------------------------- codebook pattern
def golden_codebook():
    kat = np.array([[[1, 2, 3], [4, 5, 6], [7, 8, 9]]], dtype=np.int64)        # codebook_pattern.py:26-29
    M = 1025
    want = np.array([[[M, 1, 2, 3, M, M], [M, M, 4, 5, 6, M], [M, M, M, 7, 8, 9]]], dtype=np.int64)
    got = ref_apply(torch.from_numpy(kat), M).numpy()
    assert (got == want).all()
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(-1, 1026, (3, 9, 37), generator=g)
    delayed = ref_apply(codes, M)
    back = ref_revert(delayed)
    assert (back == codes).all()
    assert (o_cb.apply_delay_pattern(kat, M) == want).all()
    assert (o_cb.apply_delay_pattern(codes.numpy(), M) == delayed.numpy()).all()
    assert (o_cb.revert_delay_pattern(delayed.numpy()) == codes.numpy()).all()
    np.savez_compressed(os.path.join(HERE, "codebook_pattern.npz"), kat_in=kat, kat_out=want,
                        codes=codes.numpy(), delayed=delayed.numpy())


# ---------------------------------------------------------------- sampler
SAMPLER_CASES = [
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
    """Deterministic inputs shared by this script and tests/ (regenerated, not stored)."""
    g = torch.Generator().manual_seed(1000 + case_idx)
    scale = (0.5, 2.0, 6.0)[case_idx % 3]
    logits = torch.randn(B, Q, V, generator=g) * scale
    logits[:, 1:, 1024] = -math.inf                          # what logit_bias does (model.py:434)
    if case_idx % 4 == 1:
        logits[0, 0, 100:400] = -math.inf
    window = torch.randint(0, 1026, (B, Q, W), generator=g)
    window[0, 0, -1] = window[0, 0, -2]                      # duplicate -> factor penalty**2
    window[1, 2, -1] = 1025                                  # mask token clamps to index 1024
    q = torch.empty(B, Q, V).exponential_(1, generator=g)
    return logits, window, q


def golden_sampler():
    toks, margins = [], []
    for i, params in enumerate(SAMPLER_CASES):
        logits, window, q = sampler_inputs(i)
        # the reference draws q from the global generator: make it draw OUR q by seeding and replaying
        seed = 7000 + i
        torch.manual_seed(seed)
        q_ref = torch.empty_like(logits).exponential_(1)
        torch.manual_seed(seed)
        ref = ref_sampling.sample_from_logits(logits.clone(), generated_tokens=window, **params).squeeze(-1)
        mine, margin = o_samp.sample_from_logits(logits.clone(), q=q_ref, generated_tokens=window,
                                                 return_margin=True, **params)
        assert (ref == mine).all(), (i, params)
        toks.append(ref.numpy()); margins.append(margin.numpy())
        # and with the stored-q variant used by the CUDA tests
        mine2 = o_samp.sample_from_logits(logits.clone(), q=q, generated_tokens=window, **params)
        toks.append(mine2.numpy())
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), tokens=np.stack(toks), margins=np.stack(margins),
                        seeds=np.array([7000 + i for i in range(len(SAMPLER_CASES))]))


# ---------------------------------------------------------------- generate (tiny transformer)
class LogitTap:
    """Record what the reference hands to sample_from_logits (model.py:423,481)."""
    def __init__(self):
        self.logits = []
        self.orig = ref_model.sample_from_logits

    def __enter__(self):
        def tapped(logits, **kw):
            self.logits.append(logits.clone())
            return self.orig(logits, **kw)
        ref_model.sample_from_logits = tapped
        return self

    def __exit__(self, *a):
        ref_model.sample_from_logits = self.orig


def _prefill_repeat(prefix_hidden_states, input_ids, inference_params, cfg_scale, embed_codes_fn, compute_logits_fn):
    # harness patch for B > 1: `repeat` where the reference `expand`s
    if cfg_scale != 1.0:
        input_ids = input_ids.repeat(prefix_hidden_states.shape[0] // input_ids.shape[0], 1, 1)
    hidden = torch.cat([prefix_hidden_states, embed_codes_fn(input_ids)], dim=1)
    return compute_logits_fn(hidden, inference_params, cfg_scale)


def run_generate_case(name, dims, weights, B, Lc, N, P, seed, sampling_params, eos_boost=0.0):
    w = dict(weights)
    if eos_boost:
        hw = w["fused_heads.weight"].clone()
        hw[1024] = (eos_boost * w["backbone.norm_f.bias"].float()).to(hw.dtype)   # make codebook-0 EOS likely
        w["fused_heads.weight"] = hw
    ref = build_reference(dims, w)
    cond = make_conditioning(2 * B, Lc, dims["d_model"], seed=1234 + B)
    prefix = None
    if P:
        prefix = torch.randint(0, 1024, (B, 9, P), generator=torch.Generator().manual_seed(7))
    orig_prefill = ref_model.prefill_static
    if B > 1:
        ref_model.prefill_static = _prefill_repeat
    try:
        with LogitTap() as tap:
            torch.manual_seed(seed)
            codes = ref.generate(cond, audio_prefix_codes=prefix, max_new_tokens=N, cfg_scale=2.0, batch_size=B,
                                 sampling_params=dict(sampling_params), disable_torch_compile=True)
    finally:
        ref_model.prefill_static = orig_prefill
    dims_o = BackboneDims(d_model=dims["d_model"], n_layer=dims["n_layer"], n_heads=dims["n_heads"],
                          n_heads_kv=dims["n_heads_kv"], d_ff=dims["d_ff"])
    oracle = TransformerOracle(w, dims_o, torch.bfloat16)
    trace = {}
    torch.manual_seed(seed)
    mine = o_gen.generate(oracle, cond, prefix, N, 2.0, B, dict(sampling_params), trace=trace)
    assert mine.shape == codes.shape and (mine == codes).all(), name
    assert len(trace["logits"]) == len(tap.logits)
    for a, b in zip(trace["logits"], tap.logits):
        assert torch.equal(a, b), name                       # same primitives on the same CPU: bit-identical
    n_eos = int((trace["delayed"] == 1024).sum())
    print(f"{name}: codes {tuple(codes.shape)} steps {trace['steps']} offset {trace['offset']} eos_tokens {n_eos}")
    keep = sorted(set([0, 1, 2, len(tap.logits) // 2, len(tap.logits) - 1]))
    np.savez_compressed(os.path.join(HERE, f"generate_{name}.npz"), codes=codes.numpy(),
                        delayed=trace["delayed"].numpy(), offset=trace["offset"], steps=trace["steps"],
                        logit_steps=np.array(keep), logits=torch.stack([tap.logits[i] for i in keep]).numpy(),
                        meta=np.array([B, Lc, N, P, seed]), eos_boost=eos_boost)
    return n_eos


def golden_generate():
    w = make_backbone_weights(**TINY_DIMS, seed=11)
    run_generate_case("tiny_b1", TINY_DIMS, w, B=1, Lc=24, N=40, P=0, seed=420, sampling_params=dict(min_p=0.1))
    n = run_generate_case("tiny_b1_eos", TINY_DIMS, w, B=1, Lc=17, N=96, P=0, seed=421,
                          sampling_params=dict(min_p=0.1), eos_boost=5.0)
    assert n > 0, "EOS case never hit EOS; pick another seed/boost"
    n = run_generate_case("tiny_b2_prefix_eos", TINY_DIMS, w, B=2, Lc=19, N=72, P=11, seed=422,
                          sampling_params=dict(linear=0.5, conf=0.4, quad=0.0), eos_boost=5.0)
    run_generate_case("tiny_b1_greedy", TINY_DIMS, w, B=1, Lc=8, N=24, P=3, seed=1,
                      sampling_params=dict(temperature=0.0))


# ---------------------------------------------------------------- DAC decode
def golden_dac():
    ae = ref_ae.DACAutoencoder()
    codes = torch.randint(0, 1024, (2, 9, 10), generator=torch.Generator().manual_seed(3))
    wav = ae.decode(codes)                                    # CPU: autocast disabled => fp32 (autoencoder.py:138)
    mine = o_dac.decode(make_dac_weights(seed=1), codes)
    err = (mine - wav).abs().max().item()
    print("dac: wav", tuple(wav.shape), "absmax", wav.abs().max().item(), "oracle max err", err)
    assert wav.shape == (2, 1, 5120) and err < 2e-5
    np.savez_compressed(os.path.join(HERE, "dac_decode.npz"), codes=codes.numpy(), wav=wav.numpy())


# ---------------------------------------------------------------- DAC encode (audio prefix -> codes)
def golden_dac_encode():
    """`DACAutoencoder.encode` (zonos/autoencoder.py:104-117) = transformers' `DacModel.encode(wav).audio_codes` in fp32, on
    seeded encoder weights and a seeded waveform of 12 frames (2 utterances)."""
    w = make_dac_weights(seed=1, with_encoder=True)
    ae = ref_ae.DACAutoencoder()
    missing = ae.dac.load_state_dict(w, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    g = torch.Generator().manual_seed(8)
    t = torch.arange(12 * 512) / 44100.0
    wav = torch.stack([0.4 * torch.sin(2 * math.pi * 220 * t) + 0.05 * torch.randn(t.shape, generator=g),
                       0.3 * torch.sin(2 * math.pi * 523 * t) * torch.sin(2 * math.pi * 3 * t) + 0.05 * torch.randn(t.shape, generator=g)]).unsqueeze(1)
    codes = ae.encode(wav)
    z = ae.dac.encoder(wav)
    mine_z = o_dac.encode_latents(w, wav)
    margins = []
    mine = o_dac.quantize(w, mine_z, margins)
    print("dac encode: codes", tuple(codes.shape), "latent absmax", z.abs().max().item(), "oracle latent err", (mine_z - z).abs().max().item(),
          "codes equal", bool(torch.equal(mine, codes)), "smallest margin", min(m.min().item() for m in margins))
    assert codes.shape == (2, 9, 12) and torch.equal(mine, codes) and torch.equal(mine_z, z)
    np.savez_compressed(os.path.join(HERE, "dac_encode.npz"), wav=wav.numpy(), codes=codes.numpy(), latents=z.numpy())


# ---------------------------------------------------------------- prefix conditioner (tensor conditioners)
COND_CFG = dict(projection="linear", conditioners=[
    dict(type="PassthroughConditioner", name="speaker", cond_dim=128, uncond_type="learned", projection="linear"),
    dict(type="FourierConditioner", name="emotion", input_dim=8, uncond_type="learned"),
    dict(type="FourierConditioner", name="fmax", min_val=0, max_val=24000, uncond_type="learned"),
    dict(type="FourierConditioner", name="pitch_std", min_val=0, max_val=400, uncond_type="learned"),
    dict(type="IntegerConditioner", name="language_id", min_val=-1, max_val=126, uncond_type="learned")])


def golden_conditioner():
    from zonos.conditioning import PrefixConditioner as RefPC
    from zonos.config import PrefixConditionerConfig as RefCfg
    from zonos.utilities.conditioning_cache import prepare_conditioning_with_cache as ref_prepare
    from zonos_b200.conditioning import PrefixConditioner, prepare_conditioning_with_cache
    from zonos_b200.config import PrefixConditionerConfig
    torch.manual_seed(21)
    ref = RefPC(RefCfg(**COND_CFG), 64).eval()
    with torch.no_grad():
        for c in ref.conditioners:
            c.uncond_vector.normal_()
    g = torch.Generator().manual_seed(22)
    cond = {"speaker": torch.randn(1, 1, 128, generator=g), "emotion": torch.rand(1, 1, 8, generator=g),
            "fmax": torch.tensor([[[22050.0]]]), "pitch_std": torch.tensor([[[45.0]]]), "language_id": torch.tensor([[[24]]])}
    uncond = {"emotion": cond["emotion"]}
    out = ref_prepare(ref, cond, uncond, cfg_scale=2.0)
    mine = PrefixConditioner(PrefixConditionerConfig(**COND_CFG), 64).eval()
    mine.load_state_dict(ref.state_dict())
    got = prepare_conditioning_with_cache(mine, cond, uncond, cfg_scale=2.0)
    assert torch.equal(got, out), (got - out).abs().max()
    np.savez_compressed(os.path.join(HERE, "prefix_conditioner.npz"), out=out.numpy(),
                        **{"sd__" + k: v.numpy() for k, v in ref.state_dict().items()},
                        **{"in__" + k: v.numpy() for k, v in cond.items()})
    print("prefix conditioner:", tuple(out.shape))


if __name__ == "__main__":
    golden_conditioner()
    golden_codebook()
    golden_sampler()
    golden_generate()
    golden_dac()
    golden_dac_encode()
    print("golden fixtures written to", HERE)
