"""CPU: the oracle reproduces every golden fixture recorded from the reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import SAMPLER_CASES, eos_boosted, load_golden, oracle_dims, sampler_inputs
from oracle import codebook as o_cb, dac as o_dac, generate as o_gen, sampling as o_samp
from oracle.transformer import TransformerOracle
from zonos_b200.synthetic import TINY_DIMS, make_backbone_weights, make_conditioning, make_dac_weights


def test_codebook_pattern_known_answer():
    g = load_golden("codebook_pattern.npz")
    assert (o_cb.apply_delay_pattern(g["kat_in"], 1025) == g["kat_out"]).all()      # zonos/codebook_pattern.py:26-29
    assert (o_cb.apply_delay_pattern(g["codes"], 1025) == g["delayed"]).all()
    assert (o_cb.revert_delay_pattern(g["delayed"]) == g["codes"]).all()


def test_codebook_pattern_edge_cases():
    empty = np.zeros((2, 9, 0), dtype=np.int64)
    d = o_cb.apply_delay_pattern(empty, 1025)
    assert d.shape == (2, 9, 9) and (d == 1025).all()
    assert o_cb.revert_delay_pattern(d).shape == (2, 9, 0)


@pytest.mark.parametrize("i", range(len(SAMPLER_CASES)))
def test_sampler_matches_reference_tokens(i):
    g = load_golden("sampler.npz")
    logits, window, q = sampler_inputs(i)
    torch.manual_seed(int(g["seeds"][i]))
    q_ref = torch.empty_like(logits).exponential_(1)          # the draws the reference made under that seed
    tok = o_samp.sample_from_logits(logits.clone(), q=q_ref, generated_tokens=window, **SAMPLER_CASES[i])
    assert (tok.numpy() == g["tokens"][2 * i]).all()
    tok2 = o_samp.sample_from_logits(logits.clone(), q=q, generated_tokens=window, **SAMPLER_CASES[i])
    assert (tok2.numpy() == g["tokens"][2 * i + 1]).all()


GEN_CASES = {
    "tiny_b1": dict(sp=dict(min_p=0.1), boost=0.0),
    "tiny_b1_eos": dict(sp=dict(min_p=0.1), boost=5.0),
    "tiny_b2_prefix_eos": dict(sp=dict(linear=0.5, conf=0.4, quad=0.0), boost=5.0),
    "tiny_b1_greedy": dict(sp=dict(temperature=0.0), boost=0.0),
}


@pytest.mark.parametrize("name", list(GEN_CASES))
def test_generate_matches_reference(name):
    g = load_golden(f"generate_{name}.npz")
    B, Lc, N, P, seed = (int(v) for v in g["meta"])
    w = make_backbone_weights(**TINY_DIMS, seed=11)
    if GEN_CASES[name]["boost"]:
        w = eos_boosted(w, GEN_CASES[name]["boost"])
    oracle = TransformerOracle(w, oracle_dims(TINY_DIMS), torch.bfloat16)
    cond = make_conditioning(2 * B, Lc, TINY_DIMS["d_model"], seed=1234 + B)
    prefix = torch.randint(0, 1024, (B, 9, P), generator=torch.Generator().manual_seed(7)) if P else None
    trace = {}
    torch.manual_seed(seed)
    codes = o_gen.generate(oracle, cond, prefix, N, 2.0, B, dict(GEN_CASES[name]["sp"]), trace=trace)
    assert codes.shape == g["codes"].shape and (codes.numpy() == g["codes"]).all()
    assert (trace["delayed"].numpy() == g["delayed"]).all() and trace["offset"] == int(g["offset"])
    for j, step in enumerate(g["logit_steps"]):
        assert np.array_equal(trace["logits"][int(step)].numpy(), g["logits"][j])      # bit-identical on CPU


def test_dac_encode_matches_reference():
    """`DACAutoencoder.encode` of the reference (transformers `DacModel.encode`, fp32) on the seeded encoder: the oracle's
    latents are bit-identical to the recorded ones on the CPU and every code index is equal."""
    g = load_golden("dac_encode.npz")
    w = make_dac_weights(seed=1, with_encoder=True)
    wav = torch.from_numpy(g["wav"])
    z = o_dac.encode_latents(w, wav)
    assert np.abs(z.numpy() - g["latents"]).max() < 1e-5
    margins = []
    codes = o_dac.quantize(w, torch.from_numpy(g["latents"]), margins)
    assert np.array_equal(codes.numpy(), g["codes"]) and codes.dtype == torch.int64
    assert np.array_equal(o_dac.encode(w, wav).numpy(), g["codes"])
    assert min(m.min().item() for m in margins) > 1e-4        # the fixture holds no decision within fp32 summation noise


def test_dac_decode_matches_reference():
    g = load_golden("dac_decode.npz")
    wav = o_dac.decode(make_dac_weights(seed=1), torch.from_numpy(g["codes"]))
    assert wav.shape == g["wav"].shape
    assert np.abs(wav.numpy() - g["wav"]).max() < 2e-5
