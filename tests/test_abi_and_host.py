"""CPU: the C-ABI library loads and exports what include/zonos_b200.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT
from oracle import codebook as o_cb


def header_symbols():
    text = open(os.path.join(ROOT, "include", "zonos_b200.h")).read()
    return sorted(set(re.findall(r"ZB_API\s+[\w\s\*]+?\b(zb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from zonos_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    assert _lib.load().zb_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header_layout():
    from zonos_b200 import _lib
    assert ctypes.sizeof(_lib.zb_layer) == 8 + 14 * 8
    assert ctypes.sizeof(_lib.zb_sampling) == 36
    assert ctypes.sizeof(_lib.zb_cache) == 16 + 5 * 8
    assert ctypes.sizeof(_lib.zb_gen_progress) == 16
    assert ctypes.sizeof(_lib.zb_model_desc) == 12 * 4 + 4 + 4 + 6 * 8 + 6 * 4


def test_no_cpu_fallback():
    from zonos_b200 import _lib, sample_from_logits
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.context("cpu")
    with pytest.raises(RuntimeError):
        sample_from_logits(torch.zeros(1, 9, 1025))
    with pytest.raises(RuntimeError):
        _lib.Context(0)          # no device: zb_ctx_create fails loudly


def test_torch_custom_ops_are_registered_and_have_no_cpu_path():
    """north_star: "thin PyTorch custom ops over a C-ABI layer" - zonos_b200/ops.py registers torch.ops.zonos_b200.*; every
    op refuses CPU tensors (no fallback) and has a fake implementation (shape inference without a GPU)."""
    from zonos_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.zonos_b200, name), name
    codes = torch.zeros(2, 9, 5, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        torch.ops.zonos_b200.embed_codes(0, codes, 2, 256)
    with pytest.raises(RuntimeError):
        torch.ops.zonos_b200.dac_decode(0, codes, 512)
    with pytest.raises(RuntimeError):
        torch.ops.zonos_b200.sample_update(torch.zeros(1, 9, 1025), None, None, [1.0, 0.0, 0.1, 0.0, 0.0, 0.0, 3.0], 0, 2, 0, 0, False)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        c = torch.zeros(2, 9, 5, dtype=torch.int64)
        assert torch.ops.zonos_b200.embed_codes(0, c, 2, 256).shape == (4, 5, 256)
        assert torch.ops.zonos_b200.dac_decode(0, c, 512).shape == (2, 1, 2560)
        assert torch.ops.zonos_b200.heads_cfg(0, torch.zeros(4, 256, dtype=torch.bfloat16), 2.0, 9, 1025).shape == (2, 9, 1025)


def test_host_codebook_pattern_matches_oracle():
    from zonos_b200 import apply_delay_pattern, revert_delay_pattern
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(-1, 1026, (2, 9, 23), generator=g)
    d = apply_delay_pattern(codes, 1025)
    assert (d.numpy() == o_cb.apply_delay_pattern(codes.numpy(), 1025)).all()
    assert (revert_delay_pattern(d) == codes).all()


def test_config_and_state_dict_names():
    from zonos_b200 import B200ZonosBackbone, Zonos, ZonosConfig, transformer_config_dict
    from zonos_b200.model import find_multiple
    from zonos_b200.synthetic import TINY_DIMS, make_backbone_weights
    assert find_multiple(1026, 8) == 1032 and find_multiple(16, 8) == 16 and find_multiple(5, 0) == 5
    cfg = ZonosConfig.from_dict(transformer_config_dict(**TINY_DIMS))
    assert cfg.eos_token_id == 1024 and cfg.masked_token_id == 1025 and cfg.codebook_dimension == 9
    m = Zonos(cfg)
    keys = set(m.state_dict())
    assert keys == set(make_backbone_weights(**TINY_DIMS))          # the reference's checkpoint key names
    assert "backbone.layers.0.mixer.in_proj.weight" in keys and "fused_heads.weight" in keys
    # checkpoints store heads.{i}.weight: they are fused on load (zonos/model.py:208-223)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    fused = sd.pop("fused_heads.weight")
    for i in range(9):
        sd[f"heads.{i}.weight"] = fused[i * 1025:(i + 1) * 1025]
    m2 = Zonos(cfg)
    m2.load_state_dict(sd)
    assert torch.equal(m2.fused_heads.weight, fused)
    assert isinstance(m.backbone, B200ZonosBackbone)


def test_finalize_matches_oracle():
    from oracle.generate import finalize_codes
    from zonos_b200 import Zonos, ZonosConfig, transformer_config_dict
    from zonos_b200.synthetic import TINY_DIMS
    m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**TINY_DIMS)))
    g = torch.Generator().manual_seed(9)
    for trial in range(20):
        B, T = 1 + trial % 3, 30 + trial
        delayed = torch.randint(0, 1026, (B, 9, T + 9), generator=g)
        if trial % 2:
            delayed[:, :, T - 8:] = 1024
        offset = T + 9 - (trial % 5)
        assert torch.equal(m._finalize(delayed.clone(), offset), finalize_codes(delayed.clone(), offset))


def test_prefix_conditioner_matches_reference_golden():
    """Tensor conditioners + projection + LayerNorm + CFG concatenation vs the fixture recorded from the reference."""
    from helpers import load_golden
    from zonos_b200.conditioning import PrefixConditioner, make_cond_dict, prepare_conditioning_with_cache, tokenize_phonemes
    from zonos_b200.config import PrefixConditionerConfig
    cfg = dict(projection="linear", conditioners=[
        dict(type="PassthroughConditioner", name="speaker", cond_dim=128, uncond_type="learned", projection="linear"),
        dict(type="FourierConditioner", name="emotion", input_dim=8, uncond_type="learned"),
        dict(type="FourierConditioner", name="fmax", min_val=0, max_val=24000, uncond_type="learned"),
        dict(type="FourierConditioner", name="pitch_std", min_val=0, max_val=400, uncond_type="learned"),
        dict(type="IntegerConditioner", name="language_id", min_val=-1, max_val=126, uncond_type="learned")])
    g = load_golden("prefix_conditioner.npz")
    pc = PrefixConditioner(PrefixConditionerConfig(**cfg), 64).eval()
    pc.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd__")})
    cond = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in__")}
    cache = {}
    with torch.no_grad():
        out = prepare_conditioning_with_cache(pc, cond, {"emotion": cond["emotion"]}, use_cache=True, cfg_scale=2.0, cache=cache)
        again = prepare_conditioning_with_cache(pc, cond, {"emotion": cond["emotion"]}, use_cache=True, cfg_scale=2.0, cache=cache)
    assert out.shape == (2, 5, 64) and np.array_equal(out.numpy(), g["out"])
    assert again is out and len(cache) == 1
    with pytest.raises(ValueError):
        PrefixConditioner(PrefixConditionerConfig(projection="none", conditioners=[dict(type="PassthroughConditioner", name="x")]), 8)({})
    ids, lens = tokenize_phonemes(["həlˈoʊ", "a"])
    assert ids.shape == (2, 8) and ids[1, :5].tolist() == [0] * 5 and ids[0, 0] == 2 and ids[0, -1] == 3 and lens == [8, 3]
    d = make_cond_dict(text="hi", speaker=torch.zeros(1, 1, 128), device="cpu")
    assert set(d) == {"espeak", "speaker", "emotion", "fmax", "pitch_std", "speaking_rate", "language_id", "ctc_loss", "speaker_noised"}
    assert d["language_id"].item() == 24 and abs(float(d["emotion"].sum()) - 1.0) < 1e-6


def test_shipped_configs_load():
    """configs/*.json (the hybrid one states the assumed shape) parse into ZonosConfig."""
    import json
    from zonos_b200.config import ZonosConfig, hybrid_config_dict, transformer_config_dict
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")
    hyb = json.load(open(os.path.join(root, "zonos_v0.1_hybrid.json")))
    assert "ASSUMED" in hyb["_comment"]
    c = ZonosConfig.from_dict(hyb)
    assert c.backbone.n_layer == hybrid_config_dict()["backbone"]["n_layer"] and c.backbone.ssm_cfg == {"layer": "Mamba2"}
    t = ZonosConfig.from_dict(json.load(open(os.path.join(root, "zonos_v0.1_transformer.json"))))
    assert t.backbone.n_layer == 26 and t.backbone.attn_layer_idx == list(range(26))


def test_from_local_reads_a_reference_format_checkpoint(tmp_path):
    """zonos/model.py:128-176: `config.json` + `model.safetensors` as the reference's checkpoints store them - one
    `heads.{i}.weight` per codebook (fused row-wise on load, model.py:208-223), embedding tables with 1026 rows (padded to the
    multiple of 8 the module holds, model.py:164-172), bf16 tensors.  Loaded on the CPU here (state-dict plumbing only; the
    compute path has no CPU kernels)."""
    import json
    import safetensors.torch
    from zonos_b200 import Zonos, transformer_config_dict
    from zonos_b200.synthetic import TINY_DIMS, make_backbone_weights
    w = make_backbone_weights(**TINY_DIMS, seed=21)
    ckpt = {k: v.clone() for k, v in w.items() if k.startswith("backbone.")}
    fused = w["fused_heads.weight"]
    for i in range(9):
        ckpt[f"heads.{i}.weight"] = fused[i * 1025:(i + 1) * 1025].clone().contiguous()
        ckpt[f"embeddings.{i}.weight"] = w[f"embeddings.{i}.weight"][:1026].clone().contiguous()
    cfg_path, st_path = tmp_path / "config.json", tmp_path / "model.safetensors"
    cfg_path.write_text(json.dumps(transformer_config_dict(**TINY_DIMS)))
    safetensors.torch.save_file(ckpt, str(st_path))
    m = Zonos.from_local(str(cfg_path), str(st_path), device="cpu")
    sd = m.state_dict()
    assert sd["fused_heads.weight"].dtype == torch.bfloat16 and torch.equal(sd["fused_heads.weight"], fused)
    for i in range(9):
        e = sd[f"embeddings.{i}.weight"]
        assert e.shape[0] == 1032 and torch.equal(e[:1026], w[f"embeddings.{i}.weight"][:1026]) and not e[1026:].any()
    for k, v in w.items():
        if k.startswith("backbone."):
            assert torch.equal(sd[k], v), k


def test_fp8_quantiser_contract_on_the_host():
    """The host restatement of the FP8 mode's quantiser (tests/helpers.py: fp8_dequantised = decode.cu: quant_e4m3_kernel):
    every dequantised weight is exactly a bf16 number, is within e4m3's half-ulp (1/16 relative) of the original where the
    original is a normal e4m3 multiple of its row scale, and quantising the dequantised weights changes nothing - the property
    the GPU test relies on when it runs the bf16 kernel and the FP8 kernel on the same dequantised weights."""
    from helpers import FP8_KEYS, fp8_dequantised
    from zonos_b200.synthetic import TINY_DIMS, make_backbone_weights
    w = make_backbone_weights(**TINY_DIMS, seed=11)
    w["fused_heads.weight"][3] = 0                            # an all-zero row keeps scale 1 and stays zero
    d1 = fp8_dequantised(w)
    d2 = fp8_dequantised(d1)
    touched = 0
    for k, v in w.items():
        if not (k.endswith(FP8_KEYS) or k == "fused_heads.weight"):
            assert d1[k] is v
            continue
        touched += 1
        a, b = v.float(), d1[k].float()
        assert d1[k].dtype == v.dtype and torch.equal(b, b.bfloat16().float())
        assert torch.equal(d1[k], d2[k]), k
        amax = a.abs().amax(dim=1, keepdim=True)
        big = a.abs() >= amax / 64                            # normal e4m3 range of the row (scale <= amax < 2 scale, normals from scale / 64)
        rel = ((a - b).abs() / a.abs().clamp_min(1e-30))[big]
        assert float(rel.max()) <= 1 / 16 + 1e-6, (k, float(rel.max()))
        assert float((a - b).abs().max()) <= float(amax.max()) / 16
        # 8 distinct magnitudes per binade, sign kept
        assert torch.equal(torch.sign(b)[big], torch.sign(a)[big])
    assert touched == 4 * TINY_DIMS["n_layer"] + 1
    assert float(d1["fused_heads.weight"][3].abs().max()) == 0.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm): rank 0 prints ONE JSON line with the contract keys, running the
    unmodified reference from oracle/_ref when build() staged it (kind "reference") and the oracle port otherwise; every other
    rank exits 0 without work.  One layer and four frames: this checks the plumbing, not a number."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-frames", "4", "--layers", "1"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_per_second" and d["higher_is_better"] is True and d["value"] > 0
    staged = os.path.isdir(os.path.join(root, "oracle", "_ref", "zonos"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=root)
    assert other.returncode == 0 and other.stdout.strip() == ""
