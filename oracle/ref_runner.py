"""Test / measurement infrastructure, NOT product code: drives the UNMODIFIED reference (langfod/Zonos) on the CPU.

`__graft_entry__.build()` stages the reference's pure-Python package into the git-ignored `oracle/_ref/zonos` when
`/root/reference` is present (build container); the directory travels to the GPU box with the snapshot.  Only
`bench.py`'s `--impl reference` / `cpu_baseline` legs use this module.  The reference's text front end needs packages
that are not in this image (phonemizer, inflect, kanjize, sudachipy): they are stubbed exactly as in
`tests/golden/make_golden.py` - the hot path (zonos/model.py:354-548 `Zonos.generate`, zonos/autoencoder.py:119-140
`DACAutoencoder.decode`) never touches them.  The DAC weights come from `transformers`' `DacModel` with the seeded
synthetic state dict (no checkpoint download), as in SURVEY.md Appendix A.
"""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FRAME_RATE = 44100 / 512


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "zonos", "model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_loaded = None


def load():
    """Import the staged reference (once) and return (zonos.model, zonos.autoencoder, BACKBONES, ZonosConfig)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref/zonos is missing: run __graft_entry__.build() where /root/reference exists")
    _stub("phonemizer"); _stub("phonemizer.backend", EspeakBackend=object)
    _stub("inflect", engine=lambda: types.SimpleNamespace(number_to_words=lambda *a, **k: ""))
    _stub("kanjize", number2kanji=str)

    class _D:
        def __init__(self, dict=None): pass
        def create(self): return self

    _stub("sudachipy", Dictionary=_D, SplitMode=types.SimpleNamespace(A=0))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import contextlib
    try:                                         # the reference logs every generate call at DEBUG level
        from loguru import logger
        logger.remove()
    except Exception:
        pass
    with contextlib.redirect_stdout(sys.stderr):  # its backbone registry prints the mamba_ssm ImportError traceback to stdout
        import zonos.autoencoder as ref_ae
        import zonos.model as ref_model
        from zonos.backbone import BACKBONES
        from zonos.config import ZonosConfig
    _loaded = (ref_model, ref_ae, BACKBONES, ZonosConfig)
    return _loaded


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


def build_model(dims, weights, dac_weights, device="cpu"):
    """The reference's own `Zonos` (torch backbone) with the synthetic weights; its DAC built from `transformers`."""
    import torch
    from transformers.models.dac import DacConfig, DacModel
    ref_model, ref_ae, BACKBONES, ZonosConfig = load()

    def _dac_init(self):
        self.dac = DacModel(DacConfig(sampling_rate=44100)).eval().requires_grad_(False)
        self.dac.load_state_dict(dac_weights, strict=False)
        self.codebook_size, self.num_codebooks = self.dac.config.codebook_size, self.dac.quantizer.n_codebooks
        self.sampling_rate = self.dac.config.sampling_rate

    ref_ae.DACAutoencoder.__init__ = _dac_init
    m = ref_model.Zonos(ZonosConfig.from_dict(ref_config(dims)), BACKBONES["torch"]).to(device, torch.bfloat16)
    sd = m.state_dict()
    sd.update({k: v.to(device) for k, v in weights.items()})
    m.load_state_dict(sd)
    m.autoencoder.dac.to(device)
    return m.eval().requires_grad_(False)


def step(model, cond, frames, seed=420):
    """One bounded sample of the workload through the reference's public API: generate `frames` frames + DAC decode.
    Returns (audio seconds produced, wall seconds)."""
    import torch
    torch.manual_seed(seed)
    dev = cond.device
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    import contextlib
    t0 = time.perf_counter()
    with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):
        codes = model.generate(cond, max_new_tokens=frames, cfg_scale=2.0, batch_size=cond.shape[0] // 2,
                               sampling_params=dict(min_p=0.1), disable_torch_compile=True)
        wav = model.autoencoder.decode(codes)
        if dev.type == "cuda":
            wav = wav.cpu()
            torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return codes.shape[-1] / FRAME_RATE * codes.shape[0], dt
