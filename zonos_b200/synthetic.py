"""Seeded synthetic weights / inputs for tests and benchmarks (no checkpoints, no network).

Shapes and key names are the reference's state_dict:
`backbone.layers.{i}.{norm,norm2}.{weight,bias}`, `...mixer.{in_proj,out_proj}.weight`,
`...mlp.{fc1,fc2}.weight`, `backbone.norm_f.*`, `embeddings.{k}.weight [1032,D]`,
`fused_heads.weight [9*1025, D]`; DAC keys as `transformers.DacModel.state_dict()`.
Initial scales follow torch's defaults (Linear/Conv: U(+-1/sqrt(fan_in)), Embedding: N(0,1))
so activations look like the reference's random-init model; norm weights/biases are
perturbed so that parity tests exercise them.  Everything is drawn from a CPU
`torch.Generator`, so the same seed gives the same tensors on every machine.
"""
import math
import torch

TRANSFORMER_DIMS = dict(d_model=2048, n_layer=26, n_heads=16, n_heads_kv=4, d_ff=8192)
TINY_DIMS = dict(d_model=512, n_layer=2, n_heads=4, n_heads_kv=2, d_ff=1024)
DAC_STRIDES = (8, 8, 4, 2)


def _uniform(gen, shape, bound, dtype):
    return ((torch.rand(shape, generator=gen) * 2 - 1) * bound).to(dtype)


def make_backbone_weights(d_model=2048, n_layer=26, n_heads=16, n_heads_kv=4, d_ff=8192,
                          n_codebooks=9, head_vocab=1025, emb_vocab=1032, seed=0,
                          dtype=torch.bfloat16, heads_scale=1.0, eos_off=False) -> dict:
    g = torch.Generator().manual_seed(seed)
    hd = d_model // n_heads
    w = {}
    for i in range(n_layer):
        p = f"backbone.layers.{i}."
        for n in ("norm", "norm2"):
            w[p + n + ".weight"] = (1.0 + 0.1 * torch.randn(d_model, generator=g)).to(dtype)
            w[p + n + ".bias"] = (0.05 * torch.randn(d_model, generator=g)).to(dtype)
        w[p + "mixer.in_proj.weight"] = _uniform(g, ((n_heads + 2 * n_heads_kv) * hd, d_model), 1 / math.sqrt(d_model), dtype)
        w[p + "mixer.out_proj.weight"] = _uniform(g, (d_model, n_heads * hd), 1 / math.sqrt(n_heads * hd), dtype)
        w[p + "mlp.fc1.weight"] = _uniform(g, (2 * d_ff, d_model), 1 / math.sqrt(d_model), dtype)
        w[p + "mlp.fc2.weight"] = _uniform(g, (d_model, d_ff), 1 / math.sqrt(d_ff), dtype)
    w["backbone.norm_f.weight"] = (1.0 + 0.1 * torch.randn(d_model, generator=g)).to(dtype)
    w["backbone.norm_f.bias"] = (0.05 * torch.randn(d_model, generator=g)).to(dtype)
    for k in range(n_codebooks):
        w[f"embeddings.{k}.weight"] = torch.randn(emb_vocab, d_model, generator=g).to(dtype)
    w["fused_heads.weight"] = _uniform(g, (n_codebooks * head_vocab, d_model), heads_scale / math.sqrt(d_model), dtype)
    if eos_off:
        # Throughput runs want a deterministic length: pin one coordinate of the final norm's output
        # to a constant (weight 0, bias 8) and let ONLY the codebook-0 EOS head row read it with a large negative
        # weight, so that logit is ~ -240 on every step and EOS is never sampled.
        w["backbone.norm_f.weight"][0] = 0.0
        w["backbone.norm_f.bias"][0] = 8.0
        w["fused_heads.weight"][:, 0] = 0.0
        w["fused_heads.weight"][head_vocab - 1] = 0.0
        w["fused_heads.weight"][head_vocab - 1, 0] = -30.0
    return w


DAC_ENC_STRIDES = (2, 4, 8, 8)


def make_dac_weights(seed=1, dtype=torch.float32, n_codebooks=9, gain=1.3, with_encoder=False) -> dict:
    """DacModel(DacConfig(sampling_rate=44100)) tensors: quantizer tables + decoder; with_encoder adds the encoder and the
    quantizers' in_proj (drawn from a second generator, so the decode-side tensors do not depend on the flag)."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(name, cout, cin, k, transposed=False):
        fan_in = (cout if transposed else cin) * k          # torch's fan_in for ConvTranspose1d uses dim 1
        b = gain / math.sqrt(fan_in)
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        w[name + ".weight"] = _uniform(g, shape, b, dtype)
        w[name + ".bias"] = _uniform(g, (cout,), b, dtype)

    for k in range(n_codebooks):
        p = f"quantizer.quantizers.{k}."
        w[p + "codebook.weight"] = torch.randn(1024, 8, generator=g).to(dtype)
        conv(p + "out_proj", 1024, 8, 1)
    conv("decoder.conv1", 1536, 1024, 7)
    ch = 1536
    for i, s in enumerate(DAC_STRIDES):
        p = f"decoder.block.{i}."
        w[p + "snake1.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
        conv(p + "conv_t1", ch // 2, ch, 2 * s, transposed=True)
        ch //= 2
        for j in (1, 2, 3):
            r = p + f"res_unit{j}."
            w[r + "snake1.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
            conv(r + "conv1", ch, ch, 7)
            w[r + "snake2.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
            conv(r + "conv2", ch, ch, 1)
    w["decoder.snake1.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
    conv("decoder.conv2", 1, ch, 7)
    if with_encoder:
        g = torch.Generator().manual_seed(seed + 1000)        # (conv() draws from `g`)
        for k in range(n_codebooks):
            conv(f"quantizer.quantizers.{k}.in_proj", 8, 1024, 1)
        conv("encoder.conv1", 64, 1, 7)
        ch = 64
        for i, s in enumerate(DAC_ENC_STRIDES):
            p = f"encoder.block.{i}."
            for j in (1, 2, 3):
                r = p + f"res_unit{j}."
                w[r + "snake1.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
                conv(r + "conv1", ch, ch, 7)
                w[r + "snake2.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
                conv(r + "conv2", ch, ch, 1)
            w[p + "snake1.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
            conv(p + "conv1", 2 * ch, ch, 2 * s)
            ch *= 2
        w["encoder.snake1.alpha"] = (0.5 + torch.rand(1, ch, 1, generator=g)).to(dtype)
        conv("encoder.conv2", 1024, ch, 3)
    return w


def make_conditioning(rows: int, cond_len: int, d_model: int, seed=1234) -> torch.Tensor:
    """bf16 [rows, cond_len, d_model] ~ N(0,1): the post-LayerNorm scale of zonos/conditioning.py:522."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, cond_len, d_model, generator=g).bfloat16()


HYBRID_TINY_DIMS = dict(d_model=512, n_layer=4, attn_layer_idx=(2,), n_heads=4, n_heads_kv=2, d_ff=1024)


def make_hybrid_weights(d_model=2048, n_layer=46, attn_layer_idx=(9, 18, 27, 36, 45), n_heads=16, n_heads_kv=4, d_ff=8192,
                        d_state=128, d_conv=4, expand=2, m_headdim=64, ngroups=1, n_codebooks=9, head_vocab=1025, emb_vocab=1032,
                        seed=0, dtype=torch.bfloat16, heads_scale=1.0) -> dict:
    """Hybrid (Mamba2 + attention) backbone with mamba_ssm's parameter names (`mixer.in_proj/conv1d/dt_bias/A_log/D/norm/
    out_proj` for Mamba2 layers; attention layers as in the transformer variant) and its default initialisers' scales."""
    g = torch.Generator().manual_seed(seed)
    hd = d_model // n_heads
    d_inner = expand * d_model
    nheads = d_inner // m_headdim
    conv_dim = d_inner + 2 * ngroups * d_state
    w = {}
    for i in range(n_layer):
        p = f"backbone.layers.{i}."
        w[p + "norm.weight"] = (1.0 + 0.1 * torch.randn(d_model, generator=g)).to(dtype)
        w[p + "norm.bias"] = (0.05 * torch.randn(d_model, generator=g)).to(dtype)
        if i in attn_layer_idx:
            w[p + "mixer.in_proj.weight"] = _uniform(g, ((n_heads + 2 * n_heads_kv) * hd, d_model), 1 / math.sqrt(d_model), dtype)
            w[p + "mixer.out_proj.weight"] = _uniform(g, (d_model, n_heads * hd), 1 / math.sqrt(n_heads * hd), dtype)
            w[p + "norm2.weight"] = (1.0 + 0.1 * torch.randn(d_model, generator=g)).to(dtype)
            w[p + "norm2.bias"] = (0.05 * torch.randn(d_model, generator=g)).to(dtype)
            w[p + "mlp.fc1.weight"] = _uniform(g, (2 * d_ff, d_model), 1 / math.sqrt(d_model), dtype)
            w[p + "mlp.fc2.weight"] = _uniform(g, (d_model, d_ff), 1 / math.sqrt(d_ff), dtype)
        else:
            w[p + "mixer.in_proj.weight"] = _uniform(g, (2 * d_inner + 2 * ngroups * d_state + nheads, d_model), 1 / math.sqrt(d_model), dtype)
            w[p + "mixer.conv1d.weight"] = _uniform(g, (conv_dim, 1, d_conv), 1 / math.sqrt(d_conv), dtype)
            w[p + "mixer.conv1d.bias"] = _uniform(g, (conv_dim,), 1 / math.sqrt(d_conv), dtype)
            dt = torch.exp(torch.rand(nheads, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
            w[p + "mixer.dt_bias"] = (dt + torch.log(-torch.expm1(-dt))).to(dtype)          # inverse softplus
            w[p + "mixer.A_log"] = torch.log(1 + 15 * torch.rand(nheads, generator=g)).to(dtype)
            w[p + "mixer.D"] = torch.ones(nheads).to(dtype)
            w[p + "mixer.norm.weight"] = (1.0 + 0.1 * torch.randn(d_inner, generator=g)).to(dtype)
            w[p + "mixer.out_proj.weight"] = _uniform(g, (d_model, d_inner), 1 / math.sqrt(d_inner), dtype)
    w["backbone.norm_f.weight"] = (1.0 + 0.1 * torch.randn(d_model, generator=g)).to(dtype)
    w["backbone.norm_f.bias"] = (0.05 * torch.randn(d_model, generator=g)).to(dtype)
    for k in range(n_codebooks):
        w[f"embeddings.{k}.weight"] = torch.randn(emb_vocab, d_model, generator=g).to(dtype)
    w["fused_heads.weight"] = _uniform(g, (n_codebooks * head_vocab, d_model), heads_scale / math.sqrt(d_model), dtype)
    return w
