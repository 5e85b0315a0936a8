"""CPU: the hybrid (Mamba2) oracle against the independent pure-PyTorch Mamba2 of `transformers` (second opinion; the
reference's own hybrid backbone cannot be imported - PARITY UNPINNED, see oracle/hybrid.py)."""
import torch

from oracle.hybrid import HybridDims, HybridOracle
from zonos_b200.synthetic import HYBRID_TINY_DIMS, make_hybrid_weights


def test_mamba2_mixer_matches_transformers_torch_forward():
    from transformers.models.mamba2.configuration_mamba2 import Mamba2Config
    from transformers.models.mamba2.modeling_mamba2 import Mamba2Mixer
    D = 256
    cfg = Mamba2Config(hidden_size=D, state_size=128, conv_kernel=4, expand=2, head_dim=64, num_heads=8, n_groups=1, rms_norm=True,
                       use_bias=False, use_conv_bias=True, chunk_size=4, layer_norm_epsilon=1e-5)
    torch.manual_seed(0)
    mixer = Mamba2Mixer(cfg, layer_idx=0).float().eval()
    with torch.no_grad():
        mixer.A_log.copy_(torch.log(1 + 15 * torch.rand(8)))
        mixer.dt_bias.copy_(torch.randn(8) * 0.5)
        mixer.D.copy_(torch.rand(8) + 0.5)
        mixer.norm.weight.copy_(1 + 0.1 * torch.randn(512))
    dims = HybridDims(d_model=D, n_layer=1, attn_layer_idx=(), n_heads=2, n_heads_kv=1, d_ff=512)
    w = {"backbone.layers.0.mixer." + k: v.detach().clone() for k, v in mixer.state_dict().items()}
    w.update({"backbone.layers.0.norm.weight": torch.ones(D), "backbone.layers.0.norm.bias": torch.zeros(D),
              "backbone.norm_f.weight": torch.ones(D), "backbone.norm_f.bias": torch.zeros(D)})
    oracle = HybridOracle(w, dims, torch.float32)
    R, T = 2, 8
    u = torch.randn(R, T, D)
    with torch.no_grad():
        ref = mixer.torch_forward(u)
    st = oracle.allocate(R, T)
    got = oracle._mamba(u, st, 0)
    assert (got - ref).abs().max() < 2e-4, (got - ref).abs().max()
    # token-by-token decode on the carried state equals the one-shot prefill (fp32: no state rounding)
    st2 = oracle.allocate(R, T)
    steps = torch.cat([oracle._mamba(u[:, t:t + 1], st2, 0) for t in range(T)], dim=1)
    assert (steps - got).abs().max() < 2e-4
    assert (st2.ssm[0] - st.ssm[0]).abs().max() < 1e-4 and (st2.conv[0] - st.conv[0]).abs().max() < 1e-5


def test_hybrid_oracle_generate_runs_and_is_deterministic():
    from oracle import generate as o_gen
    from zonos_b200.synthetic import make_conditioning
    dims = HybridDims(**HYBRID_TINY_DIMS)
    w = make_hybrid_weights(**HYBRID_TINY_DIMS, seed=3)
    oracle = HybridOracle(w, dims, torch.bfloat16)
    cond = make_conditioning(2, 7, dims.d_model, seed=2)
    torch.manual_seed(5)
    a = o_gen.generate(oracle, cond, None, 10, 2.0, 1, dict(min_p=0.1))
    torch.manual_seed(5)
    b = o_gen.generate(oracle, cond, None, 10, 2.0, 1, dict(min_p=0.1))
    assert a.shape == (1, 9, 10) and torch.equal(a, b)
