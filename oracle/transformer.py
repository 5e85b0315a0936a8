"""Oracle: transformer backbone, codebook embedding sum, output heads + CFG mix.

CPU torch restatement of
  zonos/backbone/_torch.py:9-34   (rotary table)
  zonos/backbone/_torch.py:37-68  (interleaved-pair RoPE in fp32)
  zonos/backbone/_torch.py:71-107 (KV cache write, contiguous [R,S,2,Hkv,hd])
  zonos/backbone/_torch.py:213-238,307-328,376-420,456-474 (blocks)
  zonos/utilities/codec_utils.py:37,68-79 (embedding sum, fused heads)
  zonos/model.py:225-234 (last token -> heads -> fp32 -> CFG mix)

Weights are a flat dict keyed exactly like the reference's state_dict
(`backbone.layers.{i}.mixer.in_proj.weight`, ..., `embeddings.{k}.weight`,
`fused_heads.weight`).  With dtype=torch.bfloat16 the rounding points are the
reference's (every Linear / LayerNorm / SiLU / residual add rounds to bf16),
with dtype=torch.float32 nothing is rounded ("truth" for error budgets).

Reference quirks kept on purpose: `out_proj` is applied TWICE
(_torch.py:419-420, `out_proj_repeats=2`), LayerNorm with bias regardless of
`rms_norm` (_torch.py:155,278,280), value-first / gate-second SiLU MLP (:473).
"""
from dataclasses import dataclass, field
import torch
import torch.nn.functional as F


@dataclass
class BackboneDims:
    d_model: int = 2048
    n_layer: int = 26
    n_heads: int = 16
    n_heads_kv: int = 4
    d_ff: int = 8192
    norm_eps: float = 1e-5
    out_proj_repeats: int = 2
    n_codebooks: int = 9
    head_vocab: int = 1025
    emb_vocab: int = 1032

    @property
    def head_dim(self):
        return self.d_model // self.n_heads


def rotary_table(seq_len: int, head_dim: int, base: float = 10000.0) -> torch.Tensor:
    """_torch.py:29-34 -> fp32 [seq_len, head_dim/2, 2] (cos, sin)."""
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(seq_len), inv)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)


def apply_rope(x: torch.Tensor, cs: torch.Tensor) -> torch.Tensor:
    """_torch.py:57-68.  x [R,T,H,hd] ; cs fp32 [R,T,hd/2,2].  Pairs are (2i, 2i+1)."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    c = cs[:, :, None, :, 0]
    s = cs[:, :, None, :, 1]
    out = torch.stack([xf[..., 0] * c - xf[..., 1] * s, xf[..., 1] * c + xf[..., 0] * s], dim=-1)
    return out.flatten(3).to(x.dtype)


@dataclass
class KVState:
    """Mirror of zonos/config.py:8-52 reduced to what the path reads/writes."""
    kv: list                      # per layer [R, S, 2, Hkv, hd]
    lengths: torch.Tensor         # int32 [R]   (lengths_per_sample)
    seqlen_offset: int = 0
    max_seqlen: int = 0


class TransformerOracle:
    def __init__(self, weights: dict, dims: BackboneDims, dtype=torch.bfloat16):
        self.d = dims
        self.dtype = dtype
        self.w = {k: v.to(dtype) for k, v in weights.items()
                  if k.startswith(("backbone.", "embeddings.", "fused_heads."))}
        self.freqs = rotary_table(16384, dims.head_dim)  # _torch.py:206

    # ---- cache -------------------------------------------------------------
    def allocate(self, rows: int, max_seqlen: int) -> KVState:
        """model.py:333-338 (+ find_multiple(.,8)) and _torch.py:305."""
        S = (max_seqlen + 7) // 8 * 8
        d = self.d
        kv = [torch.zeros(rows, S, 2, d.n_heads_kv, d.head_dim, dtype=self.dtype) for _ in range(d.n_layer)]
        return KVState(kv=kv, lengths=torch.zeros(rows, dtype=torch.int32), seqlen_offset=0, max_seqlen=S)

    # ---- pieces ------------------------------------------------------------
    def embed(self, codes: torch.Tensor) -> torch.Tensor:
        """codec_utils.py:37: Python `sum` => ((0 + E0) + E1) + ... sequential, rounding each add."""
        acc = 0
        for k in range(self.d.n_codebooks):
            acc = acc + F.embedding(codes[:, k], self.w[f"embeddings.{k}.weight"])
        return acc  # [B, T, D]

    def _ln(self, x, prefix):
        return F.layer_norm(x, (self.d.d_model,), self.w[prefix + ".weight"], self.w[prefix + ".bias"], self.d.norm_eps)

    def _attention(self, x, st: KVState, li: int, cs):
        d = self.d
        R, T, _ = x.shape
        p = f"backbone.layers.{li}.mixer."
        qkv = F.linear(x, self.w[p + "in_proj.weight"])
        qs, ks = d.n_heads * d.head_dim, d.n_heads_kv * d.head_dim
        q, k, v = qkv.split([qs, ks, ks], dim=-1)                      # _torch.py:401
        q = apply_rope(q.view(R, T, d.n_heads, d.head_dim), cs)
        k = apply_rope(k.view(R, T, d.n_heads_kv, d.head_dim), cs)
        v = v.view(R, T, d.n_heads_kv, d.head_dim)
        s0 = st.seqlen_offset
        st.kv[li][:R, s0:s0 + T, 0] = k                                # _torch.py:105-106
        st.kv[li][:R, s0:s0 + T, 1] = v
        kk = st.kv[li][:R, :s0 + T, 0].transpose(1, 2)                 # [R,Hkv,S,hd]
        vv = st.kv[li][:R, :s0 + T, 1].transpose(1, 2)
        y = F.scaled_dot_product_attention(q.transpose(1, 2), kk, vv, is_causal=T > 1, enable_gqa=True)
        y = y.transpose(1, 2).contiguous().view(R, T, qs)
        for _ in range(d.out_proj_repeats):                            # _torch.py:419-420 (twice)
            y = F.linear(y, self.w[p + "out_proj.weight"])
        return y

    def _mlp(self, x, li: int):
        p = f"backbone.layers.{li}.mlp."
        y, gate = F.linear(x, self.w[p + "fc1.weight"]).chunk(2, dim=-1)   # value first, gate second
        return F.linear(y * F.silu(gate), self.w[p + "fc2.weight"])

    def forward(self, hidden: torch.Tensor, st: KVState, taps: dict | None = None) -> torch.Tensor:
        """_torch.py:232-238.  Does NOT advance st (the generate loop does, model.py:430-431)."""
        R, T, _ = hidden.shape
        pos = torch.arange(T).unsqueeze(0) + st.lengths[:R].long().unsqueeze(1)
        cs = self.freqs[pos]                                           # [R,T,hd/2,2]
        x = hidden
        for li in range(self.d.n_layer):
            x = x + self._attention(self._ln(x, f"backbone.layers.{li}.norm"), st, li, cs)
            x = x + self._mlp(self._ln(x, f"backbone.layers.{li}.norm2"), li)
            if taps is not None:
                taps[li] = x.clone()
        return self._ln(x, "backbone.norm_f")

    def logits(self, hidden: torch.Tensor, st: KVState, cfg_scale: float) -> torch.Tensor:
        """model.py:225-234: last position -> fused heads [R,Q,V] -> fp32 -> u + (c-u)*s."""
        d = self.d
        last = self.forward(hidden, st)[:, -1, :]
        out = F.linear(last, self.w["fused_heads.weight"]).view(-1, d.n_codebooks, d.head_vocab).float()
        if cfg_scale != 1.0:
            c, u = out.chunk(2)
            out = u + (c - u) * cfg_scale
        return out
