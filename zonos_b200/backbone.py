"""Backbone plugin for the reference's registry (`zonos/backbone/__init__.py:24-36`).

`B200ZonosBackbone` keeps the class contract the reference's `Zonos` consumes
(`zonos/model.py:75,228,336`): `supported_architectures`, `__init__(BackboneConfig)`,
`allocate_inference_cache(batch_size, max_seqlen, dtype) -> dict`, `forward(hidden[R,T,D], InferenceParams)`,
and parameter names `layers.{i}.norm|mixer.in_proj|mixer.out_proj|norm2|mlp.fc1|mlp.fc2`, `norm_f`
(`zonos/backbone/_torch.py:278-281,369-370,470-471,154-155`), so reference checkpoints load unchanged.
All arithmetic runs in libzonos_b200.so; the nn.Modules below only OWN the weights.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .config import BackboneConfig, InferenceParams

ROPE_TABLE_LEN = 16384   # zonos/backbone/_torch.py:206


def rotary_table(seq_len: int, head_dim: int, base: float = 10000.0, device=None) -> torch.Tensor:
    """fp32 [seq_len, head_dim/2, 2] = (cos, sin)(pos * base^(-2i/head_dim)) as zonos/backbone/_torch.py:29-34."""
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2, device=device)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(seq_len, device=device), inv)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()


class _Mixer(nn.Module):
    def __init__(self, d_model, n_heads, n_heads_kv, head_dim):
        super().__init__()
        self.in_proj = nn.Linear(d_model, (n_heads + 2 * n_heads_kv) * head_dim, bias=False)
        self.out_proj = nn.Linear(n_heads * head_dim, d_model, bias=False)


class _MLP(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.fc1 = nn.Linear(d_model, 2 * d_ff, bias=False)
        self.fc2 = nn.Linear(d_ff, d_model, bias=False)


class _Weight(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))


class _Mamba2Mixer(nn.Module):
    """Parameter holder with mamba_ssm's Mamba2 names and shapes (defaults of ssm_cfg={"layer": "Mamba2"}:
    d_state 128, d_conv 4, expand 2, headdim 64, ngroups 1)."""

    def __init__(self, d_model, ssm_cfg: dict):
        super().__init__()
        self.d_state = ssm_cfg.get("d_state", 128)
        self.d_conv = ssm_cfg.get("d_conv", 4)
        self.headdim = ssm_cfg.get("headdim", 64)
        self.ngroups = ssm_cfg.get("ngroups", 1)
        self.d_inner = ssm_cfg.get("expand", 2) * d_model
        self.nheads = self.d_inner // self.headdim
        conv_dim = self.d_inner + 2 * self.ngroups * self.d_state
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner + 2 * self.ngroups * self.d_state + self.nheads, bias=False)
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, self.d_conv, groups=conv_dim, padding=self.d_conv - 1, bias=True)
        self.dt_bias = nn.Parameter(torch.zeros(self.nheads))
        self.A_log = nn.Parameter(torch.zeros(self.nheads))
        self.D = nn.Parameter(torch.ones(self.nheads))
        self.norm = _Weight(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)


def _make_norm(cfg: BackboneConfig, rms: bool) -> nn.Module:
    """Parameter holder of a block norm: LayerNorm (weight + bias) or, for hybrid checkpoints trained with rms_norm=true,
    mamba_ssm's RMSNorm (weight only - such a state dict has no `.bias` keys)."""
    return _Weight(cfg.d_model) if rms else nn.LayerNorm(cfg.d_model, eps=cfg.norm_epsilon)


def _bias_ptr(norm: nn.Module):
    b = getattr(norm, "bias", None)
    return b.data_ptr() if b is not None else None


class _Block(nn.Module):
    def __init__(self, cfg: BackboneConfig, is_attention: bool = True, rms: bool = False):
        super().__init__()
        self.is_attention = is_attention
        self.norm = _make_norm(cfg, rms)
        if is_attention:
            heads, heads_kv = cfg.attn_cfg["num_heads"], cfg.attn_cfg["num_heads_kv"]
            self.mixer = _Mixer(cfg.d_model, heads, heads_kv, cfg.d_model // heads)
            self.norm2 = _make_norm(cfg, rms)
            self.mlp = _MLP(cfg.d_model, cfg.attn_mlp_d_intermediate)
        else:
            assert cfg.d_intermediate == 0, "Mamba2 layers with an MLP (d_intermediate > 0) are not supported"
            self.mixer = _Mamba2Mixer(cfg.d_model, cfg.ssm_cfg)


class PagedKVCache:
    """Caller-owned paged KV memory in the layout include/zonos_b200.h documents."""

    def __init__(self, n_attn_layers, rows, max_seqlen, n_heads_kv, head_dim, device, dtype=torch.bfloat16, mamba=None):
        P = _lib.PAGE_TOKENS
        self.rows = rows
        self.pages_per_row = (max_seqlen + P - 1) // P
        self.num_pages = rows * self.pages_per_row
        self.kv_pages = torch.empty(max(n_attn_layers, 1), self.num_pages, 2, n_heads_kv, P, head_dim, dtype=dtype, device=device)
        # Mamba2 layers: rolling conv window [n_mamba, R, conv_dim, d_conv] and SSM state [n_mamba, R, nheads, headdim, d_state],
        # zero-initialised like mamba_ssm's allocate_inference_cache
        self.conv_state = self.ssm_state = None
        if mamba is not None:
            n_mamba, conv_dim, d_conv, nheads, headdim, d_state = mamba
            self.conv_state = torch.zeros(n_mamba, rows, conv_dim, d_conv, dtype=dtype, device=device)
            self.ssm_state = torch.zeros(n_mamba, rows, nheads, headdim, d_state, dtype=dtype, device=device)
        # simplest allocation policy: row r owns pages [r*ppr, (r+1)*ppr); the kernels only ever go through the table
        self.page_table = torch.arange(self.num_pages, dtype=torch.int32, device=device).view(rows, self.pages_per_row).contiguous()

    def desc(self, lengths: torch.Tensor) -> _lib.zb_cache:
        assert lengths.dtype == torch.int32 and lengths.is_cuda and lengths.numel() >= self.rows
        d = _lib.zb_cache()
        d.rows, d.num_pages, d.max_pages_per_row = self.rows, self.num_pages, self.pages_per_row
        d.kv_pages, d.page_table, d.lengths = self.kv_pages.data_ptr(), self.page_table.data_ptr(), lengths.data_ptr()
        d.conv_state = self.conv_state.data_ptr() if self.conv_state is not None else None
        d.ssm_state = self.ssm_state.data_ptr() if self.ssm_state is not None else None
        return d


class B200ZonosBackbone(nn.Module):
    supported_architectures = ["transformer", "hybrid"]

    def __init__(self, config: BackboneConfig):
        super().__init__()
        self.config = config
        self.hybrid = bool(config.ssm_cfg)
        if self.hybrid:
            assert config.ssm_cfg.get("layer", "Mamba2") == "Mamba2", "only Mamba2 SSM layers are supported"
        attn = set(config.attn_layer_idx) if self.hybrid else set(range(config.n_layer))
        # zonos/backbone/_torch.py ignores rms_norm (LayerNorm always); the mamba_ssm blocks of the hybrid
        # variant honour it (zonos/backbone/_mamba_ssm.py:18-40 -> create_block(rms_norm=...))
        rms = self.hybrid and bool(config.rms_norm)
        self.layers = nn.ModuleList(_Block(config, i in attn, rms) for i in range(config.n_layer))
        self.norm_f = _make_norm(config, rms)
        # Reference quirks kept as switches.  Transformer variant = zonos/backbone/_torch.py: out_proj applied
        # twice, interleaved-pair RoPE with an fp32 table, LayerNorm.  Hybrid variant = mamba_ssm blocks
        # (zonos/backbone/_mamba_ssm.py): out_proj once, rotate-half RoPE with cos/sin cached in bf16, norm per `rms_norm`.
        self.out_proj_repeats = 1 if self.hybrid else 2
        self.rope_interleaved = not self.hybrid
        self._handles = {}        # key -> (zb_model handle, keepalive)
        self._versions = {}       # key -> parameter versions the handle has seen
        self._tables = None       # (embeddings, heads, n_codebooks, head_vocab) registered by the owning Zonos
        self._rope = None
        self._cache: PagedKVCache | None = None

    # ---- native model handle ------------------------------------------------------------------
    def _weights_key(self):
        return tuple(p.data_ptr() for p in self.parameters()) + (self.out_proj_repeats, self.rope_interleaved)

    def register_tables(self, embeddings, heads, n_codebooks=9, head_vocab=1025):
        """The owner of the embedding tables / fused heads (`Zonos`) registers them once, so that the plugin-level
        `forward()` and the model-level calls share ONE native handle instead of re-creating it back and forth."""
        self._tables = (list(embeddings), heads, n_codebooks, head_vocab)

    def native_model(self, embeddings=None, heads=None, n_codebooks=9, head_vocab=1025):
        """zb_model for the current weights (+ the registered or given embedding tables / fused heads).  Handles are
        cached per (weights, tables) key and only released with the module: a live generate session keeps a pointer to
        its zb_model, so a handle must never be destroyed because another call shape came along."""
        p0 = next(self.parameters())
        if p0.device.type != "cuda" or p0.dtype != torch.bfloat16:
            raise RuntimeError("B200ZonosBackbone needs bf16 weights on a CUDA device (model.to(device, torch.bfloat16))")
        if embeddings is None and heads is None and self._tables is not None:
            embeddings, heads, n_codebooks, head_vocab = self._tables
        key = self._weights_key() + tuple(e.data_ptr() for e in (embeddings or [])) + ((heads.data_ptr(),) if heads is not None else ())
        hit = self._handles.get(key)
        # torch bumps a tensor's version on every in-place write (load_state_dict, optimizer steps): same pointers, new
        # values - the library must rebuild what it derived from the weights (zb_model_weights_changed)
        versions = tuple(p._version for p in self.parameters()) + tuple(e._version for e in (embeddings or [])) + \
            ((heads._version,) if heads is not None else ())
        if hit is not None:
            if self._versions.get(key) != versions:
                _lib.load().zb_model_weights_changed(hit[0])
                self._versions[key] = versions
            return hit[0]
        cfg = self.config
        ctx = _lib.context(p0.device)
        H, Hkv = cfg.attn_cfg["num_heads"], cfg.attn_cfg["num_heads_kv"]
        hd = cfg.d_model // H
        if self._rope is None or self._rope.device != p0.device:
            rope = rotary_table(ROPE_TABLE_LEN, hd, device=p0.device)
            if self.hybrid:      # flash_attn RotaryEmbedding keeps cos/sin in the activation dtype
                rope = rope.to(torch.bfloat16).float().contiguous()
            self._rope = rope
        rope = self._rope
        layers = (_lib.zb_layer * cfg.n_layer)()
        mm = None
        for i, blk in enumerate(self.layers):
            L = layers[i]
            L.kind = 0 if blk.is_attention else 1
            L.norm_w, L.norm_b = blk.norm.weight.data_ptr(), _bias_ptr(blk.norm)
            L.in_proj, L.out_proj = blk.mixer.in_proj.weight.data_ptr(), blk.mixer.out_proj.weight.data_ptr()
            if blk.is_attention:
                L.norm2_w, L.norm2_b = blk.norm2.weight.data_ptr(), _bias_ptr(blk.norm2)
                L.fc1, L.fc2 = blk.mlp.fc1.weight.data_ptr(), blk.mlp.fc2.weight.data_ptr()
            else:
                mm = blk.mixer
                L.conv_w, L.conv_b = mm.conv1d.weight.data_ptr(), mm.conv1d.bias.data_ptr()
                L.dt_bias, L.A_log, L.D = mm.dt_bias.data_ptr(), mm.A_log.data_ptr(), mm.D.data_ptr()
                L.mnorm_w = mm.norm.weight.data_ptr()
        d = _lib.zb_model_desc()
        d.d_model, d.n_layer, d.n_heads, d.n_heads_kv, d.head_dim = cfg.d_model, cfg.n_layer, H, Hkv, hd
        d.d_ff = cfg.attn_mlp_d_intermediate
        d.n_codebooks, d.head_vocab = n_codebooks, head_vocab
        d.emb_vocab = embeddings[0].shape[0] if embeddings else 0
        d.norm_kind = 1 if (self.hybrid and cfg.rms_norm) else 0        # _torch.py ignores rms_norm (LayerNorm always)
        if mm is not None:
            d.d_inner, d.d_state, d.d_conv, d.m_headdim, d.m_ngroups = mm.d_inner, mm.d_state, mm.d_conv, mm.headdim, mm.ngroups
        d.rope_interleaved, d.out_proj_repeats = int(self.rope_interleaved), self.out_proj_repeats
        d.norm_eps, d.rope_len, d.rope_table = cfg.norm_epsilon, ROPE_TABLE_LEN, rope.data_ptr()
        d.layers = layers
        d.norm_f_w, d.norm_f_b = self.norm_f.weight.data_ptr(), _bias_ptr(self.norm_f)
        emb_arr = None
        if embeddings:
            emb_arr = (C.c_void_p * len(embeddings))(*[e.data_ptr() for e in embeddings])
            d.embeddings = emb_arr
        d.heads = heads.data_ptr() if heads is not None else None
        h = C.c_void_p()
        with ctx.lock:
            ctx.check(ctx.lib.zb_model_create(ctx.handle, C.byref(d), C.byref(h)))
        self._handles[key] = (h, (rope, layers, emb_arr))
        self._versions[key] = versions
        return h

    def __del__(self):
        try:
            lib = _lib.load()
            for h, _ in getattr(self, "_handles", {}).values():
                lib.zb_model_destroy(h)
        except Exception:
            pass

    # ---- reference plugin contract ---------------------------------------------------------------
    def allocate_inference_cache(self, batch_size: int, max_seqlen: int, dtype: torch.dtype = torch.bfloat16):
        """zonos/backbone/_torch.py:157-211: returns {layer: (cache, aux)}; here cache = that layer's KV pages and
        aux = the page table.  Allocation follows the current default device like the reference (`with torch.device`)."""
        assert dtype == torch.bfloat16, "the B200 path computes in bf16"
        cfg = self.config
        H, Hkv = cfg.attn_cfg["num_heads"], cfg.attn_cfg["num_heads_kv"]
        device = next(self.parameters()).device
        attn_layers = [i for i, b in enumerate(self.layers) if b.is_attention]
        mamba_layers = [i for i, b in enumerate(self.layers) if not b.is_attention]
        mamba = None
        if mamba_layers:
            mm = self.layers[mamba_layers[0]].mixer
            mamba = (len(mamba_layers), mm.d_inner + 2 * mm.ngroups * mm.d_state, mm.d_conv, mm.nheads, mm.headdim, mm.d_state)
        self._cache = PagedKVCache(len(attn_layers), batch_size, max_seqlen, Hkv, cfg.d_model // H, device, dtype, mamba=mamba)
        out = {i: (self._cache.kv_pages[a], self._cache.page_table) for a, i in enumerate(attn_layers)}
        out.update({i: (self._cache.conv_state[m], self._cache.ssm_state[m]) for m, i in enumerate(mamba_layers)})
        return out

    def forward(self, hidden_states: torch.Tensor, inference_params: InferenceParams, last_only: bool = False) -> torch.Tensor:
        """zonos/backbone/_torch.py:213-238.  Positions come from `inference_params.lengths_per_sample`."""
        assert self._cache is not None, "call allocate_inference_cache first (Zonos.setup_cache does)"
        R, T, D = hidden_states.shape
        assert R == self._cache.rows, f"cache was allocated for {self._cache.rows} rows, got {R}"
        x = hidden_states.contiguous()
        assert x.dtype == torch.bfloat16 and x.is_cuda
        lengths = inference_params.lengths_per_sample
        if not lengths.is_cuda:
            raise RuntimeError("lengths_per_sample must live on the model's device")
        assert lengths.dtype == torch.int32 and lengths.numel() >= R
        c = self._cache
        return torch.ops.zonos_b200.backbone_forward(self.native_model().value, x, c.kv_pages, c.page_table, lengths, c.conv_state,
                                                     c.ssm_state, bool(last_only))


# Same registry shape as zonos/backbone/__init__.py:24-36.  This framework has one implementation, so every
# name the reference knows resolves to it.
BACKBONES = {"b200": B200ZonosBackbone, "torch": B200ZonosBackbone}
