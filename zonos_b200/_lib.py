"""ctypes binding of libzonos_b200.so (include/zonos_b200.h).  No torch types cross this boundary:
tensors are passed as raw device pointers (`tensor.data_ptr()`) plus sizes, streams as `cudaStream_t`.

There is NO fallback: if the shared library is missing or no B200 is present, loading / context creation
raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` or `make -C zonos_b200/csrc`.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzonos_b200.so")
ABI_VERSION = 1
PAGE_TOKENS = 64

c_void_p, c_int32, c_int64, c_uint64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class zb_layer(C.Structure):
    _fields_ = [("kind", c_int32), ("_pad", c_int32)] + [(n, c_void_p) for n in (
        "norm_w", "norm_b", "in_proj", "out_proj", "norm2_w", "norm2_b", "fc1", "fc2",
        "conv_w", "conv_b", "dt_bias", "A_log", "D", "mnorm_w")]


class zb_model_desc(C.Structure):
    _fields_ = [(n, c_int32) for n in ("d_model", "n_layer", "n_heads", "n_heads_kv", "head_dim", "d_ff", "n_codebooks",
                                       "head_vocab", "emb_vocab", "norm_kind", "rope_interleaved", "out_proj_repeats")] + [
        ("norm_eps", c_float), ("rope_len", c_int32), ("rope_table", c_void_p), ("layers", C.POINTER(zb_layer)),
        ("norm_f_w", c_void_p), ("norm_f_b", c_void_p), ("embeddings", C.POINTER(c_void_p)), ("heads", c_void_p)] + [
        (n, c_int32) for n in ("d_inner", "d_state", "d_conv", "m_headdim", "m_ngroups", "_pad")]


class zb_cache(C.Structure):
    _fields_ = [("rows", c_int32), ("num_pages", c_int32), ("max_pages_per_row", c_int32), ("_pad", c_int32),
                ("kv_pages", c_void_p), ("page_table", c_void_p), ("lengths", c_void_p), ("conv_state", c_void_p),
                ("ssm_state", c_void_p)]


class zb_sampling(C.Structure):
    _fields_ = [(n, c_float) for n in ("temperature", "top_p", "min_p", "linear", "conf", "quad", "repetition_penalty")] + [
        ("top_k", c_int32), ("repetition_penalty_window", c_int32)]


class zb_gen_desc(C.Structure):
    _fields_ = [(n, c_int32) for n in ("B", "Q", "T_delayed", "prefix_audio_len", "cond_len", "max_new_tokens")] + [
        ("delayed", c_void_p), ("prefix_conditioning", c_void_p), ("cfg_scale", c_float), ("sampling", zb_sampling),
        ("q_stream", c_void_p), ("q_calls", c_int32), ("_pad", c_int32), ("seed", c_uint64), ("logits_trace", c_void_p),
        ("trace_calls", c_int32), ("_pad2", c_int32)]


class zb_gen_progress(C.Structure):
    _fields_ = [(n, c_int32) for n in ("done", "offset", "steps", "max_steps")]


class zb_dac_desc(C.Structure):
    _fields_ = [(n, c_int32) for n in ("n_codebooks", "codebook_size", "codebook_dim", "latent_dim", "channels", "n_blocks")] + [
        ("strides", c_int32 * 8), ("tensors", C.POINTER(c_void_p)), ("n_tensors", c_int32), ("_pad", c_int32)]


class zb_dac_enc_desc(C.Structure):
    _fields_ = [(n, c_int32) for n in ("n_codebooks", "codebook_size", "codebook_dim", "latent_dim", "hidden", "n_blocks")] + [
        ("strides", c_int32 * 8), ("tensors", C.POINTER(c_void_p)), ("n_tensors", c_int32), ("_pad", c_int32)]


# name -> (restype, argtypes); every symbol declared in include/zonos_b200.h
PROTOTYPES = {
    "zb_abi_version": (c_int32, []),
    "zb_ctx_create": (c_int32, [c_int32, C.POINTER(c_void_p)]),
    "zb_ctx_destroy": (c_int32, [c_void_p]),
    "zb_last_error": (C.c_char_p, [c_void_p]),
    "zb_launch_count": (c_int64, [c_void_p]),
    "zb_model_create": (c_int32, [c_void_p, C.POINTER(zb_model_desc), C.POINTER(c_void_p)]),
    "zb_model_destroy": (c_int32, [c_void_p]),
    "zb_model_weights_changed": (c_int32, [c_void_p]),
    "zb_embed_codes": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "zb_backbone_forward": (c_int32, [c_void_p, c_void_p, C.POINTER(zb_cache), c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "zb_heads_cfg": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p]),
    "zb_sample_from_logits": (c_int32, [c_void_p, C.POINTER(zb_sampling), c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64,
                                        c_int64, c_int32, c_void_p, c_uint64, c_uint64, c_int32, c_void_p, c_void_p]),
    "zb_generate_begin": (c_int32, [c_void_p, c_void_p, C.POINTER(zb_cache), C.POINTER(zb_gen_desc), C.POINTER(c_void_p), c_void_p]),
    "zb_generate_steps": (c_int32, [c_void_p, c_int32, c_void_p]),
    "zb_generate_poll": (c_int32, [c_void_p, C.POINTER(zb_gen_progress), c_void_p]),
    "zb_generate_peek": (c_int32, [c_void_p, C.POINTER(zb_gen_progress)]),
    "zb_generate_end": (c_int32, [c_void_p]),
    "zb_dac_create": (c_int32, [c_void_p, C.POINTER(zb_dac_desc), C.POINTER(c_void_p), c_void_p]),
    "zb_dac_destroy": (c_int32, [c_void_p]),
    "zb_dac_decode": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "zb_dac_encode_workspace_bytes": (C.c_size_t, [c_int32, c_int64]),
    "zb_dac_encode": (c_int32, [c_void_p, C.POINTER(zb_dac_enc_desc), c_void_p, c_int32, c_int64, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "zb_bench_kernel": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
}

_lib = None
_lock = threading.Lock()
_contexts: dict = {}


def load():
    """dlopen the library and bind every prototype.  Raises if it is absent (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: the zonos_b200 CUDA library has not been built "
                        "(run `make -C zonos_b200/csrc` or `__graft_entry__.build()`); there is no CPU fallback")
                lib = C.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(lib, name)
                    fn.restype, fn.argtypes = res, args
                if lib.zb_abi_version() != ABI_VERSION:
                    raise ImportError(f"libzonos_b200.so ABI {lib.zb_abi_version()} != binding ABI {ABI_VERSION}")
                _lib = lib
    return _lib


class Context:
    """One zb_ctx per CUDA device (calls on it are serialised with a lock)."""

    def __init__(self, device_index: int):
        self.lib = load()
        h = c_void_p()
        st = self.lib.zb_ctx_create(device_index, C.byref(h))
        if st != 0:
            raise RuntimeError(self.lib.zb_last_error(None).decode())
        self.handle = h
        self.device_index = device_index
        self.lock = threading.RLock()

    def check(self, status: int):
        if status != 0:
            raise RuntimeError(self.lib.zb_last_error(self.handle).decode())

    def launch_count(self) -> int:
        return int(self.lib.zb_launch_count(self.handle))


def context(device) -> Context:
    """Context for a torch.device / index; created on first use."""
    import torch
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"zonos_b200 runs on CUDA (B200) only, got device '{dev}'; there is no CPU path")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _lock:
        ctx = _contexts.get(idx)
    if ctx is None:
        with torch.cuda.device(idx):
            new = Context(idx)
        with _lock:
            ctx = _contexts.setdefault(idx, new)
    return ctx


def stream_ptr(device=None) -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(None)
