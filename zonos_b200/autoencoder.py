"""DAC 44.1 kHz autoencoder front-end with the reference's surface (`zonos/autoencoder.py:49-170`):
`decode(codes[B,9,T]) -> wav[B,1,512*T]` runs in libzonos_b200.so; `sampling_rate`, `num_codebooks`,
`codebook_size` are the attributes callers read.  `encode(wav[B,1,L]) -> codes[B,9,L/512]` (the audio-prefix path, SURVEY.md
8(f)) runs there too when the state dict carries the encoder tensors; `preprocess` is the reference's torchaudio resample."""
import ctypes as C
import math

import torch

from . import _lib

DAC_STRIDES = (8, 8, 4, 2)


def dac_tensor_order(n_codebooks=9, n_blocks=4) -> list[str]:
    """state_dict keys in the order zb_dac_desc.tensors expects (include/zonos_b200.h)."""
    keys = []
    for k in range(n_codebooks):
        p = f"quantizer.quantizers.{k}."
        keys += [p + "codebook.weight", p + "out_proj.weight", p + "out_proj.bias"]
    keys += ["decoder.conv1.weight", "decoder.conv1.bias"]
    for i in range(n_blocks):
        p = f"decoder.block.{i}."
        keys += [p + "snake1.alpha", p + "conv_t1.weight", p + "conv_t1.bias"]
        for j in (1, 2, 3):
            r = p + f"res_unit{j}."
            keys += [r + "snake1.alpha", r + "conv1.weight", r + "conv1.bias", r + "snake2.alpha", r + "conv2.weight", r + "conv2.bias"]
    keys += ["decoder.snake1.alpha", "decoder.conv2.weight", "decoder.conv2.bias"]
    return keys


def dac_encoder_tensor_order(n_codebooks=9, n_blocks=4) -> list[str]:
    """state_dict keys in the order zb_dac_enc_desc.tensors expects (include/zonos_b200.h)."""
    keys = ["encoder.conv1.weight", "encoder.conv1.bias"]
    for i in range(n_blocks):
        p = f"encoder.block.{i}."
        for j in (1, 2, 3):
            r = p + f"res_unit{j}."
            keys += [r + "snake1.alpha", r + "conv1.weight", r + "conv1.bias", r + "snake2.alpha", r + "conv2.weight", r + "conv2.bias"]
        keys += [p + "snake1.alpha", p + "conv1.weight", p + "conv1.bias"]
    keys += ["encoder.snake1.alpha", "encoder.conv2.weight", "encoder.conv2.bias"]
    for k in range(n_codebooks):
        p = f"quantizer.quantizers.{k}."
        keys += [p + "in_proj.weight", p + "in_proj.bias", p + "codebook.weight", p + "out_proj.weight", p + "out_proj.bias"]
    return keys


class DACAutoencoder:
    def __init__(self, state_dict: dict | None = None, device="cuda", dac_module=None):
        """state_dict: DacModel tensors (decode side).  None -> `DacModel.from_pretrained("descript/dac_44khz")`
        exactly like zonos/autoencoder.py:74 (needs network / a local HF cache)."""
        self.dac = dac_module
        if state_dict is None:
            from transformers.models.dac import DacModel
            self.dac = DacModel.from_pretrained("descript/dac_44khz").eval().requires_grad_(False)
            state_dict = self.dac.state_dict()
        self.codebook_size = 1024
        self.num_codebooks = sum(1 for k in state_dict if k.endswith("codebook.weight") and k.startswith("quantizer."))
        self.sampling_rate = 44100
        self.device = torch.device(device)
        self._handle = None
        self._create(state_dict)
        # encode side (fp32 tensors kept on the device; the library borrows them per call)
        ekeys = dac_encoder_tensor_order(self.num_codebooks, len(DAC_STRIDES))
        self._enc_weights = [state_dict[k].detach().to(self.device, torch.float32).contiguous() for k in ekeys] \
            if all(k in state_dict for k in ekeys) else None

    def _create(self, sd: dict):
        ctx = _lib.context(self.device)
        keys = dac_tensor_order(self.num_codebooks, len(DAC_STRIDES))
        tensors = [sd[k].detach().to(self.device, torch.float32).contiguous() for k in keys]
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        d = _lib.zb_dac_desc()
        d.n_codebooks, d.codebook_size, d.codebook_dim = self.num_codebooks, sd[keys[0]].shape[0], sd[keys[0]].shape[1]
        d.latent_dim, d.channels, d.n_blocks = sd[keys[1]].shape[0], sd["decoder.conv1.weight"].shape[0], len(DAC_STRIDES)
        for i, s in enumerate(DAC_STRIDES):
            d.strides[i] = s
        d.tensors, d.n_tensors = arr, len(tensors)
        self.codebook_size = int(d.codebook_size)
        h = C.c_void_p()
        with ctx.lock:
            ctx.check(ctx.lib.zb_dac_create(ctx.handle, C.byref(d), C.byref(h), _lib.stream_ptr(self.device)))
        torch.cuda.synchronize(self.device)      # the source tensors may be freed after this
        self._handle = h

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().zb_dac_destroy(self._handle)
        except Exception:
            pass

    # ---- reference surface ---------------------------------------------------------------------
    def preprocess(self, wav: torch.Tensor, sr: int) -> torch.Tensor:
        """zonos/autoencoder.py:80-100: resample to 44.1 kHz, left-pad to a multiple of 512."""
        import torchaudio
        wav = torchaudio.functional.resample(wav, sr, 44_100)
        left = math.ceil(wav.shape[-1] / 512) * 512 - wav.shape[-1]
        return torch.nn.functional.pad(wav, (left, 0), value=0)

    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        """zonos/autoencoder.py:104-117: preprocessed fp32 wav [B,1,L] (L a multiple of 512) -> int64 codes [B,Q,L/512]."""
        if self._enc_weights is None:
            raise RuntimeError("encode needs the encoder tensors (`encoder.*`, `quantizer.quantizers.*.in_proj.*`) in the state dict "
                               "this DACAutoencoder was built from")
        assert wav.dim() == 3 and wav.shape[1] == 1 and wav.shape[-1] % 512 == 0, "expected [B, 1, L] with L a multiple of 512 (preprocess)"
        return torch.ops.zonos_b200.dac_encode(wav.to(self.device), self._enc_weights, self.num_codebooks)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """zonos/autoencoder.py:119-140: int64 [B,Q,T] -> fp32 [B,1,512*T]."""
        assert codes.dim() == 3 and codes.shape[1] == self.num_codebooks
        codes = codes.to(self.device, torch.int64).contiguous()
        return torch.ops.zonos_b200.dac_decode(self._handle.value, codes, math.prod(DAC_STRIDES))

    def decode_to_int16(self, codes: torch.Tensor):
        """zonos/autoencoder.py:142-170."""
        wav = self.decode(codes)[:, 0]
        return torch.clamp(wav * 32767.0, -32767.0, 32767.0).to(torch.int16).squeeze(0).unsqueeze(1)
