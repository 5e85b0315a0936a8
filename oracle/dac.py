"""Oracle: DAC 44.1 kHz decode (codes -> waveform) and encode (waveform -> codes), CPU torch.

Restates what `zonos/autoencoder.py:119-140` reaches in the un-vendored
third-party dependency `transformers` (requirements.txt:30 `>=4.48.1`; 5.5.0
installed in the build container):
  transformers/models/dac/modeling_dac.py:345-369  quantizer.from_codes
  :85-99    Snake1d      x + (alpha + 1e-9)^-1 * sin(alpha x)^2
  :173-207  residual unit Snake -> Conv1d(k7, dil d, pad 3d) -> Snake -> Conv1d(k1) -> + input
  :234-262  decoder block Snake -> ConvTranspose1d(k=2s, stride s, pad ceil(s/2)) -> 3 res units (dil 1,3,9)
  :405-439  decoder       Conv1d(1024->1536,k7,p3) -> 4 blocks (s=8,8,4,2) -> Snake -> Conv1d(96->1,k7,p3) -> tanh
Encode (`zonos/autoencoder.py:80-117` -> `DacModel.encode`, fp32, no autocast):
  :442-473  encoder       Conv1d(1->64,k7,p3) -> 4 blocks (s=2,4,8,8) -> Snake -> Conv1d(1024->1024,k3,p1)
  :210-231  encoder block 3 res units (dil 1,3,9) -> Snake -> Conv1d(k=2s, stride s, pad ceil(s/2)), channels doubled
  :281-343  residual VQ   per codebook: in_proj (k1, 1024->8) -> nearest L2-normalised code (:154-170) -> out_proj of
            `p + (q - p)` (:146-147, the straight-through expression is kept: it rounds) -> residual -= that
The reference itself holds no test or golden vector for this; parity is pinned
by outputs of `transformers.DacModel.decode` generated in the build container
(tests/golden/make_golden.py) on seeded random-init weights.

Weights: flat dict keyed like DacModel.state_dict()
(`quantizer.quantizers.{k}.codebook.weight`, `...out_proj.{weight,bias}`,
`decoder.conv1.*`, `decoder.block.{i}.snake1.alpha`, `decoder.block.{i}.conv_t1.*`,
`decoder.block.{i}.res_unit{j}.{snake1,snake2}.alpha`, `...{conv1,conv2}.*`,
`decoder.snake1.alpha`, `decoder.conv2.*`).
"""
import math
import torch
import torch.nn.functional as F

STRIDES = (8, 8, 4, 2)
DILATIONS = (1, 3, 9)


def snake(x, alpha):
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def from_codes(w: dict, codes: torch.Tensor) -> torch.Tensor:
    """codes int64 [B,Q,T] -> fp32 [B,1024,T]: sum_k Conv1d_k1(Embedding_k[codes_k])."""
    z = 0.0
    for k in range(codes.shape[1]):
        p = f"quantizer.quantizers.{k}."
        lat = F.embedding(codes[:, k], w[p + "codebook.weight"]).transpose(1, 2)   # [B,8,T]
        z = z + F.conv1d(lat, w[p + "out_proj.weight"], w[p + "out_proj.bias"])
    return z


def _r(x):
    """Round to bf16 and back: one rounding point of the reference's CUDA autocast path."""
    return x.bfloat16().float()


def decode(w: dict, codes: torch.Tensor, taps: dict | None = None, autocast_bf16: bool = False) -> torch.Tensor:
    """codes int64 [B,Q,T] -> fp32 [B,1,512*T]  (autoencoder.py:140 adds the channel dim).

    autocast_bf16=False is the reference's CPU path (autocast disabled, autoencoder.py:138 -> full fp32).
    autocast_bf16=True emulates the dtype flow of its CUDA path (`torch.autocast(cuda, bf16)`):
    conv operands and results rounded to bf16, fp32 accumulation, Snake evaluated in fp32 on the bf16 activation,
    bf16 residual adds.  (The final tanh and the codebook-projection sum are left in fp32.)"""
    w = {k: v.float() for k, v in w.items() if k.startswith(("quantizer.", "decoder."))}
    if autocast_bf16:
        return _decode_autocast(w, codes)
    x = from_codes(w, codes)
    x = F.conv1d(x, w["decoder.conv1.weight"], w["decoder.conv1.bias"], padding=3)
    if taps is not None:
        taps["conv1"] = x.clone()
    for i, s in enumerate(STRIDES):
        p = f"decoder.block.{i}."
        x = snake(x, w[p + "snake1.alpha"])
        x = F.conv_transpose1d(x, w[p + "conv_t1.weight"], w[p + "conv_t1.bias"], stride=s,
                               padding=math.ceil(s / 2))
        for j, dil in enumerate(DILATIONS, start=1):
            r = p + f"res_unit{j}."
            y = F.conv1d(snake(x, w[r + "snake1.alpha"]), w[r + "conv1.weight"], w[r + "conv1.bias"],
                         dilation=dil, padding=3 * dil)
            y = F.conv1d(snake(y, w[r + "snake2.alpha"]), w[r + "conv2.weight"], w[r + "conv2.bias"])
            x = x + y
        if taps is not None:
            taps[f"block{i}"] = x.clone()
    x = snake(x, w["decoder.snake1.alpha"])
    x = F.conv1d(x, w["decoder.conv2.weight"], w["decoder.conv2.bias"], padding=3)
    return torch.tanh(x)


def _decode_autocast(w: dict, codes: torch.Tensor) -> torch.Tensor:
    def conv(x, name, **kw):
        return _r(F.conv1d(_r(x), _r(w[name + ".weight"]), None, **kw) + w[name + ".bias"].view(1, -1, 1))

    x = _r(from_codes(w, codes))
    x = conv(x, "decoder.conv1", padding=3)
    for i, s in enumerate(STRIDES):
        p = f"decoder.block.{i}."
        x = snake(x, w[p + "snake1.alpha"])
        x = _r(F.conv_transpose1d(_r(x), _r(w[p + "conv_t1.weight"]), None, stride=s, padding=math.ceil(s / 2))
               + w[p + "conv_t1.bias"].view(1, -1, 1))
        for j, dil in enumerate(DILATIONS, start=1):
            r = p + f"res_unit{j}."
            y = conv(snake(x, w[r + "snake1.alpha"]), r + "conv1", dilation=dil, padding=3 * dil)
            y = conv(snake(y, w[r + "snake2.alpha"]), r + "conv2")
            x = _r(x + y)
    x = snake(x, w["decoder.snake1.alpha"])
    return torch.tanh(conv(x, "decoder.conv2", padding=3))


ENC_STRIDES = (2, 4, 8, 8)


def encode_latents(w: dict, wav: torch.Tensor) -> torch.Tensor:
    """wav fp32 [B,1,L] (L a multiple of 512) -> encoder output fp32 [B,1024,L/512] (modeling_dac.py:463-473)."""
    w = {k: v.float() for k, v in w.items() if k.startswith("encoder.")}
    x = F.conv1d(wav.float(), w["encoder.conv1.weight"], w["encoder.conv1.bias"], padding=3)
    for i, s in enumerate(ENC_STRIDES):
        p = f"encoder.block.{i}."
        for j, dil in enumerate(DILATIONS, start=1):
            r = p + f"res_unit{j}."
            y = F.conv1d(snake(x, w[r + "snake1.alpha"]), w[r + "conv1.weight"], w[r + "conv1.bias"], dilation=dil, padding=3 * dil)
            y = F.conv1d(snake(y, w[r + "snake2.alpha"]), w[r + "conv2.weight"], w[r + "conv2.bias"])
            x = x + y
        x = F.conv1d(snake(x, w[p + "snake1.alpha"]), w[p + "conv1.weight"], w[p + "conv1.bias"], stride=s, padding=math.ceil(s / 2))
    x = snake(x, w["encoder.snake1.alpha"])
    return F.conv1d(x, w["encoder.conv2.weight"], w["encoder.conv2.bias"], padding=1)


def quantize(w: dict, z: torch.Tensor, margins: list | None = None) -> torch.Tensor:
    """Residual VQ (modeling_dac.py:281-343, :122-170): z fp32 [B,1024,T] -> codes int64 [B,Q,T].  `margins` receives, per
    codebook, best minus second-best score per position (a small margin marks a decision other fp32 summation orders may flip)."""
    w = {k: v.float() for k, v in w.items() if k.startswith("quantizer.")}
    B, _, T = z.shape
    residual, codes, k = z, [], 0
    while f"quantizer.quantizers.{k}.codebook.weight" in w:
        p = f"quantizer.quantizers.{k}."
        proj = F.conv1d(residual, w[p + "in_proj.weight"], w[p + "in_proj.bias"])            # [B,8,T]
        enc = F.normalize(proj.permute(0, 2, 1).reshape(B * T, -1))
        cb = F.normalize(w[p + "codebook.weight"])
        l2 = enc.pow(2).sum(1, keepdim=True)
        dist = -(l2 - 2 * enc @ cb.t()) + cb.pow(2).sum(1, keepdim=True).t()               # (sic: + |c|^2; constant, codes are unit vectors)
        idx = dist.max(1)[1].reshape(B, T)
        if margins is not None:
            top2 = dist.topk(2, dim=1).values
            margins.append((top2[:, 0] - top2[:, 1]).reshape(B, T))
        q = F.embedding(idx, w[p + "codebook.weight"]).transpose(1, 2)
        q = proj + (q - proj)                                                               # straight-through expression, forward value
        q = F.conv1d(q, w[p + "out_proj.weight"], w[p + "out_proj.bias"])
        residual = residual - q
        codes.append(idx)
        k += 1
    return torch.stack(codes, dim=1)


def encode(w: dict, wav: torch.Tensor) -> torch.Tensor:
    """wav fp32 [B,1,L] -> codes int64 [B,9,L/512] (`DacModel.encode(...).audio_codes`, modeling_dac.py:581-640)."""
    return quantize(w, encode_latents(w, wav))
