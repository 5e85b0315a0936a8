"""DAC decode alone: random codes [B, 9, T] -> waveform, timed with CUDA events (and under ncu for the launch list).

  python scripts/dac_times.py --batch 8 --frames 861 [--reps 3] [--cudnn]
--cudnn additionally times transformers' DacModel.decode under bf16 autocast on the same GPU (the cuDNN bar).
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=861)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cudnn", action="store_true")
    args = ap.parse_args()
    from zonos_b200 import DACAutoencoder
    from zonos_b200.synthetic import make_dac_weights
    dev = torch.device("cuda:0")
    dacw = make_dac_weights(seed=1)
    ae = DACAutoencoder(dacw, device=dev)
    codes = torch.randint(0, 1024, (args.batch, 9, args.frames), generator=torch.Generator().manual_seed(3)).to(dev)
    wav = ae.decode(codes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        wav = ae.decode(codes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    frames = args.batch * args.frames
    tf = frames * 1.608e9 / (ms * 1e-3) / 1e12
    print(f"zonos_b200 DAC decode B={args.batch} T={args.frames}: {ms:.2f} ms  {ms * 1e3 / frames:.2f} us/frame  {tf:.0f} TFLOP/s  wav {tuple(wav.shape)}")
    if args.cudnn:
        from transformers.models.dac import DacConfig, DacModel
        dac = DacModel(DacConfig(sampling_rate=44100)).eval().requires_grad_(False)
        dac.load_state_dict(dacw, strict=False)
        dac = dac.to(dev)
        with torch.autocast("cuda", torch.bfloat16), torch.no_grad():
            ref = dac.decode(audio_codes=codes).audio_values
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                ref = dac.decode(audio_codes=codes).audio_values
            e1.record()
            torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / args.reps
        print(f"transformers DacModel.decode (cuDNN, bf16 autocast) same box: {ms2:.2f} ms  {frames * 1.608e9 / (ms2 * 1e-3) / 1e12:.0f} TFLOP/s  "
              f"max |diff| {float((ref.float().view(-1) - wav.view(-1)[: ref.numel()]).abs().max()):.4f}")


if __name__ == "__main__":
    main()
