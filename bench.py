#!/usr/bin/env python
"""Benchmark of the Zonos inference hot path on B200 (contract: the driver's prompt, section "Measurement").

A "step" is ONE pass of the hot path over one batch of synthetic utterances: prefill of the conditioning prefix,
the whole autoregressive loop (N = 861 frames = 10 s of audio) and the DAC decode of the result.
Workload at N=1 GPU = BASELINE.json configs[1]: Zonos-v0.1-transformer, bf16, batch 1 with CFG 2.0, Lc = 160
synthetic conditioning tokens.  With --gpus N every rank runs the same per-GPU workload on its own utterances
(request sharding, no collective on the data path) => "scaling": "weak".

The default run measures BOTH halves of BASELINE.json's metric: the headline fields are batch 1 (configs[1]); the
`batch64` object is configs[3] (64 utterances per GPU) measured the same way (value, e2e, roofline, breakdown), and
under torchrun every rank runs both, so the line carries the whole-job aggregate of each.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--frames F] [--cond-len Lc] [--no-batch64]
  python bench.py --impl reference ...      # the UNMODIFIED reference (oracle/_ref, staged by build()) on the host cores;
                                            # falls back to the oracle port when oracle/_ref is absent

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAME_RATE = 44100 / 512          # 86.13 codec frames per audio second
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=861, help="new frames per utterance (861 = 10 s)")
    ap.add_argument("--cond-len", type=int, default=160)
    ap.add_argument("--ref-frames", type=int, default=128, help="frames per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="transformer", choices=["transformer", "hybrid"],
                    help="headline model: Zonos-v0.1-transformer (configs[1]/[3]) or Zonos-v0.1-hybrid (Mamba2 + attention, configs[2]/[4])")
    ap.add_argument("--prefix-frames", type=int, default=0, help="synthetic audio-prefix codes per utterance (configs[4]: 258 = 3 s)")
    ap.add_argument("--no-hybrid", action="store_true", help="skip the hybrid batch-1 leg of the default run")
    ap.add_argument("--no-batch64", action="store_true", help="skip the 64-utterances-per-GPU leg")
    ap.add_argument("--batch64-steps", type=int, default=2)
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the informative reference-torch-on-this-GPU timing")
    ap.add_argument("--no-fp8", action="store_true", help="skip the opt-in FP8 leg of the default run")
    ap.add_argument("--layers", type=int, default=26, help="(debug) fewer layers; invalidates the number")
    return ap.parse_args()


def config_dict(args, n_gpus, batch=None, variant=None):
    batch = args.batch if batch is None else batch
    variant = args.variant if variant is None else variant
    if variant == "hybrid":
        base = "configs[2]" if batch == 1 and not args.prefix_frames else "configs[4]-shaped"
    else:
        base = "configs[1]" if batch == 1 else ("configs[3]" if batch == 64 else "configs[3]-shaped")
    return {"workload": "Zonos-v0.1-%s random-init bf16, CFG 2.0, %d utterance(s)/GPU x %d frames (%.1f s)%s + DAC 44.1 kHz decode"
                        % (variant, batch, args.frames, args.frames / FRAME_RATE,
                           " after a %d-frame audio prefix" % args.prefix_frames if args.prefix_frames else ""),
            "baseline_config": base, "variant": variant, "prefix_frames": args.prefix_frames,
            "batch_per_gpu": batch, "frames": args.frames, "cond_len": args.cond_len, "cfg_scale": 2.0,
            "sampling": "min_p=0.1, repetition_penalty=3.0 (generate defaults)",
            # the hybrid leg runs the assumed shape of configs/zonos_v0.1_hybrid.json (46 layers, attention at every 9th)
            "n_layer": 46 if variant == "hybrid" else args.layers,
            "parallelism": "request-sharded replicas x%d, no data-path collective" % n_gpus,
            "l2": "each decode step streams 3.2 GB of weights (>> 126 MB L2), no flush needed"}


# --------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference): the reference's algorithm restated in oracle/
# --------------------------------------------------------------------------------------------------
def cpu_setup(args):
    """The CPU arm: the unmodified reference from oracle/_ref when build() staged it (kind "reference"), else the oracle
    port (kind "port").  Returns (kind, step_fn(frames) -> (audio_s, wall_s))."""
    import torch
    from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning, make_dac_weights
    from oracle import ref_runner
    dims = dict(TRANSFORMER_DIMS, n_layer=args.layers)
    torch.set_num_threads(os.cpu_count())
    dacw = make_dac_weights(seed=1)
    if args.variant == "hybrid":
        # the reference's hybrid backbone needs mamba_ssm (CUDA-only, not in the image): only the restatement can run here
        from oracle import dac as o_dac, generate as o_gen
        from oracle.hybrid import HybridDims, HybridOracle
        from zonos_b200.synthetic import make_hybrid_weights
        hd = dict(d_model=2048, n_layer=46, attn_layer_idx=(9, 18, 27, 36, 45), n_heads=16, n_heads_kv=4, d_ff=8192)
        hw = make_hybrid_weights(**hd, seed=0, heads_scale=8.0)
        horacle = HybridOracle(hw, HybridDims(**hd), torch.bfloat16)
        hcond = make_conditioning(2, args.cond_len, 2048)

        def hybrid_step(frames):
            torch.manual_seed(420)
            t0 = time.perf_counter()
            codes = o_gen.generate(horacle, hcond, None, frames, 2.0, 1, dict(min_p=0.1))
            o_dac.decode(dacw, codes)
            return codes.shape[2] / FRAME_RATE, time.perf_counter() - t0
        return "port", hybrid_step
    w = make_backbone_weights(**dims, seed=0, heads_scale=8.0, eos_off=True)
    cond = make_conditioning(2, args.cond_len, dims["d_model"])
    if ref_runner.available():
        try:
            model = ref_runner.build_model(dims, w, dacw)
            del w
            return "reference", lambda frames: ref_runner.step(model, cond, frames)
        except Exception as e:                          # a dependency of the reference is missing on this box
            print("reference unavailable, timing the oracle port instead: %r" % (e,), file=sys.stderr)
    from oracle import dac as o_dac, generate as o_gen
    from oracle.transformer import BackboneDims, TransformerOracle
    oracle = TransformerOracle(w, BackboneDims(**dims), torch.bfloat16)

    def port_step(frames):
        torch.manual_seed(420)
        t0 = time.perf_counter()
        codes = o_gen.generate(oracle, cond, None, frames, 2.0, 1, dict(min_p=0.1))
        o_dac.decode(dacw, codes)
        return codes.shape[2] / FRAME_RATE, time.perf_counter() - t0
    return "port", port_step


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (zonos/model.py:354-548 generate +
    zonos/autoencoder.py:119-140 decode, unmodified, from oracle/_ref) with all host threads, same config/metric, each
    step a bounded sample.  Rank 0 only: under torchrun the other ranks exit at once, so for N > 1 the driver's ratio
    compares N GPUs with ONE host process (stated in `note`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    kind, step = cpu_setup(args)
    for _ in range(max(0, min(args.warmup, 1))):
        step(max(4, args.ref_frames // 4))
    audio = t = 0.0
    for _ in range(args.steps):
        a, d = step(args.ref_frames)
        audio += a; t += d
    value = audio / t
    sample = "batch 1, Lc=%d prefill + %d frames + DAC decode per step (of the %d-frame utterance)" % (
        args.cond_len, args.ref_frames, args.frames)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "one host process on rank 0 (the reference's CPU path is single-process); with --gpus N > 1 a ratio against "
                    "this line compares N GPUs with one host"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import torch
    import torch.distributed as dist
    from zonos_b200 import DACAutoencoder, Zonos, ZonosConfig, _lib, transformer_config_dict
    from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning, make_dac_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    dims = dict(TRANSFORMER_DIMS, n_layer=args.layers)
    N, Lc, P = args.frames, args.cond_len, args.prefix_frames

    # every rank builds the same replica (seeded CPU generator) - no weight broadcast needed for synthetic runs
    dacw = make_dac_weights(seed=1, with_encoder=P > 0)        # configs[4]: the audio prefix is ENCODED on the device, too
    autoenc = DACAutoencoder(dacw, device=dev)

    def build(variant):
        """(model, host weights, spec): spec carries what the roofline needs - weight bytes streamed per step, KV bytes per
        cached token and row, recurrent-state bytes read + written per row and step."""
        if variant == "transformer":
            w_ = make_backbone_weights(**dims, seed=0, heads_scale=8.0, eos_off=True)
            m_ = Zonos(ZonosConfig.from_dict(transformer_config_dict(**dims)), autoencoder=autoenc)
            n_attn, n_mamba, d_model, hkv = dims["n_layer"], 0, dims["d_model"], dims["n_heads_kv"]
        else:
            from zonos_b200 import hybrid_config_dict
            from zonos_b200.synthetic import make_hybrid_weights
            hd = dict(d_model=2048, n_layer=46, attn_layer_idx=(9, 18, 27, 36, 45), n_heads=16, n_heads_kv=4, d_ff=8192)
            w_ = make_hybrid_weights(**hd, seed=0, heads_scale=8.0)
            w_["fused_heads.weight"][1024] = 0                 # codebook-0 EOS pinned off like eos_off=True: deterministic length
            m_ = Zonos(ZonosConfig.from_dict(hybrid_config_dict(**hd)), autoencoder=autoenc)
            n_attn, n_mamba, d_model, hkv = len(hd["attn_layer_idx"]), hd["n_layer"] - len(hd["attn_layer_idx"]), hd["d_model"], hd["n_heads_kv"]
        m_ = m_.to(dev, torch.bfloat16)
        m_.load_state_dict(w_)
        w_bytes = 2 * sum(v.numel() for k, v in w_.items() if k.startswith("backbone.") or k.startswith("fused_heads"))
        d_inner = 2 * d_model
        state = n_mamba * 2 * 2 * ((d_inner // 64) * 64 * 128 + (d_inner + 2 * 128) * 4)    # SSM state + conv window, read + write, bf16
        spec = {"variant": variant, "w_bytes": w_bytes, "kv_tok": n_attn * 2 * hkv * 128 * 2, "state_bytes_per_row": state,
                "d_model": d_model, "n_layer": n_attn + n_mamba, "hybrid": variant == "hybrid"}
        return m_, w_, spec

    model, w, spec0 = build(args.variant)
    ctx = model._ctx()
    stream = torch.cuda.current_stream(dev)
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(model, spec, B, steps, warmup, with_clocks):
        """One leg: `steps` passes (prefill + N frames + DAC decode) of B utterances per GPU.  value = inputs resident in
        HBM; e2e = pinned host conditioning (and prefix codes) in, host waveform out, all copies inside the timed region."""
        cond_host = make_conditioning(2 * B, Lc, spec["d_model"], seed=1234 + rank).pin_memory()
        cond_dev = cond_host.to(dev)
        prefix_host = prefix_dev = None
        if P:
            # synthetic prefix AUDIO (P frames = P * 512 samples at 44.1 kHz); value: its codes are resident, e2e: the waveform
            # comes from pinned host memory and is encoded (autoencoder.encode) inside the timed region
            prefix_host = (0.1 * torch.randn(B, 1, P * 512, generator=torch.Generator().manual_seed(7 + rank))).pin_memory()
            prefix_dev = model.autoencoder.encode(prefix_host.to(dev))
        wav_host = torch.empty((B, 1, 512 * (P + N)), dtype=torch.float32).pin_memory()

        def gen(c, pfx, n_frames, seed):
            return model.generate(c, pfx, max_new_tokens=n_frames, cfg_scale=2.0, batch_size=B, sampling_params=dict(min_p=0.1), seed=seed)

        def step_device(seed):
            codes = gen(cond_dev, prefix_dev, N, seed)
            wav = model.autoencoder.decode(codes)
            return codes, wav

        def step_e2e(seed):
            c = cond_host.to(dev, non_blocking=True)                       # H2D inside the timed region
            pfx = model.autoencoder.encode(prefix_host.to(dev, non_blocking=True)) if P else None
            codes = gen(c, pfx, N, seed)
            wav = model.autoencoder.decode(codes)
            wav_host[..., : wav.shape[-1]].copy_(wav, non_blocking=True)   # D2H of the result
            torch.cuda.current_stream(dev).synchronize()
            return codes, wav

        def timed(fn, K):
            barrier()
            l0 = ctx.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            frames = 0
            for i in range(K):
                codes, _ = fn(1000 + i)
                frames += codes.shape[0] * (codes.shape[2] - P)            # NEW audio only: the prefix was given
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            launches = ctx.launch_count() - l0
            barrier()
            t = torch.tensor([ms, float(frames)], dtype=torch.float64, device=dev)
            if world > 1:
                tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
                return float(tmax[0]), float(tsum[1]), launches
            return ms, float(frames), launches

        for i in range(warmup):
            step_device(i)
        clocks = ClockSampler(local)
        if rank == 0 and with_clocks:
            clocks.start()
        ms, frames, launches = timed(step_device, steps)
        clock_info = clocks.stop() if rank == 0 and with_clocks else None
        value = (frames / FRAME_RATE) / (ms / 1e3)
        step_e2e(7)
        ms_e, frames_e, _ = timed(step_e2e, steps)
        e2e = (frames_e / FRAME_RATE) / (ms_e / 1e3)

        # ---- roofline leg.  The dominant kernel is the persistent decode step: one launch per codec frame = embed + all
        # layers + heads, > 90 % of the pass (decode_step_kernel<R> for B <= 2, decode_tc_kernel above).  Its average
        # launch duration is measured live with CUDA events as the slope of generate() time over the number of steps (two
        # lengths, same prefill), which also contains the sampler launch that follows every step.  Algorithmic bytes per
        # launch per SURVEY.md 8(d): W (3.2 GB, out_proj counted once) + KV read at the mean context + KV append. ----
        roof = breakdown = None
        if rank == 0:
            def time_generate(n_frames, reps=2):
                gen(cond_dev, prefix_dev, n_frames, 5)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for r_ in range(reps):
                    c_ = gen(cond_dev, prefix_dev, n_frames, 6 + r_)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                return e0.elapsed_time(e1) / reps, c_

            n_small = max(8, N // 8)
            reps = 2 if B <= 8 else 1
            t_full, codes_full = time_generate(N, reps)
            t_small, _ = time_generate(n_small, reps)
            step_ms = (t_full - t_small) / (N - n_small)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            model.autoencoder.decode(codes_full)
            e0.record(stream)
            for _ in range(3):
                model.autoencoder.decode(codes_full)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            dac_ms = e0.elapsed_time(e1) / 3
            w_bytes, kv_tok = spec["w_bytes"], spec["kv_tok"]
            mean_s = Lc + P + 1 + (n_small + N + 16) / 2
            step_bytes = w_bytes + 2 * B * mean_s * kv_tok + 2 * B * kv_tok + 2 * B * spec["state_bytes_per_row"] \
                + B * (9 * spec["d_model"] * 2 + 9 * 1025 * 4)
            achieved = step_bytes / (step_ms * 1e-3) / 1e9
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath) and not P:
                traffic = json.load(open(tpath)).get(("hybrid_" if spec["hybrid"] else "") + "batch%d_decode_step_dram_bytes_per_launch" % B)
            if spec["hybrid"] and 2 * B <= 4:
                kname = "decode_step_kernel<R=%d> (persistent FFMA2 consumer, hybrid stack: Mamba2 layers as three tagged-word phases - in_proj, conv1d " \
                        "step + selective state update, gated norm + out_proj - next to the attention layers; embed + %d layers + heads, one launch " \
                        "per frame) + sample kernel" % (2 * B, spec["n_layer"])
            elif spec["hybrid"]:
                kname = "decode step of the hybrid stack as ONE CUDA graph of %d layers (gemv3_kernel / gemm_tc_kernel Linears, mamba_scan_kernel, " \
                        "gated_norm_kernel, attn_kernel) + sample kernel, R=%d rows" % (spec["n_layer"], 2 * B)
            elif 2 * B <= 4:
                kname = "decode_step_kernel<R=%d> (persistent FFMA2 consumer: embed + %d layers + heads, one launch per frame) + sample kernel" % (2 * B, spec["n_layer"])
            else:
                kname = "decode_tc_kernel, R=%d rows (persistent tcgen05/TMEM consumer + mma.sync attention: embed + %d layers + heads, one launch per frame) + sample kernel" % (2 * B, spec["n_layer"])
            roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": peak_src, "us_per_launch": step_ms * 1e3, "algorithmic_bytes_per_launch": step_bytes,
                    "how": "slope of generate() time between %d and %d frames (CUDA events)" % (n_small, N)}
            steps_full = N + 8
            breakdown = {"decode_ms": step_ms * steps_full, "prefill_and_setup_ms": max(0.0, t_full - step_ms * steps_full), "dac_ms": dac_ms}
        return {"value": value, "ms_per_step": ms / steps, "frames_per_second": frames / (ms / 1e3), "launches": int(launches),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": cond_host.numel() * 2 + (prefix_host.numel() * 4 if P else 0), "d2h_bytes_per_step": wav_host.numel() * 4,
                        "ms_per_step": ms_e / steps},
                "clocks": clock_info, "roofline": roof, "breakdown_ms": breakdown, "cond_dev": cond_dev}

    B = args.batch
    head = measure(model, spec0, B, args.steps, max(args.warmup, 3), True)
    cond_dev = head.pop("cond_dev")
    roof = head["roofline"]

    if rank == 0 and roof is not None and 2 * B <= 8 and args.variant == "transformer":
        # stand-alone norm2 + fc1 + SiLU GEMV (55 % of the weight bytes) in a C-side launch loop over all layers
        native = model._native_model()
        sp = _lib.stream_ptr(dev)
        Rr = 2 * B
        iters = 8 * dims["n_layer"]
        ctx.check(ctx.lib.zb_bench_kernel(ctx.handle, native, 0, 2, Rr, dims["n_layer"], sp))
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.check(ctx.lib.zb_bench_kernel(ctx.handle, native, 0, 2, Rr, iters, sp))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        fc1_us = 1e3 * e0.elapsed_time(e1) / iters
        fc1_bytes = 2 * dims["d_ff"] * dims["d_model"] * 2 + Rr * dims["d_model"] * 2 + Rr * dims["d_ff"] * 2 + 4 * dims["d_model"]
        roof["gemv_fc1_alone"] = {"kernel": "gemv3_kernel (norm2+fc1+SiLU), PDL-chained launches over all layers",
                                  "us_per_launch": fc1_us, "achieved_gbs": fc1_bytes / (fc1_us * 1e-6) / 1e9,
                                  "frac": fc1_bytes / (fc1_us * 1e-6) / 1e9 / peak}

    default_run_early = B == 1 and args.variant == "transformer" and not P
    # ---- p50 time to first audio: generate_stream() entry -> first 43-frame (0.5 s) chunk decoded and on the host.  The
    # reference has no streaming (its TTFA is the whole generate + decode). ----
    ttfa = None
    if rank == 0:
        samples = []
        for i in range(5):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            gen_it = model.generate_stream(cond_dev, None, max_new_tokens=N, cfg_scale=2.0, batch_size=B, sampling_params=dict(min_p=0.1), seed=40 + i)
            wav0, _ = next(gen_it)
            wav0.cpu()
            samples.append((time.perf_counter() - t0) * 1e3)
            gen_it.close()
            torch.cuda.synchronize(dev)
        samples.sort()
        ttfa = {"p50_ms": samples[len(samples) // 2], "min_ms": samples[0], "max_ms": samples[-1], "chunk_frames": 43,
                "definition": "generate_stream() entry -> first 0.5 s chunk DAC-decoded and copied to the host (prefill + 77 steps + chunk decode)"}
    # ---- the other half of the metric: 64 utterances per GPU (BASELINE.json configs[3]); every rank runs it ----
    batch64 = None
    default_run = B == 1 and args.variant == "transformer" and not P
    if default_run and not args.no_batch64:
        b64 = measure(model, spec0, 64, max(1, args.batch64_steps), 1, False)
        b64.pop("cond_dev")
        if rank == 0:
            batch64 = {"metric": METRIC, "value": b64["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": max(1, args.batch64_steps), "warmup": 1,
                       "ms_per_step": b64["ms_per_step"], "frames_per_second": b64["frames_per_second"], "e2e": b64["e2e"],
                       "gpu_launches": b64["launches"], "roofline": b64["roofline"], "breakdown_ms": b64["breakdown_ms"],
                       "config": config_dict(args, n_gpus, 64), "target_audio_s_per_s_per_box_of_8": 3000.0}

    # ---- BASELINE.json configs[2]: the hybrid (Mamba2 + attention) variant at batch 1; every rank runs it ----
    hybrid = None
    if default_run and not args.no_hybrid:
        hm, hw, hspec = build("hybrid")
        del hw
        hy = measure(hm, hspec, 1, 2, 1, False)
        hy.pop("cond_dev")
        if rank == 0:
            hybrid = {"metric": METRIC, "value": hy["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": 2, "warmup": 1, "ms_per_step": hy["ms_per_step"],
                      "e2e": hy["e2e"], "gpu_launches": hy["launches"], "roofline": hy["roofline"], "breakdown_ms": hy["breakdown_ms"],
                      "config": config_dict(args, n_gpus, 1, "hybrid"),
                      "note": "parity of this variant is pinned by oracle/hybrid.py and transformers' Mamba2Mixer only (the reference's "
                              "hybrid backbone needs mamba_ssm, which is not in the image); layer count / attention positions are the "
                              "assumption of configs/zonos_v0.1_hybrid.json"}
        del hm
        torch.cuda.empty_cache()

    # ---- informative: the reference's own torch path on THIS GPU (eager; SURVEY K1/K2/K16 bars), bounded sample ----
    ref_gpu = None
    if rank == 0 and not args.no_ref_gpu and default_run:
        try:
            from oracle import ref_runner
            if ref_runner.available():
                rm = ref_runner.build_model(dims, w, dacw, device=dev)
                rc = make_conditioning(2, Lc, dims["d_model"], seed=1234).to(dev)
                ref_runner.step(rm, rc, 16)
                n_a, n_b = max(16, args.ref_frames // 4), args.ref_frames
                a0_, d0_ = ref_runner.step(rm, rc, n_a)
                a_, d_ = ref_runner.step(rm, rc, n_b)
                ref_gpu = {"value": a_ / d_, "unit": UNIT, "sample": "batch 1, Lc=%d prefill + %d frames + DAC decode, wall clock" % (Lc, n_b),
                           "ms_per_decode_step": 1e3 * (d_ - d0_) / (n_b - n_a),
                           "how": "value: one whole call (its per-call setup, e.g. CUDA-graph capture, included); ms_per_decode_step: slope "
                                  "between %d and %d frames" % (n_a, n_b),
                           "what": "unmodified reference (zonos torch backbone + transformers DAC) on this GPU, its own code path"}
                del rm
                torch.cuda.empty_cache()
        except Exception as e:
            ref_gpu = {"unavailable": repr(e)[:300]}

    # ---- opt-in FP8 weight mode of the batch-1 decode step (ZB_FP8=1: e4m3 copy with power-of-two row scales,
    # HFMA2 consumer; prefill and DAC stay bf16).  Reported beside the bf16 headline with its own tolerance figures, never mixed in. ----
    fp8 = None
    if rank == 0 and default_run_early and not args.no_fp8:
        try:
            def gen1(n_frames, seed, sp, trace=None):
                return model.generate(cond_dev, None, max_new_tokens=n_frames, cfg_scale=2.0, batch_size=1, sampling_params=sp, seed=seed, trace=trace)

            def timed_gen(n_frames):
                gen1(n_frames, 5, dict(min_p=0.1))
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for r_ in range(2):
                    gen1(n_frames, 6 + r_, dict(min_p=0.1))
                e1.record(stream)
                torch.cuda.synchronize(dev)
                return e0.elapsed_time(e1) / 2

            # Tolerance figures.  Free-running greedy decoding is useless on random-init weights (flat heads: the histories fork in the
            # first frames and then share nothing), so logits are compared on IDENTICAL histories: (a) the same seeded draws in both
            # modes for 500 frames, per-call logits while the sampled histories still agree; (b) the first decode step after 16
            # different conditionings (same prefill, same first token in both modes).
            def masked(l_):
                return torch.where(torch.isfinite(l_), l_, torch.full_like(l_, -1e30))

            def compare(lb_, lf_):
                """rows [..., 1025] on the same history -> (sum sq err, max err, n, argmax agreements, rows, sum KL(bf16 || fp8))"""
                fin_ = torch.isfinite(lb_) & torch.isfinite(lf_)
                e_ = torch.where(fin_, lf_ - lb_, torch.zeros_like(lb_))
                pb_, pf_ = torch.log_softmax(masked(lb_), -1), torch.log_softmax(masked(lf_), -1)
                kl_ = (pb_.exp() * torch.where(fin_, pb_ - pf_, torch.zeros_like(pb_))).sum(-1)
                agree_ = (masked(lb_).argmax(-1) == masked(lf_).argmax(-1)).float()
                return float(e_.pow(2).sum()), float(e_.abs().max()), int(fin_.sum()), float(agree_.sum()), int(agree_.numel()), float(kl_.sum()), \
                    float(torch.where(fin_, lb_, torch.zeros_like(lb_)).pow(2).sum())

            n_tol = min(500, N)
            conds16 = [make_conditioning(2, Lc, spec0["d_model"], seed=5000 + i).to(dev) for i in range(16)]

            def first_steps():
                out_ = []
                for c_ in conds16:
                    t_ = {}
                    model.generate(c_, None, max_new_tokens=2, cfg_scale=2.0, batch_size=1, sampling_params=dict(min_p=0.1), seed=77, trace=t_)
                    out_.append(t_["logits"][:2].float().cpu())
                return torch.stack(out_)                           # [16, 2 calls, 1, 9, 1025]

            tr_b, tr_f = {}, {}
            gen1(n_tol, 3, dict(min_p=0.1), tr_b)
            fs_b = first_steps()
            os.environ["ZB_FP8"] = "1"
            gen1(n_tol, 3, dict(min_p=0.1), tr_f)
            fs_f = first_steps()
            n_small = max(8, N // 8)
            t_full, t_small = timed_gen(N), timed_gen(n_small)
            step_ms = (t_full - t_small) / (N - n_small)
            Lb, Lf = tr_b["logits"].float().cpu(), tr_f["logits"].float().cpu()     # [calls, 1, 9, 1025]
            db, df = tr_b["delayed"].cpu(), tr_f["delayed"].cpu()
            ncall = min(Lb.shape[0], Lf.shape[0], db.shape[-1] - 1)
            same_hist = 1                                          # call c was computed from the columns <= c: equal while they agree
            while same_hist < ncall and bool((db[..., 1:1 + same_hist] == df[..., 1:1 + same_hist]).all()):
                same_hist += 1
            same_hist = max(2, same_hist)
            ra = compare(Lb[1:same_hist], Lf[1:same_hist])
            rb_ = compare(fs_b[:, 1], fs_f[:, 1])
            prefill_same = bool(torch.equal(fs_b[:, 0], fs_f[:, 0]) and torch.equal(Lb[0], Lf[0]))

            def figures(r_):
                return {"rows": r_[4], "logit_rms_err": (r_[0] / max(1, r_[2])) ** 0.5, "logit_max_abs_err": r_[1], "rms_of_bf16_logits": (r_[6] / max(1, r_[2])) ** 0.5,
                        "argmax_agreement": r_[3] / max(1, r_[4]), "mean_kl_bf16_to_fp8": r_[5] / max(1, r_[4])}
            wq_bytes = spec0["w_bytes"] // 2
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            fp8_traffic = json.load(open(tpath)).get("fp8_batch1_decode_step_dram_bytes_per_launch") if os.path.exists(tpath) else None
            mean_s = Lc + 1 + (n_small + N + 16) / 2
            step_bytes = wq_bytes + 2 * mean_s * spec0["kv_tok"] + 2 * spec0["kv_tok"] + 9 * spec0["d_model"] * 2 + 9 * 1025 * 4
            fp8 = {"value": (N / FRAME_RATE) / (t_full / 1e3 + head["breakdown_ms"]["dac_ms"] / 1e3), "unit": UNIT,
                   "us_per_decode_step": step_ms * 1e3, "speedup_of_the_step_vs_bf16": head["roofline"]["us_per_launch"] / (step_ms * 1e3),
                   "roofline": {"bound": "hbm", "kernel": "decode_step_kernel<R=2, FP8 mode> (e4m3 weights, HFMA2 consumer) + sample kernel",
                                "algorithmic_bytes_per_launch": step_bytes, "achieved": step_bytes / (step_ms * 1e-3) / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak, "traffic": fp8_traffic},
                   "tolerance": {"how": "logits of the two modes on identical histories (same prefill: %s)" % prefill_same,
                                 "same_seeded_draws_%d_frames" % n_tol: dict(figures(ra), decode_calls_until_the_histories_fork=same_hist - 1),
                                 "first_decode_step_16_conditionings": figures(rb_)},
                   "what": "weights of the decode step quantised once per model to e4m3 with one power-of-two scale per row (decode.cu: "
                           "quant_e4m3_kernel); prefill, KV cache, activations, DAC unchanged; value = prefill + %d FP8 decode steps + DAC" % (N + 8)}
        except Exception as e:                                   # the bf16 line must not depend on the opt-in mode
            fp8 = {"error": repr(e)[:300]}
        finally:
            os.environ.pop("ZB_FP8", None)
    del cond_dev

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        del w
        kind, cstep = cpu_setup(args)
        a, d = cstep(args.ref_frames)
        cpu = {"value": a / d, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": "batch 1, Lc=%d prefill + %d frames + DAC decode (%.1f s of CPU work)" % (Lc, args.ref_frames, d)}

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config_dict(args, n_gpus), "frames_per_second": head["frames_per_second"],
                "e2e": head["e2e"], "gpu_launches": head["launches"], "clocks": head["clocks"], "roofline": roof,
                "breakdown_ms": head["breakdown_ms"], "ttfa": ttfa, "batch64": batch64, "hybrid_batch1": hybrid, "fp8_batch1": fp8, "reference_gpu_eager": ref_gpu,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
