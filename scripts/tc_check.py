"""GPU check of the tcgen05 persistent decode step (decode_tc.cu) against the other CUDA decode paths and, at tiny
dims, against the CPU oracle.  Prints one line per configuration:

  python scripts/tc_check.py --dims tiny --batch 1 8 64 --frames 12 --oracle
  python scripts/tc_check.py --dims full --batch 1 --frames 861 --time

Both CUDA paths run inside one process (ZB_DECODE_TC is read per generate session) on the same explicit Exp(1) draws.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def history_aware_diff(t1, t0, P=0):
    """Per sample call: max |logits difference| over the utterances whose delayed-code history still equals the other
    run's, and how many utterances have parted ways (a float near-tie in one sampler decision legitimately forks one)."""
    n = min(int(t1["steps"]), int(t0["steps"])) + 1
    l1 = t1["logits"][:n].cpu()
    l0 = t0["logits"][:n]
    l0 = (torch.stack(list(l0)) if isinstance(l0, list) else l0).cpu()
    d1, d0 = t1["delayed"].cpu(), t0["delayed"].cpu()
    B = d1.shape[0]
    alive = torch.ones(B, dtype=torch.bool)
    worst, forks = [], []
    for call in range(n):
        fin = torch.isfinite(l0[call])
        diff = torch.where(fin, (l1[call] - l0[call]).abs(), torch.zeros_like(l0[call])).flatten(1).max(dim=1).values
        worst.append(float(diff[alive].max()) if alive.any() else 0.0)
        col = P + 1 + call
        alive &= (d1[..., col] == d0[..., col]).all(dim=1)
        forks.append(int((~alive).sum()))
    return worst, forks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="tiny", choices=["tiny", "full"])
    ap.add_argument("--batch", type=int, nargs="+", default=[1])
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--cond-len", type=int, default=10)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--oracle", action="store_true", help="also compare with the CPU oracle (tiny dims)")
    ap.add_argument("--time", action="store_true", help="time generate() with CUDA events, both paths")
    ap.add_argument("--no-compare", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="print the phase timeline of CTA 0, layer 1 (last step)")
    ap.add_argument("--mega", action="store_true", help="compare the two consumers of the B <= 2 persistent kernel (ZB_MEGA_TC=1 tcgen05 / 0 FFMA2) instead")
    args = ap.parse_args()

    from helpers import build_b200_model, oracle_dims, q_stream_from_seed
    from zonos_b200 import _lib
    from zonos_b200.synthetic import TINY_DIMS, TRANSFORMER_DIMS, make_backbone_weights, make_conditioning
    dims = dict(TINY_DIMS if args.dims == "tiny" else TRANSFORMER_DIMS)
    if args.layers:
        dims["n_layer"] = args.layers
    dev = "cuda:0"
    w = make_backbone_weights(**dims, seed=11, heads_scale=8.0 if args.dims == "full" else 1.0, eos_off=args.dims == "full")
    model = build_b200_model(dims, w, dev)
    oracle = None
    if args.oracle:
        from oracle import generate as o_gen
        from oracle.transformer import TransformerOracle
        oracle = TransformerOracle(w, oracle_dims(dims), torch.bfloat16)

    for B in args.batch:
        N, Lc = args.frames, args.cond_len
        cond = make_conditioning(2 * B, Lc, dims["d_model"], seed=9)
        q = q_stream_from_seed(77, N + 9, B) if N <= 64 else None
        res = {}
        for tc in ((1,) if args.no_compare else (1, 0)):
            if args.mega:
                os.environ["ZB_MEGA_TC"] = "1" if tc else "0"
            else:
                os.environ["ZB_DECODE_TC"] = "2" if tc else "0"
            trace = {}
            try:
                codes = model.generate(cond.to(dev), max_new_tokens=N, batch_size=B, q_stream=q, seed=5, trace=trace if N <= 64 else None)
                torch.cuda.synchronize()
            except Exception as e:                      # keep going: the other configurations still tell something
                print(f"B={B} tc={tc}: FAILED {type(e).__name__}: {e}", flush=True)
                res[tc] = None
                continue
            ms = None
            if args.time:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                model.generate(cond.to(dev), max_new_tokens=N, batch_size=B, q_stream=q, seed=6)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            res[tc] = (codes.cpu(), trace, ms)
        line = f"dims={args.dims} B={B} Lc={Lc} N={N}:"
        if res.get(1) and res.get(0):
            c1, t1, m1 = res[1]
            c0, t0, m0 = res[0]
            if t1 and t0:
                worst, forks = history_aware_diff(t1, t0)
                line += f" tc-vs-other same-history logits max {max(worst):.4f} (call1 {worst[1] if len(worst) > 1 else 0:.4f}) forked {forks[-1]}/{B} nan={int(torch.isnan(t1['logits']).sum())}"
            line += f" tokens_equal={bool(c1.shape == c0.shape and torch.equal(c1, c0))}"
            if m1 is not None:
                line += f" ms tc={m1:.1f} other={m0:.1f} per-step tc={m1 / (N + 8):.3f} other={m0 / (N + 8):.3f}"
        elif res.get(1):
            c1, t1, m1 = res[1]
            line += f" tc only: codes {tuple(c1.shape)}" + (f" ms={m1:.1f} per-step {m1 / (N + 8):.3f}" if m1 is not None else "")
        if oracle is not None and res.get(1):
            otrace = {}
            from oracle import generate as o_gen
            ref = o_gen.generate(oracle, cond, None, N, 2.0, B, dict(min_p=0.1), q_stream=q, trace=otrace)
            worst, forks = history_aware_diff(res[1][1], otrace)
            same = bool(res[1][0].shape == ref.shape and torch.equal(res[1][0], ref))
            line += f" | vs oracle: same-history logits max {max(worst):.4f} per call {[round(v, 3) for v in worst[:8]]} forked {forks[-1]}/{B} codes_equal={same}"
        print(line, flush=True)

    if args.timeline and args.mega:
        import ctypes as C
        lib = C.CDLL(_lib.LIB_PATH)
        lib.zb_debug_timeline.argtypes = [C.c_void_p]
        for tc in (1, 0):
            os.environ["ZB_MEGA_TC"] = str(tc)
            buf = torch.zeros(256, dtype=torch.int64, device=dev)
            lib.zb_debug_timeline(C.c_void_p(buf.data_ptr()))
            B = args.batch[-1]
            cond = make_conditioning(2 * B, args.cond_len, dims["d_model"], seed=9)
            model.generate(cond.to(dev), max_new_tokens=args.frames, batch_size=B, seed=5)
            torch.cuda.synchronize()
            lib.zb_debug_timeline(C.c_void_p(0))
            t = buf.cpu().double()
            # stamps of CTA 0: [0] start, [1] embed done, then per layer and phase (in_proj, attention, out x repeats, fc1, fc2): inputs ready, work done
            names = ["in_proj", "attention", "out1", "out2", "fc1", "fc2"]
            li = 5
            base = 2 + li * 12
            print(f"ZB_MEGA_TC={tc}: CTA 0, layer {li} (us): phase = wait for inputs + work")
            prev = float(t[base - 1])
            tot = 0.0
            for i, nme in enumerate(names):
                ready, done = float(t[base + 2 * i]), float(t[base + 2 * i + 1])
                print(f"   {nme:10s} wait {1e-3 * (ready - prev):6.2f}  work {1e-3 * (done - ready):6.2f}")
                tot += done - prev
                prev = done
            print(f"   layer total {1e-3 * tot:.2f} us")
            if tc:
                for nme, off in (("in_proj", 200), ("out2", 208), ("fc1", 216), ("fc2", 224)):
                    x = t[off:off + 6]
                    lab = ["normalised", "image written + CTA barrier", "accumulator complete", "read back + CTA barrier", "epilogue done"]
                    print(f"   layer 1 {nme:8s}: " + "  ".join(f"{l} +{1e-3 * float(v - x[0]):.2f}" for l, v in zip(lab, x[1:])))
    elif args.timeline:
        import ctypes as C
        lib = _lib.load()
        buf = torch.zeros(148 * 128, dtype=torch.int64, device=dev)
        lib.zb_debug_tc_timeline(C.c_void_p(buf.data_ptr()))
        os.environ["ZB_DECODE_TC"] = "2"
        B = args.batch[-1]
        cond = make_conditioning(2 * B, args.cond_len, dims["d_model"], seed=9)
        model.generate(cond.to(dev), max_new_tokens=args.frames, batch_size=B, seed=5)
        torch.cuda.synchronize()
        t = buf.cpu().view(148, 128).double()
        lib.zb_debug_tc_timeline(C.c_void_p(0))
        names = ["qkv: operand released", "qkv: accumulator done", "qkv: dumped -> barrier", "barrier passed", "qkv epilogue -> barrier", "barrier passed",
                 "attention (+ merge) -> barrier", "barrier passed"]
        for g in ("out1", "out2", "fc1", "fc2"):
            names += [f"{g}: operand released", f"{g}: accumulator done", f"{g}: dumped -> barrier", "barrier passed", f"{g} epilogue -> barrier", "barrier passed"]
        t0 = t[:, 0][t[:, 0] > 0].min()
        print(f"timeline of the middle layer, last step, B={B} (us; min / median / max over CTAs; last CTA):")
        prev = 0.0
        for i, name in enumerate(names):
            col = t[:, i]
            ok = col > 0
            if not ok.any():
                continue
            v = (col[ok] - t0) * 1e-3
            last = int(torch.argmax(torch.where(ok, col, torch.zeros_like(col))))
            print(f"  {i:2d} {name:28s} {float(v.min()):8.2f} {float(v.median()):8.2f} {float(v.max()):8.2f}   cta {last:3d}   (+{float(v.max()) - prev:.2f})")
            prev = float(v.max())


if __name__ == "__main__":
    main()
