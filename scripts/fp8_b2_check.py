"""Full-size FP8 mode at batch 2 (R = 4: 16 KB stages, the bf16 row geometry): FP8 kernel against the bf16 kernel on the
dequantised weights, per-call logits while the sampled histories agree (the same contract as tests/test_gpu_parity.py:
test_fp8_mode_full_size, which runs batch 1)."""
import os, sys, time, torch
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_b200_model, fp8_dequantised, q_stream_from_seed
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning
dims = dict(TRANSFORMER_DIMS, n_layer=int(os.environ.get("ZB_CHECK_LAYERS", "26")))
w = fp8_dequantised(make_backbone_weights(**dims, seed=0))
model = build_b200_model(dims, w, "cuda:0")
B, N = 2, 8
cond = make_conditioning(2 * B, 40, dims["d_model"], seed=4).to("cuda:0")
q = q_stream_from_seed(5, N + 9, B)
out = {}
for mode in ("0", "1"):
    os.environ["ZB_FP8"] = mode
    tr = {}
    model.generate(cond, max_new_tokens=N, batch_size=B, q_stream=q, trace=tr)
    out[mode] = (tr["logits"].float().cpu(), tr["delayed"].cpu())
(l0, d0), (l1, d1) = out["0"], out["1"]
alive = torch.ones(B, dtype=torch.bool); worst = 0.0; n = 0
for call in range(1, min(l0.shape[0], d0.shape[-1] - 1)):
    for b in range(B):
        if alive[b]:
            fin = torch.isfinite(l0[call, b])
            r = float(((l1[call, b][fin] - l0[call, b][fin]).abs() / (0.06 + 0.01 * l0[call, b][fin].abs())).max())
            worst = max(worst, r); n += 1
    alive &= (d1[..., 1 + call] == d0[..., 1 + call]).all(dim=1)
print(f"fp8 batch 2 full size: prefill equal {bool(torch.equal(l0[0], l1[0]))}, {n} (call, utterance) pairs compared, worst |diff| / (0.06 + 0.01 |b|) = {worst:.3f} (<= 1 passes), {time.time() - t0:.1f} s")
