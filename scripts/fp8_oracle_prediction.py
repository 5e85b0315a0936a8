"""CPU cross-check of the FP8 mode's tolerance figure (no GPU): the oracle (oracle/transformer.py) prefills with the ORIGINAL
weights, then runs the first decode step twice on copies of that cache - once with the original weights, once with the weights
the FP8 mode streams (tests/helpers.py: fp8_dequantised = the dequantised e4m3 copy).  The difference of the two logits tensors
is what e4m3 rounding alone does to that step; the GPU test (tests/test_gpu_parity.py::test_fp8_mode_full_size, same weights,
conditioning and draws) printed for the FP8 KERNEL against the bf16 KERNEL: rms err 0.0240, max 0.0956, logit spread 0.5833.

  python scripts/fp8_oracle_prediction.py > profiles/r2_fp8_oracle_prediction.txt
"""
import copy, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import fp8_dequantised, oracle_dims, q_stream_from_seed
from oracle.sampling import sample_from_logits
from oracle.transformer import TransformerOracle
from oracle.codebook import apply_delay_pattern
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning

t0 = time.time()
torch.set_num_threads(os.cpu_count())
dims = dict(TRANSFORMER_DIMS, n_layer=int(os.environ.get("ZB_CHECK_LAYERS", "26")))
w = make_backbone_weights(**dims, seed=0)
wq = fp8_dequantised(w)
oracle = TransformerOracle(w, oracle_dims(dims), torch.bfloat16)
B, Lc, N, Q = 1, 40, 4, 9
cond = make_conditioning(2 * B, Lc, dims["d_model"], seed=4)
q = q_stream_from_seed(5, N + 9, B)
st = oracle.allocate(2 * B, Lc + N + Q)
codes = torch.full((B, Q, N), -1, dtype=torch.int64)
delayed = torch.from_numpy(apply_delay_pattern(codes.numpy(), 1025))
ids = delayed[..., :1].repeat(2, 1, 1)
hidden = torch.cat([cond.to(torch.bfloat16), oracle.embed(ids)], dim=1)
logits0 = oracle.logits(hidden, st, 2.0)                                  # prefill: original weights in both modes
tok = sample_from_logits(logits0, q=q[0], min_p=0.1)
frame = delayed[..., 1]
delayed[..., 1] = torch.where(frame == -1, tok, frame)
st.seqlen_offset += Lc + 1
st.lengths += Lc + 1
ids = delayed[..., 1:2].repeat(2, 1, 1)
out = {}
for name, weights in (("bf16", w), ("fp8", wq)):
    o = TransformerOracle(weights, oracle_dims(dims), torch.bfloat16)
    s2 = copy.deepcopy(st)
    out[name] = o.logits(o.embed(ids), s2, 2.0)[0]                        # (the logit bias is the same constant in both)
a, b = out["fp8"].float(), out["bf16"].float()
err = a - b
spread = (b - b.mean()).pow(2).mean().sqrt()
print(f"oracle, {dims['n_layer']} layers, first decode step, dequantised-e4m3 weights against the original weights: "
      f"rms err {float(err.pow(2).mean().sqrt()):.4f}, max {float(err.abs().max()):.4f}, logit spread {float(spread):.4f}, "
      f"ratio {float(err.pow(2).mean().sqrt() / spread):.3f}   ({time.time() - t0:.0f} s on {torch.get_num_threads()} threads)")
print("GPU (B200), FP8 kernel against bf16 kernel, same weights / conditioning / draws: rms err 0.0240, max 0.0956, logit spread 0.5833, ratio 0.041")
