"""Tolerance of the FP8 weight mode over a long utterance, on the CPU oracle (no GPU), TEACHER-FORCED: the bf16 model generates
(seeded draws, generate defaults); a second model whose decode steps use the dequantised e4m3 weights (tests/helpers.py:
fp8_dequantised; the prefill uses the original weights, as in the FP8 mode) is fed the SAME tokens step by step and keeps its own
KV cache.  Per step: logit error, argmax (greedy-token) agreement per codebook, KL(bf16 || fp8).  The GPU kernels reproduce the
oracle's first-step figure to 1 % (profiles/r2_fp8_oracle_prediction.txt), which is what makes this CPU study meaningful.
Weights as in bench.py (random init, heads x8, codebook-0 EOS pinned off).

  python scripts/fp8_oracle_teacher_forced.py [steps=500] > profiles/r2_fp8_oracle_teacher_forced.txt
"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import fp8_dequantised, oracle_dims
from oracle.codebook import apply_delay_pattern
from oracle.generate import make_logit_bias
from oracle.sampling import sample_from_logits
from oracle.transformer import TransformerOracle
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
t0 = time.time()
torch.set_num_threads(os.cpu_count())
dims = dict(TRANSFORMER_DIMS, n_layer=int(os.environ.get("ZB_CHECK_LAYERS", "26")))
w = make_backbone_weights(**dims, seed=0, heads_scale=8.0, eos_off=True)
ref = TransformerOracle(w, oracle_dims(dims), torch.bfloat16)
fp8 = TransformerOracle(fp8_dequantised(w), oracle_dims(dims), torch.bfloat16)
B, Lc, Q, N = 1, 160, 9, steps
cond = make_conditioning(2 * B, Lc, dims["d_model"], seed=1234)
torch.manual_seed(420)
st_r, st_f = ref.allocate(2 * B, Lc + N + Q), fp8.allocate(2 * B, Lc + N + Q)
codes = torch.full((B, Q, N), -1, dtype=torch.int64)
delayed = torch.from_numpy(apply_delay_pattern(codes.numpy(), 1025))
ids = delayed[..., :1].repeat(2, 1, 1)
hidden = torch.cat([cond.to(torch.bfloat16), ref.embed(ids)], dim=1)
logits = ref.logits(hidden, st_r, 2.0)                                    # prefill with the original weights ...
for li in range(len(st_f.kv)):
    st_f.kv[li].copy_(st_r.kv[li])                                        # ... whose cache both models start from
tok = sample_from_logits(logits, min_p=0.1)
delayed[..., 1] = torch.where(delayed[..., 1] == -1, tok, delayed[..., 1])
for st in (st_r, st_f):
    st.seqlen_offset += Lc + 1
    st.lengths += Lc + 1
bias = make_logit_bias(B, Q, 1025, 1024)
offset = 1
acc = dict(se=0.0, n=0, mx=0.0, agree=0, rows=0, kl=0.0, sq=0.0)
per100 = []
for step in range(steps):
    offset += 1
    if offset >= delayed.shape[2]:
        break
    ids = delayed[..., offset - 1:offset].repeat(2, 1, 1)
    lr = ref.logits(ref.embed(ids), st_r, 2.0) + bias
    lf = fp8.logits(fp8.embed(ids), st_f, 2.0) + bias
    fin = torch.isfinite(lr) & torch.isfinite(lf)
    e = torch.where(fin, lf - lr, torch.zeros_like(lr))
    m_r, m_f = torch.where(fin, lr, torch.full_like(lr, -1e30)), torch.where(fin, lf, torch.full_like(lf, -1e30))
    p_r, p_f = torch.log_softmax(m_r, -1), torch.log_softmax(m_f, -1)
    kl = (p_r.exp() * torch.where(fin, p_r - p_f, torch.zeros_like(p_r))).sum(-1)
    ag = (m_r.argmax(-1) == m_f.argmax(-1))
    acc["se"] += float(e.pow(2).sum()); acc["n"] += int(fin.sum()); acc["mx"] = max(acc["mx"], float(e.abs().max()))
    acc["agree"] += int(ag.sum()); acc["rows"] += ag.numel(); acc["kl"] += float(kl.sum())
    acc["sq"] += float(torch.where(fin, lr, torch.zeros_like(lr)).pow(2).sum())
    window = delayed[..., max(0, offset - min(N, 100)):offset]
    tok = sample_from_logits(lr, generated_tokens=window, min_p=0.1)      # the bf16 model's token drives BOTH models
    delayed[..., offset] = torch.where(delayed[..., offset] == -1, tok, delayed[..., offset])
    for st in (st_r, st_f):
        st.seqlen_offset += 1
        st.lengths += 1
    if (step + 1) % 100 == 0:
        per100.append((step + 1, (acc["se"] / acc["n"]) ** 0.5, acc["agree"] / acc["rows"], acc["kl"] / acc["rows"]))
        print(f"  after {step + 1:4d} steps: logit rms err {per100[-1][1]:.4f}, argmax agreement {per100[-1][2]:.4f}, mean KL {per100[-1][3]:.5f}  ({time.time() - t0:.0f} s)", flush=True)
print(f"oracle, {dims['n_layer']} layers, teacher-forced over {step + 1} decode steps x {Q} codebooks = {acc['rows']} rows: logit rms err {(acc['se'] / acc['n']) ** 0.5:.4f} "
      f"(rms of the bf16 logits {(acc['sq'] / acc['n']) ** 0.5:.3f}), max |err| {acc['mx']:.3f}, greedy-token (argmax) agreement {acc['agree'] / acc['rows']:.4f}, "
      f"mean KL(bf16 || fp8) {acc['kl'] / acc['rows']:.5f} nat")
print("GPU (B200, bench.py fp8_batch1.tolerance, first decode step after 16 conditionings, 144 rows): logit rms err 0.179 (rms 5.32), max 0.75, argmax agreement 0.854, mean KL 0.0149")
