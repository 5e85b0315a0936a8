"""Long-utterance check of the persistent decode kernel: 30 s (2580 frames, kv_len up to ~2750: several attention units
per CTA) must reproduce the tokens of the 10 s run with the same seed, stay deterministic and in range."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig, transformer_config_dict
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning
dev = torch.device("cuda:0")
w = make_backbone_weights(**TRANSFORMER_DIMS, seed=0, heads_scale=8.0, eos_off=True)
m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**TRANSFORMER_DIMS))).to(dev, torch.bfloat16); m.load_state_dict(w)
for B in (1, 2):
    cond = make_conditioning(2 * B, 160, 2048).to(dev)
    short = m.generate(cond, max_new_tokens=861, batch_size=B, seed=5)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    long_ = m.generate(cond, max_new_tokens=2580, batch_size=B, seed=5)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    again = m.generate(cond, max_new_tokens=2580, batch_size=B, seed=5)
    n = min(short.shape[2], 840)
    print(f"B={B}: long {tuple(long_.shape)} in {dt:.2f} s ({long_.shape[2] / 86.1328 / dt:.1f}x real time per utterance); "
          f"first {n} frames equal to the 861-frame run: {bool(torch.equal(short[..., :n], long_[..., :n]))}; "
          f"deterministic: {bool(torch.equal(long_, again))}; range [{int(long_.min())}, {int(long_.max())}]")
