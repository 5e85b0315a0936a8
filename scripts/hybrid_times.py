"""Full-size hybrid (assumed shape: configs/zonos_v0.1_hybrid.json) generate timing, batch 1, 10 s."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig
from zonos_b200.config import hybrid_config_dict
from zonos_b200.synthetic import make_hybrid_weights, make_conditioning
dev = torch.device("cuda:0")
w = make_hybrid_weights(seed=0, heads_scale=8.0)
m = Zonos(ZonosConfig.from_dict(hybrid_config_dict())).to(dev, torch.bfloat16); m.load_state_dict(w)
cond = make_conditioning(2, 164, 2048).to(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 861
m.generate(cond, max_new_tokens=32, seed=1)
out = []
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = m.generate(cond, max_new_tokens=N, seed=10 + i)
    torch.cuda.synchronize(); out.append((round((time.perf_counter() - t0) * 1e3, 1), c.shape[-1]))
nbytes = sum(v.numel() * v.element_size() for k, v in w.items() if k.startswith(("backbone.", "fused_heads")))
print("hybrid generate ms / frames:", out, " weights streamed per step: %.2f GB" % (nbytes / 1e9))
