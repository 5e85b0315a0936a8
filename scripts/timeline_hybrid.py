"""Timeline of the persistent decode kernel on the hybrid stack (CTA 0): per phase of one Mamba2 layer, wait for inputs + work."""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig, _lib
from zonos_b200.config import hybrid_config_dict
from zonos_b200.synthetic import make_hybrid_weights, make_conditioning
dev = torch.device("cuda:0")
w = make_hybrid_weights(seed=0, heads_scale=8.0)
w["fused_heads.weight"][1024] = 0
m = Zonos(ZonosConfig.from_dict(hybrid_config_dict())).to(dev, torch.bfloat16); m.load_state_dict(w)
lib = C.CDLL(_lib.LIB_PATH); lib.zb_debug_timeline.argtypes = [C.c_void_p]
cond = make_conditioning(2, 160, 2048).to(dev)
m.generate(cond, max_new_tokens=40, seed=1)
buf = torch.zeros(512, dtype=torch.int64, device=dev)
lib.zb_debug_timeline(C.c_void_p(buf.data_ptr()))
m.generate(cond, max_new_tokens=int(os.environ.get("ZB_TL_N", "400")), seed=1)
lib.zb_debug_timeline(C.c_void_p(0))
t = buf.cpu().tolist()
print("embed", t[1] - t[0], "ns")
i = 2
for layer in range(9):                        # layers 0..8 are Mamba2 layers (attention at 9, 18, ...)
    row = []
    prev = t[i - 1]
    for ph in ("in_proj", "scan", "out_proj"):
        ready, done = t[i], t[i + 1]; i += 2
        row.append(f"{ph}: wait {ready - prev} work {done - ready}")
        prev = done
    if layer in (2, 3, 4):
        print(f"layer {layer}: " + "   ".join(row) + "  (ns)")
print("layer period:", t[2 + 6 * 4 + 1] - t[2 + 6 * 3 + 1], "ns")
