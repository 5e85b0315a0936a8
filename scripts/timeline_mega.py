"""Timeline of the persistent decode kernel (CTA 0): per phase, time from "inputs ready" to "work done" and the wait
for the next phase's inputs.  ZB_TL_FF=<d_ff> overrides the MLP width (debug experiments)."""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig, _lib, transformer_config_dict
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning
dev = torch.device("cuda:0")
dims = dict(TRANSFORMER_DIMS)
if os.environ.get("ZB_TL_FF"):
    dims["d_ff"] = int(os.environ["ZB_TL_FF"])
w = make_backbone_weights(**dims, seed=0, heads_scale=8.0, eos_off=True)
m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**dims))).to(dev, torch.bfloat16); m.load_state_dict(w)
lib = C.CDLL(_lib.LIB_PATH); lib.zb_debug_timeline.argtypes = [C.c_void_p]
B = int(os.environ.get("ZB_TL_B", "1"))
cond = make_conditioning(2 * B, 160, 2048).to(dev)
m.generate(cond, max_new_tokens=40, batch_size=B, seed=1)
buf = torch.zeros(512, dtype=torch.int64, device=dev)
lib.zb_debug_timeline(C.c_void_p(buf.data_ptr()))
m.generate(cond, max_new_tokens=int(os.environ.get("ZB_TL_N", "400")), batch_size=B, seed=1)   # the buffer keeps the stamps of the LAST step (kv_len ~ 170 + N)
lib.zb_debug_timeline(C.c_void_p(0))
t = buf.cpu().tolist()
names = ["in_proj", "attn", "out1", "out2", "fc1", "fc2"]
print("embed", t[1] - t[0], "ns; wait", t[2] - t[1])
i = 2
for layer in range(3):
    row = []
    for ph in names:
        work = t[i + 1] - t[i]; bar = t[i + 2] - t[i + 1]; i += 2
        row.append(f"{ph}:{work}+{bar}")
    print(f"layer {layer}: " + "  ".join(row) + "   (work ns + wait ns)")
print("layer period (layer1 start -> layer2 start):", t[2 + 24] - t[2 + 12], "ns")
for name, off in (("in_proj", 200), ("fc1", 220)):
    x = t[off:off + 9]
    if x[0] > 0:
        lab = ["inputs ready", "stats issued", "norm params in smem", "LN barrier", "normalized", "first stage", "stages done", "final barrier", "epilogue done"]
        print(f"layer 1 {name}: " + "  ".join(f"{l} +{v - x[0]}" for l, v in zip(lab[1:], x[1:])))
