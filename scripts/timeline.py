"""Device-side timeline (globaltimer stamps of CTA 0) of consecutive PDL-chained launches of one decode GEMV."""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig, _lib, transformer_config_dict
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights
dev = torch.device("cuda:0")
w = make_backbone_weights(**TRANSFORMER_DIMS, seed=0)
m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**TRANSFORMER_DIMS))).to(dev, torch.bfloat16); m.load_state_dict(w)
ctx = m._ctx(); native = m._native_model(); sp = _lib.stream_ptr(dev)
lib = C.CDLL(_lib.LIB_PATH); lib.zb_debug_timeline.argtypes = [C.c_void_p]
names = ["start", "copy0", "copyN", "dep_ok", "x_ready", "stage0", "drained", "end"]
for which, name in ((1, "out_proj"), (2, "fc1"), (3, "fc2")):
    iters = 12
    buf = torch.zeros(iters * 8, dtype=torch.int64, device=dev)
    ctx.check(ctx.lib.zb_bench_kernel(ctx.handle, native, 0, which, 2, 26, sp)); torch.cuda.synchronize()
    lib.zb_debug_timeline(C.c_void_p(buf.data_ptr()))
    ctx.check(ctx.lib.zb_bench_kernel(ctx.handle, native, 0, which, 2, iters, sp)); torch.cuda.synchronize()
    lib.zb_debug_timeline(C.c_void_p(0))
    t = buf.view(iters, 8).cpu()
    t0 = int(t[4, 0])
    print(f"--- {name}: per launch, ns relative to launch 4 start (CTA 0): " + " ".join(names))
    for i in range(4, 10):
        print(f"  launch {i}: " + " ".join(f"{int(v) - t0:7d}" for v in t[i]))
