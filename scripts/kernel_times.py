"""Time each decode GEMV kind alone (C-side launch loop, CUDA events), cycling layers so weights come from HBM."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig, _lib, transformer_config_dict
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights
dev = torch.device("cuda:0")
w = make_backbone_weights(**TRANSFORMER_DIMS, seed=0)
m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**TRANSFORMER_DIMS))).to(dev, torch.bfloat16); m.load_state_dict(w)
ctx = m._ctx(); native = m._native_model(); sp = _lib.stream_ptr(dev); st = torch.cuda.current_stream(dev)
D, F = 2048, 8192
bytes_ = {1: D * D * 2, 2: 2 * F * D * 2, 3: D * F * 2}
for rows in (2, 4):
    for which, name in ((1, "out_proj"), (2, "fc1"), (3, "fc2")):
        ctx.check(ctx.lib.zb_bench_kernel(ctx.handle, native, 0, which, rows, 26, sp)); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 260
        e0.record(st); ctx.check(ctx.lib.zb_bench_kernel(ctx.handle, native, 0, which, rows, iters, sp)); e1.record(st); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / iters
        print(f"rows={rows} {name:9s} {us:7.2f} us/launch  {bytes_[which] / us / 1e3:7.1f} GB/s")
