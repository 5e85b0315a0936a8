"""Distribution of generate() time (batch 1, 861 frames) over repeated calls: looks for intermittent slow runs."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import Zonos, ZonosConfig, transformer_config_dict
from zonos_b200.synthetic import TRANSFORMER_DIMS, make_backbone_weights, make_conditioning
dev = torch.device("cuda:0")
w = make_backbone_weights(**TRANSFORMER_DIMS, seed=0, heads_scale=8.0, eos_off=True)
m = Zonos(ZonosConfig.from_dict(transformer_config_dict(**TRANSFORMER_DIMS))).to(dev, torch.bfloat16); m.load_state_dict(w)
cond = make_conditioning(2, 160, 2048).to(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 861
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
m.generate(cond, max_new_tokens=64, seed=1)
import ctypes as C, numpy as np
from zonos_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH); lib.zb_debug_steplog.argtypes = [C.c_void_p]
log = torch.zeros(16400 + 8 * 4001, dtype=torch.int64, device=dev)
lib.zb_debug_steplog(C.c_void_p(log.data_ptr()))
out = []
for i in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    c = m.generate(cond, max_new_tokens=N, seed=10 + i)
    e1.record()
    torch.cuda.synchronize()
    out.append((round(e0.elapsed_time(e1), 1), round((time.perf_counter() - t0) * 1e3, 1), c.shape[-1]))
    full = log.cpu().numpy()
    t = full[:8200].reshape(-1, 2)[1:N - 2]
    ts = full[8200:16400].reshape(-1, 2)[1:N - 2]
    fine = full[16400:16400 + 8 * 4001].reshape(-1, 8)[1:N - 2]
    print("   sampler stages after start (p50 us): logits+penalty %.1f softmax %.1f min-p %.1f race %.1f syncthreads %.1f bookkeeping %.1f fence+atomic %.1f end %.1f" % tuple([np.median(fine[:, k] - ts[:, 0]) / 1e3 for k in range(7)] + [np.median(ts[:, 1] - ts[:, 0]) / 1e3]))
    print(f"   sampler: kernel end -> sampler start p50 {np.median(ts[:, 0] - t[:, 1]) / 1e3:.1f} us, sampler run p50 {np.median(ts[:, 1] - ts[:, 0]) / 1e3:.1f} us, sampler end -> next kernel start p50 {np.median(t[1:, 0] - ts[:-1, 1]) / 1e3:.1f} us")
    dur = (t[:, 1] - t[:, 0]) / 1e3
    gap = (t[1:, 0] - t[:-1, 1]) / 1e3
    print(f"run {i}: {out[-1][0]} ms; step kernel us: p50 {np.median(dur):.0f} p90 {np.percentile(dur, 90):.0f} max {dur.max():.0f}; gap between steps us: p50 {np.median(gap):.0f} p90 {np.percentile(gap, 90):.0f} max {gap.max():.0f}; first/last quarter kernel p50 {np.median(dur[:N//4]):.0f}/{np.median(dur[-N//4:]):.0f}")
print("generate ms (cuda events, wall, frames):", out)
