set -u
timeout 600 python -m pytest tests -m gpu -q -x -k "dac" 2>&1 | tail -12
python - <<'PY'
import torch, time, math, sys
sys.path.insert(0,'.')
from zonos_b200 import DACAutoencoder
from zonos_b200.synthetic import make_dac_weights
w = make_dac_weights(seed=1, with_encoder=True)
ae = DACAutoencoder(w, device='cuda:0')
for B, secs in ((1, 3), (8, 3)):
    L = 258 * 512
    wav = 0.2 * torch.randn(B, 1, L, generator=torch.Generator().manual_seed(1)).cuda()
    ae.encode(wav); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = ae.encode(wav); e1.record(); torch.cuda.synchronize()
    print(f"encode B={B} 3 s prefix: {e0.elapsed_time(e1):.1f} ms  ({B * 0.185 / (e0.elapsed_time(e1) * 1e-3):.1f} TFLOP/s fp32)  codes {tuple(c.shape)}")
PY
