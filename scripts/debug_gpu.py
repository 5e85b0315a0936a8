import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import build_b200_model, oracle_dims, load_golden
from oracle.transformer import TransformerOracle
from oracle import dac as o_dac
from zonos_b200.synthetic import TINY_DIMS, make_backbone_weights, make_dac_weights
DEV = "cuda:0"
w = make_backbone_weights(**TINY_DIMS, seed=11)
model = build_b200_model(TINY_DIMS, w, DEV)
oracle = TransformerOracle(w, oracle_dims(TINY_DIMS), torch.bfloat16)
for R in (2, 4):
    for scale in (1.0, 3.0):
        params = model.setup_cache(R, 16); st = oracle.allocate(R, 16)
        g = torch.Generator().manual_seed(1)
        for T in (1, 1, 3):
            x = (torch.randn(R, T, 512, generator=g) * scale).bfloat16()
            got = model.backbone(x.to(DEV), params).float().cpu()
            taps = {}
            ref = oracle.forward(x, st, taps).float()
            e = (got - ref).abs()
            print(f"R={R} scale={scale} T={T} maxerr={e.max():.4f} per-row={[round(v,3) for v in e.amax(dim=(1,2)).tolist()]} ref absmax {ref.abs().max():.2f}")
            params.lengths_per_sample += T; st.seqlen_offset += T; st.lengths += T
# DAC error profile
from zonos_b200 import DACAutoencoder
gd = load_golden("dac_decode.npz")
ae = DACAutoencoder(make_dac_weights(seed=1), device=DEV)
wav = ae.decode(torch.from_numpy(gd["codes"]).to(DEV)).cpu().numpy()
err = np.abs(wav - gd["wav"])[:, 0]
for b in range(2):
    prof = err[b].reshape(10, 512).max(axis=1)
    print("dac err per frame b", b, np.round(prof, 3), "argmax", err[b].argmax(), "ref there", gd["wav"][b,0,err[b].argmax()], "got", wav[b,0,err[b].argmax()])
# oracle with bf16-rounded weights to see how much is weight rounding
wd = make_dac_weights(seed=1)
wd16 = {k: (v.bfloat16().float() if ("conv" in k or "out_proj.weight" in k) and k.endswith("weight") else v) for k, v in wd.items()}
ref16 = o_dac.decode(wd16, torch.from_numpy(gd["codes"])).numpy()
print("fp32-oracle vs bf16-weight-oracle maxdiff", np.abs(ref16 - gd["wav"]).max(), "cuda vs bf16-weight-oracle", np.abs(wav - ref16).max())
