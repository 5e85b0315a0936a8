import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import _lib
dev = torch.device("cuda:0")
ctx = _lib.context(dev)
lib = C.CDLL(_lib.LIB_PATH)
lib.zb_debug_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
g = torch.Generator(device="cpu").manual_seed(0)
ok = True
for (M, N, K) in [(16, 128, 64), (16, 128, 256), (8, 256, 512), (100, 384, 2048), (322, 3072, 2048), (322, 2048, 8192), (20, 9225, 512), (256, 512, 1024), (300, 130, 128)]:
    x = (torch.randn(M, K, generator=g)).bfloat16().to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(dev)
    y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    st = lib.zb_debug_gemm(ctx.handle, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(y.data_ptr()), M, N, K, _lib.stream_ptr(dev))
    if st != 0:
        print("launch failed", ctx.lib.zb_last_error(ctx.handle).decode()); ok = False; continue
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().T)
    err = (y.float() - ref).abs().max().item()
    print(f"M={M} N={N} K={K} max err {err:.4f} ref absmax {ref.abs().max().item():.2f}")
    ok = ok and err < 0.05
print("GEMM_TC", "OK" if ok else "FAILED")
