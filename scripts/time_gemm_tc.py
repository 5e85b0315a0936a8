import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import _lib
dev = torch.device("cuda:0"); ctx = _lib.context(dev)
lib = C.CDLL(_lib.LIB_PATH)
lib.zb_debug_gemm.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p]
for (M, N, K) in [(128, 3072, 2048), (128, 2048, 2048), (128, 16384, 2048), (128, 2048, 8192), (322, 16384, 2048), (2048, 16384, 2048)]:
    ws = [(torch.randn(N, K) / K ** 0.5).bfloat16().to(dev) for _ in range(8)]      # rotate weights: stream from HBM
    x = torch.randn(M, K).bfloat16().to(dev); y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    sp = _lib.stream_ptr(dev)
    def run(w): lib.zb_debug_gemm(ctx.handle, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(y.data_ptr()), M, N, K, sp)
    for w in ws: run(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): run(ws[i % 8])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    print(f"M={M} N={N} K={K}: {us:8.1f} us  {N*K*2/us/1e3:7.1f} GB/s weights  {2*M*N*K/us/1e6:7.1f} TFLOP/s")
