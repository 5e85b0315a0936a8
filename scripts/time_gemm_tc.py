"""gemm_tc_kernel against torch.matmul (cuBLAS bf16) on the same box: the five (N, K) Linear shapes of a transformer
layer + the heads (SURVEY.md K1 bar: "torch.matmul on the same box"), at M = 128 rows (batch-64 decode), 322 (batch-1
prefill, 2 x 161 tokens) and 20,608 (batch-64 prefill).  Plain store epilogue on both sides; 8 weight copies rotated so
that the weights stream from HBM as they do inside a layer stack.  CUDA events on the launching stream.

  python scripts/time_gemm_tc.py > gpurun_out/gemm_vs_matmul.txt
"""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zonos_b200 import _lib
dev = torch.device("cuda:0"); ctx = _lib.context(dev)
lib = C.CDLL(_lib.LIB_PATH)
lib.zb_debug_gemm.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p]
SHAPES = [("in_proj", 3072, 2048), ("out_proj", 2048, 2048), ("fc1", 16384, 2048), ("fc2", 2048, 8192), ("heads*", 9216, 2048)]   # *9216 of the 9225 head rows (whole 128-row tiles)
ROWS = [128, 322, 20608]
if len(sys.argv) > 1:
    ROWS = [int(v) for v in sys.argv[1].split(",")]


def timed(fn, ws, iters):
    for w in ws: fn(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(ws[i % len(ws)])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print("%-9s %6s | %10s %8s %8s | %10s %8s %8s | %s" % ("matrix", "M", "gemm_tc us", "TFLOP/s", "GB/s W", "matmul us", "TFLOP/s", "GB/s W", "gemm_tc / matmul time"))
for M in ROWS:
    for name, N, K in SHAPES:
        ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(8)]
        x = torch.randn(M, K, device=dev).bfloat16()
        y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        y2 = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        sp = _lib.stream_ptr(dev)

        def run_tc(w):
            ctx.check(lib.zb_debug_gemm(ctx.handle, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(y.data_ptr()), M, N, K, sp))

        def run_mm(w):
            torch.matmul(x, w.t(), out=y2)

        iters = 40 if M <= 322 else 10
        t_tc, t_mm = timed(run_tc, ws, iters), timed(run_mm, ws, iters)
        run_tc(ws[0]); run_mm(ws[0]); torch.cuda.synchronize()
        err = (y.float() - y2.float()).abs().max().item()
        fl = 2.0 * M * N * K
        print("%-9s %6d | %10.1f %8.1f %8.1f | %10.1f %8.1f %8.1f | %.2f   (max |diff| %.3g)" % (
            name, M, t_tc, fl / t_tc / 1e6, N * K * 2 / t_tc / 1e3, t_mm, fl / t_mm / 1e6, N * K * 2 / t_mm / 1e3, t_tc / t_mm, err))
        del ws, x, y, y2
