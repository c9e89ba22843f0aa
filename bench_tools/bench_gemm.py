"""GEMM latency decomposition (not a pytest file): time vs K, epilogue mode, warm/cold L2."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
lib = _lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def t(fn, cold, iters=5, reps=100):
    """average over `reps` back-to-back launches captured in a CUDA graph (no host launch gaps)"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    ts = []
    for _ in range(iters):
        if cold: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / reps)
    return float(np.median(ts))
# launch-overhead yardstick: an empty-ish kernel
y0 = torch.zeros(256, device="cuda"); x0 = torch.zeros(256, device="cuda")
print("accumulate(256 floats) warm: %.1f us" % t(lambda: _lib.call("bdetr_accumulate", 256, ptr(x0), ptr(y0), stream_ptr()), False))
for mode_name, mode in (("tf32", _lib.MODE_TF32), ("fp32", _lib.MODE_FP32)):
    lib.bdetr_set_mode(mode)
    for (M, N, K, ta, tb, bias, beta) in [(6400, 256, 32, 0, 0, 0, 0), (6400, 256, 64, 0, 0, 0, 0), (6400, 256, 128, 0, 0, 0, 0), (6400, 256, 256, 0, 0, 0, 0),
                                          (6400, 256, 512, 0, 0, 0, 0), (6400, 256, 1024, 0, 0, 0, 0), (6400, 256, 256, 0, 0, 1, 0), (6400, 256, 256, 0, 0, 0, 1),
                                          (6400, 256, 256, 0, 1, 0, 1), (1600, 256, 256, 0, 0, 1, 0), (128, 64, 256, 0, 0, 0, 0), (256, 256, 6400, 1, 0, 0, 1), (12800, 512, 256, 0, 0, 1, 0)]:
        A = torch.randn((K, M) if ta else (M, K), device="cuda"); Bm = torch.randn((N, K) if tb else (K, N), device="cuda")
        bv = torch.randn(N, device="cuda"); C = torch.zeros(M, N, device="cuda")
        fn = lambda: _lib.call("bdetr_gemm", M, N, K, ptr(A), ta, ptr(Bm), tb, ptr(bv) if bias else None, 0, beta, ptr(C), stream_ptr())
        w, c = t(fn, False), t(fn, True)
        print(f"{mode_name} M{M} N{N} K{K} ta{ta} tb{tb} bias{bias} beta{beta}: warm {w:6.1f} us  cold {c:6.1f} us  ({2.0*M*N*K/w/1e6:7.1f} TF/s warm)", flush=True)
