import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
for (M, N, K) in [(6400, 256, 32), (6400, 256, 256), (128, 64, 256), (6400, 256, 1024)]:
    A = torch.randn(M, K, device="cuda"); Bm = torch.randn(K, N, device="cuda"); C = torch.zeros(M, N, device="cuda")
    fn = lambda: _lib.call("bdetr_gemm", M, N, K, ptr(A), 0, ptr(Bm), 0, None, 0, 0, ptr(C), stream_ptr())
    for _ in range(3): fn()
    lib.bdetr_debug_set_timeline(ptr(buf))
    for rep in range(3):
        fn(); torch.cuda.synchronize()
        t = buf.cpu().numpy(); d = (t - t[0])
        print(f"M{M} N{N} K{K}: setup {d[1]} | 2nd TMA issue {d[2]} | first stage landed {d[3]} | last MMA commit {d[4]} | accum ready {d[5]} | epilogue done {d[6]} | teardown {d[7]}  (cycles from CTA entry)")
    lib.bdetr_debug_set_timeline(None)
