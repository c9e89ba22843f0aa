"""Diagnostics for the tcgen05 GEMM operand layouts (not a pytest file)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr

lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
rng = np.random.default_rng(0)

def run(M, N, K, ta, tb, A, Bm):
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda()
    dC = torch.full((M, N), -7.0, device="cuda")
    _lib.call("bdetr_gemm", M, N, K, ptr(dA), ta, ptr(dB), tb, None, 0, 0, ptr(dC), stream_ptr())
    torch.cuda.synchronize()
    return dC.cpu().numpy()

for (ta, tb) in [(0, 1), (0, 0), (1, 1), (1, 0)]:
    for (M, N, K) in [(128, 128, 32), (128, 128, 64), (256, 256, 256), (128, 64, 32)]:
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        Bm = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        ref = (A.T if ta else A).astype(np.float64) @ (Bm.T if tb else Bm).astype(np.float64)
        C = run(M, N, K, ta, tb, A, Bm)
        err = np.abs(C - ref).max() / np.abs(ref).max()
        print(f"ta{ta} tb{tb} M{M} N{N} K{K}: err {err:.3e}  |C|mean {np.abs(C).mean():.3f} |ref|mean {np.abs(ref).mean():.3f} untouched {(C == -7).mean():.3f}")
    # one-hot probes at M=N=128, K=32: C = A_logical[:, k*] placed in column n*
    M, N, K = 128, 128, 32
    for (ks, ns) in [(0, 0), (1, 0), (5, 3), (9, 40), (31, 127)]:
        Al = rng.standard_normal((M, K)).astype(np.float32)
        Bl = np.zeros((K, N), np.float32); Bl[ks, ns] = 1.0
        A = np.ascontiguousarray(Al.T) if ta else Al
        Bm = np.ascontiguousarray(Bl.T) if tb else Bl
        C = run(M, N, K, ta, tb, A, Bm)
        nzcols = np.nonzero(np.abs(C).max(0) > 1e-6)[0]
        msg = f"  probe B[{ks},{ns}]=1: nonzero cols {nzcols[:6].tolist()}"
        for c in nzcols[:2]:
            # which A column does it match
            d = np.abs(Al - C[:, c:c + 1]).max(0)
            msg += f" | col {c} ~ A[:, {int(d.argmin())}] (res {d.min():.2e})"
        print(msg)
    for (ms, ks) in [(0, 0), (3, 1), (40, 9), (127, 31)]:
        Al = np.zeros((M, K), np.float32); Al[ms, ks] = 1.0
        Bl = rng.standard_normal((K, N)).astype(np.float32)
        A = np.ascontiguousarray(Al.T) if ta else Al
        Bm = np.ascontiguousarray(Bl.T) if tb else Bl
        C = run(M, N, K, ta, tb, A, Bm)
        nzrows = np.nonzero(np.abs(C).max(1) > 1e-6)[0]
        msg = f"  probe A[{ms},{ks}]=1: nonzero rows {nzrows[:6].tolist()}"
        for r in nzrows[:2]:
            d = np.abs(Bl - C[r:r + 1, :]).max(1)
            msg += f" | row {r} ~ B[{int(d.argmin())}, :] (res {d.min():.2e})"
        print(msg)
