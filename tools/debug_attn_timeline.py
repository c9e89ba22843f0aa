"""Cycle accounting of the multi-stream attention kernel's CTA (0,0,0) (not a pytest file)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32); lib.bdetr_debug_force_attention_kernel(2)
B, L, H, d = 1, 20020, 8, 32
D = H * d
for Lq in (20020, 128, 256):          # 3 live streams per CTA / 1 / 2
    q = torch.randn(B, Lq, D, device="cuda"); k = torch.randn(B, L, D, device="cuda"); v = torch.randn(B, L, D, device="cuda")
    o = torch.empty(B, H, Lq, d, device="cuda"); lse = torch.empty(B, H, Lq, device="cuda")
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    fn = lambda: _lib.call("bdetr_attention_core_fwd", B, H, Lq, L, d, ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), stream_ptr())
    for _ in range(2): fn()
    lib.bdetr_debug_set_timeline(ptr(buf))
    for rep in range(2):
        fn(); torch.cuda.synchronize()
        t = buf.cpu().numpy(); n = L // 128 + 1
        print(f"Lq {Lq}: per tile (cycles), softmax warp 2: wait S {t[0]/n:.0f} | pass {t[1]/n:.0f} | st+arrive {t[2]/n:.0f} | final wait O {t[3]:.0f}")
    lib.bdetr_debug_set_timeline(None)
