"""Three launches of the long-sequence attention core at BASELINE config 5's length, one image (for `ncu --set full`; not a
pytest file).  usage: python tools/prof_attention.py [tf32|f16]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
which = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
B, Lq, Lk, H, d = 1, 20020, 20020, 8, 32      # BASELINE config 5 sequence length, one image
D = H * d
q = torch.randn(B, Lq, D, device="cuda"); k = torch.randn(B, Lk, D, device="cuda"); v = torch.randn(B, Lk, D, device="cuda")
o = torch.empty(B, H, Lq, d, device="cuda"); lse = torch.empty(B, H, Lq, device="cuda")
ws = torch.empty(lib.bdetr_attention_f16_workspace_bytes(B, H, Lq, Lk, d) // 2, dtype=torch.float16, device="cuda")
for _ in range(3):
    if which == "f16":
        _lib.call("bdetr_attention_core_fwd_f16", B, H, Lq, Lk, d, ptr(q), ptr(k), ptr(v), ptr(ws), ptr(o), ptr(lse), stream_ptr())
    else:
        _lib.call("bdetr_attention_core_fwd", B, H, Lq, Lk, d, ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), stream_ptr())
torch.cuda.synchronize(); print("ok")
