"""clock64 stamps of CTA 0 of the fused Dense + LayerNorm kernel (not a pytest file)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr
from boosted_detr_b200.layers import Layer
from boosted_detr_b200.transformers import FeedForwardBlock
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
Layer._rng = np.random.default_rng(0)
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
for M in (6400, 1600):
    x = torch.randn(M // 100, 100, 256, device="cuda")
    ffn = FeedForwardBlock(name="f")
    for _ in range(3): ffn.forward([x], training=True, dropout_key=5)
    lib.bdetr_debug_set_timeline(ptr(buf))
    for rep in range(2):
        ffn.forward([x], training=True, dropout_key=5); torch.cuda.synchronize()
        t = buf.cpu().numpy(); d = t - t[0]
        if os.environ.get("BDETR_LN_SPLIT", "4") == "1":
            print(f"M{M}: pdl wait done {d[1]} | params staged {d[2]} | accumulator ready {d[3]} | phase 1 done {d[4]} | sync {d[5]} | phase 2 done {d[6]} | end {d[7]}")
        else:
            print(f"M{M} split: pdl wait done {d[1]} | tile staged, loops done, mask hashed {d[2]} | accumulator ready {d[3]} | partials pushed {d[7]} | z store issued, cluster barrier passed {d[4]} | statistics merged {d[8]} | out boxes read by TMA {d[5]} | end {d[6]}")
    lib.bdetr_debug_set_timeline(None)
