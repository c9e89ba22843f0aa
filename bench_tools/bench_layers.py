"""Layer-level micro-benchmark (not a pytest file): fused vs layer-level entry points of the FFN / attention blocks,
warm (CUDA graph of back-to-back launches), tensor-core mode."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from boosted_detr_b200 import _lib
from boosted_detr_b200.layers import Layer
from boosted_detr_b200.transformers import AttentionBlock, EncoderBlock, FeedForwardBlock

lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
Layer._rng = np.random.default_rng(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for M in (6400, 1600):
    x = torch.randn(M // 100, 100, 256, device="cuda") if M == 1600 else torch.randn(16, 400, 256, device="cuda")
    ffn = FeedForwardBlock(name="f")
    ffn.forward([x], training=True)
    for fused in ("1", "0"):
        os.environ["BDETR_FUSED"] = fused
        import boosted_detr_b200.transformers as T
        orig = T.fused_path
        T.fused_path = (lambda D=256: fused == "1")
        t = bench.time_kernel(lambda: ffn.forward([x], training=True, dropout_key=5), reps=20, iters=5, flush=flush)
        out, ctx = ffn.forward([x], training=True, dropout_key=5)
        go = torch.randn_like(out)
        tb = bench.time_kernel(lambda: ffn.backward(ctx, go), reps=20, iters=5, flush=flush)
        print(f"FFN M={M} fused={fused}: fwd {t*1e6:.1f} us, bwd {tb*1e6:.1f} us", flush=True)
        T.fused_path = orig
pos = torch.randn(400, 256, device="cuda")
x = torch.randn(16, 400, 256, device="cuda")
enc = EncoderBlock(8, name="e")
enc.forward([x, pos], training=True)
import boosted_detr_b200.transformers as T
orig = T.fused_path
for fused in ("1", "0"):
    T.fused_path = (lambda D=256: fused == "1")
    t = bench.time_kernel(lambda: enc.forward([x, pos], training=True, dropout_keys=(3, 4)), reps=10, iters=5, flush=flush)
    y, ctx = enc.forward([x, pos], training=True, dropout_keys=(3, 4))
    go = torch.randn_like(y); dp = torch.zeros(400, 256, device="cuda")
    tb = bench.time_kernel(lambda: enc.backward(ctx, go, dp), reps=10, iters=5, flush=flush)
    print(f"EncoderBlock B=16 L=400 fused={fused}: fwd {t*1e6:.1f} us, bwd {tb*1e6:.1f} us", flush=True)
T.fused_path = orig
