import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from boosted_detr_b200 import _lib
from boosted_detr_b200.layers import Layer
from boosted_detr_b200.transformers import EncoderBlock
from oracle import reference_path as R
lib = _lib.load()
for mode in (_lib.MODE_FP32, _lib.MODE_TF32):
    lib.bdetr_set_mode(mode)
    rng = np.random.default_rng(1)
    B, L, D, H = 2, 256, 256, 8
    Layer._rng = np.random.default_rng(1)
    x = rng.standard_normal((B, L, D)).astype(np.float32)
    pos = R.positional_table(16, 16, D, np.float32).reshape(L, D)
    go = rng.standard_normal((B, L, D)).astype(np.float32)
    blk = EncoderBlock(H, name="blk")
    dx, dpos = torch.from_numpy(x).cuda(), torch.from_numpy(pos).cuda()
    out, ctx = blk.forward([dx, dpos], training=False)
    d_pos = torch.zeros(L, D, device="cuda")
    d_x = blk.backward(ctx, torch.from_numpy(go).cuda(), d_pos)
    torch.cuda.synchronize()
    w = {"p/" + n[len("blk/"):]: o._weights[k].cpu().numpy() for n, o, k in blk.named_weights()}
    p = R.params_to_torch(w, torch.float64)
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    tp = torch.tensor(pos, dtype=torch.float64, requires_grad=True)
    ref = R.encoder_block(tx, tp.expand(B, L, D), p, "p", H, R.Dropout(None), 0, False)
    (ref * torch.tensor(go, dtype=torch.float64)).sum().backward()
    nerr = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    rl2 = lambda a, b: float(np.sqrt(((a-b)**2).sum()/ (b**2).sum()))
    print("mode", mode, "fwd", nerr(out.cpu().numpy(), ref.detach().numpy()), "d_x max", nerr(d_x.cpu().numpy(), tx.grad.numpy()), "d_x relL2", rl2(d_x.cpu().numpy(), tx.grad.numpy()),
          "d_pos", nerr(d_pos.cpu().numpy(), tp.grad.numpy()))
