"""Launches each hot kernel a few times at BASELINE sizes (for ncu captures; not a pytest file)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import synth_preds, synth_targets
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
from boosted_detr_b200.transformers import AttentionBlock

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32 if mode == "tf32" else _lib.MODE_FP32)
B, L, D, H = 16, 400, 256, 8
x = torch.randn(B * L, D, device="cuda"); w = torch.randn(D, D, device="cuda"); b = torch.randn(D, device="cuda")
y = torch.empty(B * L, D, device="cuda"); gw = torch.zeros(D, D, device="cuda")
for _ in range(3):
    _lib.call("bdetr_gemm", B * L, D, D, ptr(x), 0, ptr(w), 0, ptr(b), 1, 0, ptr(y), stream_ptr())      # forward
    _lib.call("bdetr_gemm", B * L, D, D, ptr(y), 0, ptr(w), 1, None, 0, 0, ptr(x), stream_ptr())        # dgrad
    _lib.call("bdetr_gemm", D, D, B * L, ptr(x), 1, ptr(y), 0, None, 0, 1, ptr(gw), stream_ptr())       # wgrad
q = torch.randn(B, L, D, device="cuda")
blk = AttentionBlock(H, name="probe")
for _ in range(2):
    out, ctx = blk.forward([q, q, q], training=True, dropout_key=123)
    blk.backward(ctx, torch.randn_like(out))
rng = np.random.default_rng(0)
Bm, T, Q, C, A = 256, 100, 300, 82, 3
tr = synth_targets(rng, Bm, T, C, A); pr = synth_preds(rng, Bm, Q, C, A)
d = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (*tr, *pr)]
cost = torch.empty(Bm, T, Q, device="cuda")
c4r = torch.empty(Bm, T, dtype=torch.int32, device="cuda"); r4c = torch.empty(Bm, Q, dtype=torch.int32, device="cuda")
mask = torch.empty(Bm, T, Q, device="cuda"); asg = torch.empty(Bm, Q, device="cuda"); st = torch.empty(Bm, dtype=torch.int32, device="cuda")
for _ in range(2):
    _lib.call("bdetr_cost_matrix_fwd", Bm, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[4]), ptr(d[5]), ptr(d[6]), 1000.0, 1.0, 1.0, ptr(cost), stream_ptr())
    _lib.call("bdetr_lsap_assign", Bm, T, Q, ptr(cost), ptr(d[3]), ptr(c4r), ptr(r4c), ptr(mask), ptr(asg), ptr(st), stream_ptr())
# optimizer update at BASELINE config 2 size: 8.05 M parameters in ~300 variables
from boosted_detr_b200.optimizers import SGD
sizes = ([65536, 256] * 60 + [102400] * 6 + [256 * 82, 82, 1024, 4] * 6)
sizes = sizes * max(1, 8_047_638 // sum(sizes))
offs = np.cumsum([0] + [(n + 3) // 4 * 4 for n in sizes])
tab = SGD.chunk_table([(str(i), int(offs[i]), n) for i, n in enumerate(sizes)])
tab_d = torch.from_numpy(tab.view(np.uint8).copy()).cuda()
wf = torch.randn(int(offs[-1]), device="cuda"); gf = torch.randn_like(wf) * 1e-3; af = torch.zeros_like(wf); part = torch.zeros(len(tab), device="cuda")
for _ in range(2):
    _lib.call("bdetr_sgd_step", len(tab), ptr(tab_d), ptr(wf), ptr(gf), ptr(af), ptr(part), 1e-3, None, 0.9, 1, 0.1, stream_ptr())
torch.cuda.synchronize()
print("ok")
