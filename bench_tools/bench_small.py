"""In-graph (no host launch gaps) timings of the small latency-bound kernels at BASELINE config 2 (not a pytest file)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import synth_preds, synth_targets
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
import bench
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
rng = np.random.default_rng(0)
B, T, Q, C, A = 16, 20, 100, 82, 3
tr = synth_targets(rng, B, T, C, A); pr = synth_preds(rng, B, Q, C, A)
d = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (*tr, *pr)]
cost = torch.empty(B, T, Q, device="cuda")
c4r = torch.empty(B, T, dtype=torch.int32, device="cuda"); r4c = torch.empty(B, Q, dtype=torch.int32, device="cuda")
st = torch.empty(B, dtype=torch.int32, device="cuda"); losses = torch.empty(5, B, device="cuda"); iou = torch.empty(Q, device="cuda")
s = stream_ptr
f_cost = lambda: _lib.call("bdetr_cost_matrix_fwd", B, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[4]), ptr(d[5]), ptr(d[6]), 1000.0, 1.0, 1.0, ptr(cost), s())
f_lsap = lambda: _lib.call("bdetr_lsap_assign", B, T, Q, ptr(cost), ptr(d[3]), ptr(c4r), ptr(r4c), None, None, ptr(st), s())
f_loss = lambda: _lib.call("bdetr_matched_loss_fwd", B, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), ptr(d[4]), ptr(d[5]), ptr(d[6]), ptr(c4r), ptr(r4c), 1000.0, 1.0, 1.0, 100.0, ptr(losses), ptr(iou), s())
f_cost(); f_lsap()
for name, fn in (("cost_matrix", f_cost), ("lsap_assign (memset + validate + solver)", f_lsap), ("matched_loss_fwd", f_loss)):
    print(f"{name}: {bench.time_kernel(fn, reps=20) * 1e6:.1f} us per call inside a CUDA graph (B={B}, T={T}, Q={Q})")
