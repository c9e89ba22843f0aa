import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import synth_preds, synth_targets
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
rng = np.random.default_rng(0)
Bm, T, Q, C, A = 256, 100, 300, 82, 3
tr = synth_targets(rng, Bm, T, C, A); pr = synth_preds(rng, Bm, Q, C, A)
d = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (*tr, *pr)]
cost = torch.empty(Bm, T, Q, device="cuda")
c4r = torch.empty(Bm, T, dtype=torch.int32, device="cuda"); r4c = torch.empty(Bm, Q, dtype=torch.int32, device="cuda")
mask = torch.empty(Bm, T, Q, device="cuda"); asg = torch.empty(Bm, Q, device="cuda"); st = torch.empty(Bm, dtype=torch.int32, device="cuda")
losses = torch.empty(5, Bm, device="cuda"); iou = torch.empty(Q, device="cuda")
for _ in range(2):
    _lib.call("bdetr_cost_matrix_fwd", Bm, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[4]), ptr(d[5]), ptr(d[6]), 1000.0, 1.0, 1.0, ptr(cost), stream_ptr())
    _lib.call("bdetr_lsap_assign", Bm, T, Q, ptr(cost), ptr(d[3]), ptr(c4r), ptr(r4c), ptr(mask), ptr(asg), ptr(st), stream_ptr())
    _lib.call("bdetr_matched_loss_fwd", Bm, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), ptr(d[4]), ptr(d[5]), ptr(d[6]), ptr(c4r), ptr(r4c), 1000.0, 1.0, 1.0, 100.0, ptr(losses), ptr(iou), stream_ptr())
torch.cuda.synchronize(); print("ok")
