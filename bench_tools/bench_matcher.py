"""Micro-benchmark of the matcher trio at BASELINE config 4 (not a pytest file)."""
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import synth_preds, synth_targets
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3   # us


def main():
    out = {}
    for (B, T, Q, C, A, w_attr) in [(256, 100, 300, 82, 3, 1.0), (256, 100, 300, 82, 3, 0.0), (256, 100, 300, 48, 296, 1.0),
                                    (16, 20, 100, 82, 3, 1.0)]:
        rng = np.random.default_rng(0)
        tr = synth_targets(rng, B, T, C, A)
        pr = synth_preds(rng, B, Q, C, A)
        d = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (*tr, *pr)]
        cost = torch.empty(B, T, Q, device="cuda")
        c4r = torch.empty(B, T, dtype=torch.int32, device="cuda"); r4c = torch.empty(B, Q, dtype=torch.int32, device="cuda")
        mask = torch.empty(B, T, Q, device="cuda"); asg = torch.empty(B, Q, device="cuda"); st = torch.empty(B, dtype=torch.int32, device="cuda")
        losses = torch.empty(5, B, device="cuda"); iou = torch.empty(Q, device="cuda")
        s = stream_ptr()
        f_cost = lambda: _lib.call("bdetr_cost_matrix_fwd", B, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[4]), ptr(d[5]), ptr(d[6]), 1000.0, 1.0, w_attr, ptr(cost), s)
        prep = torch.empty(_lib.load().bdetr_cost_targets_bytes(B, T, C, A), dtype=torch.uint8, device="cuda")
        f_prep = lambda: _lib.call("bdetr_cost_targets_prepare", B, T, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(prep), s)
        f_cost_p = lambda: _lib.call("bdetr_cost_matrix_prepared", B, T, Q, C, A, ptr(prep), ptr(d[4]), ptr(d[5]), ptr(d[6]), 1000.0, 1.0, w_attr, ptr(cost), s)
        f_lsap = lambda: _lib.call("bdetr_lsap_assign", B, T, Q, ptr(cost), ptr(d[3]), ptr(c4r), ptr(r4c), ptr(mask), ptr(asg), ptr(st), s)
        f_lsap_nomask = lambda: _lib.call("bdetr_lsap_assign", B, T, Q, ptr(cost), ptr(d[3]), ptr(c4r), ptr(r4c), None, None, ptr(st), s)
        f_loss = lambda: _lib.call("bdetr_matched_loss_fwd", B, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), ptr(d[4]), ptr(d[5]), ptr(d[6]), ptr(c4r), ptr(r4c), 1000.0, 1.0, w_attr, 100.0, ptr(losses), ptr(iou), s)
        t_cost = timeit(f_cost); t_prep = timeit(f_prep); t_cost_p = timeit(f_cost_p); t_lsap = timeit(f_lsap); t_lsap2 = timeit(f_lsap_nomask); t_loss = timeit(f_loss)
        bytes_cost = 4 * B * (Q * C + Q * A + 4 * Q + T * C + T * A + 4 * T + T * Q)
        key = f"B{B}_T{T}_Q{Q}_C{C}_A{A}_wattr{w_attr}"
        out[key] = {"cost_us": t_cost, "cost_GBps": bytes_cost / t_cost / 1e3, "targets_prepare_us": t_prep, "cost_prepared_us": t_cost_p, "lsap_with_mask_us": t_lsap,
                    "lsap_index_only_us": t_lsap2, "lsap_us_per_image": t_lsap2 / B, "matched_loss_us": t_loss}
        # CPU: the reference's literal loop (scipy, single thread)
        from scipy.optimize import linear_sum_assignment
        cn = cost.cpu().numpy(); n = tr[3]
        t0 = time.perf_counter()
        for b in range(min(B, 64)):
            linear_sum_assignment(cn[b, :n[b], :])
        out[key]["scipy_us_per_image"] = (time.perf_counter() - t0) / min(B, 64) * 1e6
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_matcher.json", "w"), indent=1)


if __name__ == "__main__":
    main()
