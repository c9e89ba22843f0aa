"""Generates the committed golden fixtures from the ORACLE (oracle/reference_path.py, fp64) and the installed
scipy matcher.  The reference has no golden vectors of its own and TensorFlow cannot run here,
so these fixtures pin (a) the matcher against scipy 1.18.1 and (b) the oracle against later accidental edits;
they are NOT outputs of the TensorFlow reference ("parity unpinned" at the TF boundary).

    python tests/golden/make_golden.py      # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from util import make_weights, synth_preds, synth_targets  # noqa: E402
from oracle import reference_path as R  # noqa: E402

TINY = dict(N=2, rows=3, cols=4, D=128, H=4, Q=10, T=5, C=6, A=3, B=2)


def lsap_cases():
    rng = np.random.default_rng(42)
    out = {}
    kinds = ["uniform", "ties", "identical_cols", "inf", "tall", "const"]
    for i, kind in enumerate(kinds):
        T, Q = (9, 14) if kind != "tall" else (12, 5)
        if kind == "uniform":
            c = rng.random((T, Q))
        elif kind == "ties":
            c = np.round(rng.random((T, Q)) * 4) / 4
        elif kind == "identical_cols":
            c = np.tile(rng.random((T, 1)), (1, Q))
        elif kind == "inf":
            c = rng.integers(0, 4, (T, Q)).astype(float)
            c[rng.random((T, Q)) < 0.2] = np.inf
        elif kind == "tall":
            c = np.round(rng.random((T, Q)) * 8) / 8
        else:
            c = np.full((T, Q), 1.5)
        c = c.astype(np.float32)
        r, cc = linear_sum_assignment(c)
        c4r = np.full(T, -1, np.int32)
        c4r[r] = cc
        out[f"cost_{kind}"] = c
        out[f"col4row_{kind}"] = c4r
    return out


def tiny_inputs():
    t = TINY
    rng = np.random.default_rng(7)
    w = make_weights(rng, t["N"], t["rows"], t["cols"], t["D"], t["Q"], t["C"], t["A"])
    cat, attr, box, n = synth_targets(rng, t["B"], t["T"], t["C"], t["A"], attr_p=0.3)
    feats = np.tanh(rng.standard_normal((t["B"], t["rows"], t["cols"], t["D"]))).astype(np.float32)
    return w, feats, (cat, attr, box, n)


def main():
    np.savez_compressed(os.path.join(HERE, "lsap_cases.npz"), **lsap_cases())
    # cost matrix + matched loss on a small batch
    rng = np.random.default_rng(11)
    B, T, Q, C, A = 3, 6, 9, 7, 4
    tr = synth_targets(rng, B, T, C, A, attr_p=0.3)
    pr = synth_preds(rng, B, Q, C, A, k_sum=2)
    wts = R.model_weights(1.0)
    dt = torch.float64
    y_true = [torch.tensor(tr[0], dtype=dt), torch.tensor(tr[1], dtype=dt), torch.tensor(tr[2], dtype=dt), tr[3]]
    y_pred = [torch.tensor(p, dtype=dt) for p in pr]
    losses, iou, mask, cost = R.matching_loss(y_true, y_pred, wts)
    np.savez_compressed(os.path.join(HERE, "matching_loss_small.npz"), cat_true=tr[0], attr_true=tr[1], box_true=tr[2],
                        num_objects=tr[3], cat_pred=pr[0], attr_pred=pr[1], box_pred=pr[2], cost=cost.numpy(),
                        mask=mask.numpy(), losses=np.stack([l.numpy() for l in losses]), iou=iou.numpy())
    # tiny boosted model: forward (inference + training) and gradient checksums
    w, feats, tg = tiny_inputs()
    t = TINY
    inf = R.boosted_detr_call(R.params_to_torch(w), torch.tensor(feats, dtype=dt), None, t["N"], t["H"], training=False)
    out, grads, stats = R.train_step_reference(w, feats, tg, t["N"], t["H"], dt, dropout_seed=5, weights=wts)
    gsum = {k: np.array([np.abs(v).sum(), v.sum(), np.abs(v).max()]) for k, v in grads.items()}
    keys = sorted(gsum)
    np.savez_compressed(os.path.join(HERE, "tiny_model.npz"),
                        inf_cat=inf["preds"][0].numpy(), inf_attr=inf["preds"][1].numpy(), inf_box=inf["preds"][2].numpy(),
                        train_cat=out["preds"][0].detach().numpy(), train_box=out["preds"][2].detach().numpy(),
                        loss=out["loss"].detach().numpy(), iou=out["metrics"]["IOU"].detach().numpy(),
                        grad_keys=np.array(keys), grad_stats=np.stack([gsum[k] for k in keys]),
                        grad_query=grads["DecoderPrep/init_decoder_features"],
                        grad_pos0=grads["ImageEncoderAttention_0/positional_encoding"])
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
