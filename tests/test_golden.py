"""Committed golden fixtures (tests/golden/make_golden.py): the oracle and the C restatement are checked on
CPU, the CUDA path on GPU."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import TINY, tiny_inputs  # noqa: E402
from util import lsap_c_solve  # noqa: E402

G = lambda name: np.load(os.path.join(HERE, "golden", name))


def test_lsap_oracles_match_golden():
    from scipy.optimize import linear_sum_assignment
    g = G("lsap_cases.npz")
    for key in g.files:
        if not key.startswith("cost_"):
            continue
        kind = key[5:]
        c, exp = g[key], g["col4row_" + kind]
        r, cc = linear_sum_assignment(c)
        got = np.full(c.shape[0], -1, np.int32); got[r] = cc
        assert (got == exp).all(), f"installed scipy disagrees with the golden vector {kind}"
        k, a, b = lsap_c_solve(c)
        got = np.full(c.shape[0], -1, np.int32); got[a] = b
        assert (got == exp).all(), f"oracle/lsap_ref.c disagrees with the golden vector {kind}"


def test_oracle_matching_loss_matches_golden():
    from oracle import reference_path as R
    g = G("matching_loss_small.npz")
    dt = torch.float64
    y_true = [torch.tensor(g["cat_true"], dtype=dt), torch.tensor(g["attr_true"], dtype=dt), torch.tensor(g["box_true"], dtype=dt), g["num_objects"]]
    y_pred = [torch.tensor(g[k], dtype=dt) for k in ("cat_pred", "attr_pred", "box_pred")]
    losses, iou, mask, cost = R.matching_loss(y_true, y_pred, R.model_weights(1.0))
    assert np.allclose(cost.numpy(), g["cost"], rtol=1e-12) and (mask.numpy() == g["mask"]).all()
    assert np.allclose(np.stack([l.numpy() for l in losses]), g["losses"], rtol=1e-12)
    # hand-checkable identities: padded rows are never matched, one match per real target
    n = g["num_objects"]
    assert (g["mask"].sum(axis=(1, 2)) == n).all()


def test_oracle_tiny_model_matches_golden():
    from oracle import reference_path as R
    g = G("tiny_model.npz")
    w, feats, tg = tiny_inputs()
    t = TINY
    out, grads, _ = R.train_step_reference(w, feats, tg, t["N"], t["H"], torch.float64, dropout_seed=5, weights=R.model_weights(1.0))
    assert np.allclose(out["loss"].detach().numpy(), g["loss"], rtol=1e-10)
    assert np.allclose(grads["DecoderPrep/init_decoder_features"], g["grad_query"], rtol=1e-8, atol=1e-12)
    # fp32 evaluation of the oracle agrees with its fp64 twin (bounds the oracle's own rounding)
    out32, _, _ = R.train_step_reference(w, feats, tg, t["N"], t["H"], torch.float32, dropout_seed=5, weights=R.model_weights(1.0),
                                         forced_masks=[m.to(torch.float32) for m in out["masks"]])
    assert np.abs(out32["loss"].detach().numpy() - g["loss"]).max() / np.abs(g["loss"]).max() < 1e-5


@pytest.mark.gpu
def test_gpu_matches_golden():
    from boosted_detr_b200.boosted_model import BoostedDETR
    from boosted_detr_b200.losses_and_metrics import MatchingAssignment, MatchingLoss
    g = G("lsap_cases.npz")
    for key in g.files:
        if key.startswith("cost_"):
            c = g[key]
            n = torch.tensor([c.shape[0]], dtype=torch.int32).cuda()
            c4r = MatchingAssignment().assign(torch.from_numpy(c[None]).cuda(), n)[0].cpu().numpy()[0]
            assert (c4r == g["col4row_" + key[5:]]).all(), key
    g = G("matching_loss_small.npz")
    dev = lambda k: torch.from_numpy(g[k]).cuda()
    ml = MatchingLoss(attribute_weight=1.0)
    losses, iou = ml([[dev("cat_true"), dev("attr_true"), dev("box_true"), dev("num_objects")],
                      [dev("cat_pred"), dev("attr_pred"), dev("box_pred")]])
    got = torch.stack(losses).cpu().numpy()
    assert np.abs(got - g["losses"]).max() / np.abs(g["losses"]).max() < 1e-5
    assert np.abs(ml.last_ctx["cost"].cpu().numpy() - g["cost"]).max() / np.abs(g["cost"]).max() < 1e-5
    # tiny boosted model (D=128, 4 heads): inference, training loss, gradients
    g = G("tiny_model.npz")
    w, feats, tg = tiny_inputs()
    t = TINY
    vocab = {"category": ["c"] * (t["C"] - 2), "attribute": ["a"] * (t["A"] - 2)}
    model = BoostedDETR(num_object_preds=t["Q"], image_size=(t["rows"] * 32, t["cols"] * 32), num_encoder_blocks=t["N"],
                        num_encoder_heads=t["H"], encoder_dim=t["D"], num_decoder_blocks=t["N"], num_decoder_heads=t["H"],
                        decoder_dim=t["D"], vocab_dict=vocab, attribute_weight=1.0).build()
    model.set_weights_dict(w)
    cat, attr, box = model.call({"features": feats}, training=False)
    for got_t, key in ((cat, "inf_cat"), (attr, "inf_attr"), (box, "inf_box")):
        assert np.abs(got_t.cpu().numpy() - g[key]).max() / np.abs(g[key]).max() < 1e-5, key
    model.dropout_seed = 5
    model.train_step({"features": feats, "category": tg[0], "attribute": tg[1], "bbox": tg[2], "num_objects": tg[3]})
    assert np.abs(model.metric_tensors["loss"].cpu().numpy() - g["loss"]).max() / np.abs(g["loss"]).max() < 1e-5
    gq = model.get_grads_dict()["DecoderPrep/init_decoder_features"]
    assert np.abs(gq - g["grad_query"]).max() / np.abs(g["grad_query"]).max() < 1e-4
    gp = model.get_grads_dict()["ImageEncoderAttention_0/positional_encoding"]
    assert np.abs(gp - g["grad_pos0"]).max() / np.abs(g["grad_pos0"]).max() < 1e-4
