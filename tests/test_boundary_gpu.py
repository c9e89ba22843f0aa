"""Call conventions of the reference's layer protocol that round 1 left as stubs: stand-alone
MultiheadAttention.call, MatchingMetric.call, the loss functions as callables, and the workspace-size queries."""
import numpy as np
import pytest
import torch

from test_dense_gpu import nerr
from util import synth_preds, synth_targets

pytestmark = pytest.mark.gpu


def test_multihead_attention_call_vs_oracle():
    from boosted_detr_b200.layers import Layer
    from boosted_detr_b200.transformers import MultiheadAttention
    from oracle import reference_path as R
    rng = np.random.default_rng(2)
    B, Lq, Lk, D, H = 2, 37, 53, 256, 8
    Layer._rng = np.random.default_rng(2)
    q, k, v = (rng.standard_normal((B, L, D)).astype(np.float32) for L in (Lq, Lk, Lk))
    mha = MultiheadAttention(H, D // H, name="mha")
    out = mha([torch.from_numpy(x).cuda() for x in (q, k, v)])
    w = {"p/" + n[len("mha/"):]: o._weights[kk].cpu().numpy() for n, o, kk in mha.named_weights()}
    ref = R.multihead_attention(*(torch.tensor(x, dtype=torch.float64) for x in (q, k, v)), R.params_to_torch(w), "p", H)
    assert nerr(out.cpu().numpy(), ref.numpy()) < 1e-5
    with pytest.raises(NotImplementedError):
        mha([torch.from_numpy(x).cuda() for x in (q, k, v)], attention_mask=torch.ones(1))


def test_matching_metric_and_loss_callables_vs_oracle():
    from boosted_detr_b200.losses_and_metrics import (AttributeLoss, BoxLoss, CategoryLoss, MatchingMask, MatchingMetric)
    from oracle import reference_path as R
    rng = np.random.default_rng(4)
    B, T, Q, C, A = 3, 7, 19, 82, 3
    cat, attr, box, n = synth_targets(rng, B, T, C, A, attr_p=0.3)
    pc, pa, pb = synth_preds(rng, B, Q, C, A, k_sum=2)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    t64 = lambda x: torch.tensor(x, dtype=torch.float64)
    for func, ref_fn, yt, yp in ((CategoryLoss, R.category_loss, cat, pc), (AttributeLoss, R.attribute_loss, attr, pa), (BoxLoss, R.box_loss, box, pb)):
        ref = R.cost_array(t64(yt), t64(yp), ref_fn).numpy()
        got = func(dev(yt), dev(yp)).cpu().numpy()                                    # plain [B,T,K] / [B,Q,K]
        got_b = func(dev(yt).unsqueeze(2), dev(yp).unsqueeze(1)).cpu().numpy()        # the reference's broadcast operands
        assert nerr(got, ref) < 2e-6 and (got == got_b).all(), func.__name__
    iou_ref = R.cost_array(t64(box), t64(pb), R.iou_metric).numpy()
    mm = MatchingMetric()
    assert nerr(mm([dev(box), dev(pb)]).cpu().numpy(), iou_ref) < 2e-6
    cost = dev(rng.random((B, T, Q)).astype(np.float32))
    mask, _ = MatchingMask()([cost, dev(n)])
    got = mm([dev(box), dev(pb)], assignment_mask=mask).cpu().numpy()
    assert nerr(got, mask.cpu().numpy() * iou_ref) < 1e-5        # (normalised by the few matched IoUs, not by the full matrix)


def test_workspace_queries_cover_what_the_layers_allocate():
    from boosted_detr_b200 import _lib
    lib = _lib.load()
    B, Lq, Lk, D, H, M, Dh, C, A = 4, 100, 400, 256, 8, 400, 256, 82, 3
    q, k = B * Lq * D, B * Lk * D
    assert lib.bdetr_attention_block_saved_bytes(B, Lq, Lk, D, H, 1) == 4 * (3 * q + 2 * k + B * H * Lq + 2 * B * Lq)
    assert lib.bdetr_attention_block_saved_bytes(B, Lq, Lk, D, H, 0) == 4 * (2 * q + 2 * k + B * H * Lq + 2 * B * Lq)
    assert lib.bdetr_ffn_block_saved_bytes(M, D, 1) == 4 * (2 * M * D + 2 * M)
    assert lib.bdetr_ffn_block_scratch_bytes(M, D) == 8 * M * D
    assert lib.bdetr_heads_saved_bytes(M, Dh, C, A) > 4 * 3 * M * Dh
    assert lib.bdetr_heads_scratch_bytes(M, Dh, C, A) > 4 * 3 * M * Dh
    assert lib.bdetr_attention_block_scratch_bytes(B, Lq, Lk, D, H) >= 4 * (3 * q + 2 * k)
