"""Pins the oracle's TensorFlow-side restatements against INDEPENDENT implementations of the same published
algorithms that exist in this image (TensorFlow / tensorflow_addons themselves cannot be installed):
  tfa giou_loss            <-> torchvision.ops.generalized_box_iou / box_iou
  tfa SigmoidFocalCrossEntropy <-> torchvision.ops.sigmoid_focal_loss
  Keras LayerNormalization / BatchNormalization / softmax attention <-> torch.nn.functional
  Keras SGD(momentum, nesterov, clipnorm) <-> torch.optim.SGD + per-parameter clip_grad_norm_
  CosineDecayRestarts      <-> torch.optim.lr_scheduler.CosineAnnealingWarmRestarts (m_mul = 1 case)
These are CPU tests of the checker, not of the product."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_path as R

tv = pytest.importorskip("torchvision")
from torchvision import ops as tvops  # noqa: E402


def _boxes(rng, n):
    x, y = rng.uniform(0, 0.8, n), rng.uniform(0, 0.8, n)
    w, h = rng.uniform(0.02, 0.3, n), rng.uniform(0.02, 0.3, n)
    return np.stack([x, y, w, h], -1)


def test_giou_and_iou_match_torchvision():
    rng = np.random.default_rng(0)
    t, q = _boxes(rng, 17), _boxes(rng, 23)
    tt, qq = R.coco_to_tf(torch.tensor(t)), R.coco_to_tf(torch.tensor(q))                 # [ymin,xmin,ymax,xmax]
    got_giou = R.tfa_giou(tt[:, None, :], qq[None, :, :], "giou")
    got_iou = R.tfa_giou(tt[:, None, :], qq[None, :, :], "iou")
    xyxy = lambda b: torch.stack([b[:, 1], b[:, 0], b[:, 3], b[:, 2]], -1)               # torchvision wants x1,y1,x2,y2
    assert torch.allclose(got_giou, tvops.generalized_box_iou(xyxy(tt), xyxy(qq)), atol=1e-12)
    assert torch.allclose(got_iou, tvops.box_iou(xyxy(tt), xyxy(qq)), atol=1e-12)
    # hand-worked pairs: unit squares offset by half a side along x -> IoU 1/3 and the enclosing box IS the union
    # (GIoU = IoU); offset diagonally -> intersection 1/4, union 7/4, enclosing 9/4 -> GIoU = 1/7 - (9/4 - 7/4)/(9/4)
    a = torch.tensor([[0.0, 0.0, 1.0, 1.0]], dtype=torch.float64)
    b = torch.tensor([[0.0, 0.5, 1.0, 1.5]], dtype=torch.float64); c = torch.tensor([[0.5, 0.5, 1.5, 1.5]], dtype=torch.float64)
    assert float(R.tfa_giou(a, b, "iou")) == pytest.approx(1.0 / 3.0)
    assert float(R.tfa_giou(a, b, "giou")) == pytest.approx(1.0 / 3.0)
    assert float(R.tfa_giou(a, c, "iou")) == pytest.approx(1.0 / 7.0)
    assert float(R.tfa_giou(a, c, "giou")) == pytest.approx(1.0 / 7.0 - 2.0 / 9.0)
    # the box term of the cost for that pair: 2 (1 - GIoU) + 5 mean((10 t - 10 p)^2) with all four coordinates 0.5 apart
    cost = R.box_loss(torch.tensor([[0.0, 0.0, 1.0, 1.0]], dtype=torch.float64), torch.tensor([[0.5, 0.5, 1.0, 1.0]], dtype=torch.float64))
    assert float(cost) == pytest.approx(2.0 * (1.0 - (1.0 / 7.0 - 2.0 / 9.0)) + 5.0 * 25.0)
    # degenerate boxes (the padded target rows: w = h = -10): clamped to zero area, divide_no_nan -> finite
    pad = R.coco_to_tf(torch.tensor([[-10.0, -10.0, -10.0, -10.0]]))
    assert torch.isfinite(R.tfa_giou(pad[:, None, :], qq[None, :, :], "giou")).all()


def test_focal_attribute_loss_matches_torchvision():
    rng = np.random.default_rng(1)
    p = torch.tensor(rng.uniform(0.01, 0.99, (5, 7, 11)))                                  # inside safe_clip's range
    y = torch.tensor((rng.uniform(size=(5, 7, 11)) < 0.3).astype(np.float64))
    got = R.attribute_loss(y, p)                                                           # mean over attributes
    logits = torch.log(p) - torch.log1p(-p)
    ref = tvops.sigmoid_focal_loss(logits, y, alpha=0.25, gamma=2.0, reduction="none").mean(dim=-1)
    assert torch.allclose(got, ref, rtol=2e-5, atol=1e-9)                                  # Keras' 1e-7 epsilons are the only difference


def test_category_loss_closed_form():
    """mean_c BCE(y_c, clip(p_c) * y_c) == -log(clip(p[c*]) + 1e-7) / C for one-hot y."""
    rng = np.random.default_rng(2)
    C = 9
    p = torch.tensor(rng.uniform(0, 1.2, (4, 6, C)))
    cls = rng.integers(0, C, (4, 6))
    y = torch.nn.functional.one_hot(torch.tensor(cls), C).double()
    got = R.category_loss(y, p)
    pc = torch.clamp(torch.gather(p, -1, torch.tensor(cls)[..., None])[..., 0], 0.001, 0.999)
    # every y = 0 term is -log(1 - 1e-7 + 1e-7) = -log(1) = 0 in exact arithmetic (1e-16 in float64): only c* remains
    assert torch.allclose(got, -torch.log(pc + 1e-7) / C, rtol=1e-9, atol=1e-14)


def test_layer_norm_batch_norm_attention_match_torch_functional():
    rng = np.random.default_rng(3)
    x = torch.tensor(rng.standard_normal((3, 5, 16)))
    p = {"ln/gamma": torch.tensor(rng.standard_normal(16)), "ln/beta": torch.tensor(rng.standard_normal(16)),
         "bn/gamma": torch.tensor(rng.standard_normal(16)), "bn/beta": torch.tensor(rng.standard_normal(16)),
         "bn/moving_mean": torch.zeros(16, dtype=torch.float64), "bn/moving_variance": torch.ones(16, dtype=torch.float64)}
    F = torch.nn.functional
    assert torch.allclose(R.layer_norm(x, p, "ln"), F.layer_norm(x, (16,), p["ln/gamma"], p["ln/beta"], eps=1e-3), atol=1e-12)
    stats = {}
    got = R.batch_norm(x, p, "bn", True, stats)
    rm, rv = torch.zeros(16, dtype=torch.float64), torch.ones(16, dtype=torch.float64)
    ref = F.batch_norm(x.reshape(-1, 16), rm, rv, p["bn/gamma"], p["bn/beta"], training=True, momentum=0.01, eps=1e-3)
    assert torch.allclose(got.reshape(-1, 16), ref, atol=1e-12)
    assert torch.allclose(stats["bn/moving_mean"], rm, atol=1e-12)          # Keras momentum .99 == torch momentum .01
    # (torch updates the running variance with the UNBIASED estimate, Keras with the biased one: compare by formula)
    flat = x.reshape(-1, 16)
    assert torch.allclose(stats["bn/moving_variance"], 0.99 * torch.ones(16) + 0.01 * flat.var(dim=0, unbiased=False), atol=1e-12)
    # attention core: softmax(q k^T / sqrt(d)) v per head == scaled_dot_product_attention, then the reference's raw
    # [B,H,L,d] -> [B,L,H*d] reshape WITHOUT the permute back (quirk Q1)
    B, L, H, d = 2, 6, 4, 8
    D = H * d
    eye = torch.eye(D, dtype=torch.float64)
    pa = {f"a/{n}/kernel": eye for n in ("QueryProjection", "KeyProjection", "ValueProjection", "OutputProjection")}
    pa.update({f"a/{n}/bias": torch.zeros(D, dtype=torch.float64) for n in ("QueryProjection", "KeyProjection", "ValueProjection", "OutputProjection")})
    q, k, v = (torch.tensor(rng.standard_normal((B, L, D))) for _ in range(3))
    got = R.multihead_attention(q, k, v, pa, "a", H)
    split = lambda t: t.reshape(B, L, H, d).permute(0, 2, 1, 3)
    ref = F.scaled_dot_product_attention(split(q), split(k), split(v)).contiguous().reshape(B, L, D)
    assert torch.allclose(got, ref, atol=1e-12)


@pytest.mark.parametrize("nesterov", [True, False])
def test_sgd_reference_matches_torch_optim(nesterov):
    rng = np.random.default_rng(4)
    shapes = {"a": (7, 5), "b": (13,), "c": (3, 3, 2)}
    w = {k: rng.standard_normal(s) for k, s in shapes.items()}
    acc = {k: np.zeros(s) for k, s in shapes.items()}
    tw = {k: torch.tensor(v.copy(), requires_grad=True) for k, v in w.items()}
    lr, mom, clip = 0.05, 0.9, 0.1
    opt = torch.optim.SGD(list(tw.values()), lr=lr, momentum=mom, nesterov=nesterov)
    for step in range(4):
        g = {k: rng.standard_normal(s) * (0.01 if step % 2 else 1.0) for k, s in shapes.items()}       # norms on both sides of the clip
        w, acc = R.sgd_step_reference(w, g, acc, lr, mom, nesterov, clip)
        for k in shapes:
            tw[k].grad = torch.tensor(g[k].copy())
            torch.nn.utils.clip_grad_norm_([tw[k]], clip)                                               # PER VARIABLE, like Keras clipnorm
        opt.step()
        for k in shapes:
            # torch divides by (norm + 1e-6) and clamps the factor at 1; tf.clip_by_norm divides by max(norm, c): 1e-5 apart
            assert np.allclose(w[k], tw[k].detach().numpy(), rtol=2e-5, atol=1e-8), (k, step)


def test_cosine_decay_restarts_matches_torch_scheduler():
    lr0, first, alpha = 1e-3, 40, 0.1
    prm = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([prm], lr=lr0)
    sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=first, T_mult=2, eta_min=alpha * lr0)
    for step in range(0, 300):
        ref = opt.param_groups[0]["lr"]
        got = R.cosine_decay_restarts(step, lr0, first, t_mul=2.0, m_mul=1.0, alpha=alpha)
        assert got == pytest.approx(ref, rel=1e-9, abs=1e-15), step
        opt.step(); sch.step()
    # m_mul scales the peak of every restart: period i starts at lr0 * ((1 - alpha) * m_mul^i + alpha)
    for i, start in enumerate([0, 40, 120, 280]):
        assert R.cosine_decay_restarts(start, lr0, first, 2.0, 0.95, alpha) == pytest.approx(lr0 * ((1 - alpha) * 0.95 ** i + alpha), rel=1e-9)
