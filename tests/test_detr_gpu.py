"""Plain DETR (reference model.py) on the shared layers, fp32 mode, against the fp64 oracle."""
import numpy as np
import pytest
import torch

from test_dense_gpu import _perturb, nerr
from util import synth_targets

pytestmark = pytest.mark.gpu


def test_plain_detr_train_step_vs_oracle():
    from boosted_detr_b200.model import DETR
    from boosted_detr_b200.parameters import ModelParameters
    from oracle import reference_path as R
    p = ModelParameters("COCO").default_params()
    p.pop("pad_value"); p.pop("oov_value")
    NE, ND, B, T, Q = 2, 3, 2, 6, 24
    p.update(num_object_preds=Q, num_encoder_blocks=NE, num_decoder_blocks=ND, image_size=(6 * 32, 7 * 32))
    model = DETR(**p, attribute_weight=1.0, seed=1).build()
    rng = np.random.default_rng(2)
    _perturb(model, rng)
    w = model.get_weights_dict()
    assert w["CategoryPredictionHead/DenseCateg/kernel"].shape == (256, 1024)          # hidden width 4 x decoder_dim (model.py:106)
    cat, attr, box, n = synth_targets(rng, B, T, model.num_categories, model.num_attributes, attr_p=0.05)
    feats = np.tanh(rng.standard_normal((B, 6, 7, 256))).astype(np.float32)
    model.dropout_seed = 13
    model.train_step({"features": feats, "category": cat, "attribute": attr, "bbox": box, "num_objects": n})
    torch.cuda.synchronize()
    pt = R.params_to_torch(w, torch.float64, requires_grad=True)
    tg = (torch.tensor(cat, dtype=torch.float64), torch.tensor(attr, dtype=torch.float64), torch.tensor(box, dtype=torch.float64), n)
    out = R.detr_call(pt, torch.tensor(feats, dtype=torch.float64), tg, NE, ND, 8, True, R.Dropout(13), R.model_weights(1.0), {})
    out["loss"].sum().backward()
    assert nerr(model.metric_tensors["loss"].cpu().numpy(), out["loss"].detach().numpy()) < 1e-5
    for g, r in zip(model.last_preds, out["preds"]):
        assert nerr(g.cpu().numpy(), r.detach().numpy()) < 1e-5
    g = model.get_grads_dict()
    gmax = max(float(v.grad.abs().max()) for v in pt.values() if v.grad is not None)
    for k, v in pt.items():
        if v.grad is not None and "KeyProjection/bias" not in k:
            assert nerr(g[k], v.grad.numpy(), 1e-6 * gmax) < 5e-4, k
    # inference returns the three prediction tensors; predict_indices the InverseTokenization tokens
    preds = model.call({"features": feats}, training=False)
    ref = R.detr_call(R.params_to_torch(model.get_weights_dict()), torch.tensor(feats, dtype=torch.float64), None, NE, ND, 8, False)
    for gq, r in zip(preds, ref["preds"]):
        assert nerr(gq.cpu().numpy(), r.numpy()) < 1e-5
