"""BackboneNeck: BatchNorm -> 1x1 Conv2D + tanh -> BatchNorm as one tcgen05 GEMM with folded
normalisations, against the fp64 oracle (tensor-core mode: 1e-3-class tolerances, north_star's reduced-precision bar)."""
import numpy as np
import pytest
import torch

from test_dense_gpu import nerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("training", [True, False])
def test_backbone_neck_vs_oracle(training):
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.backbone import BackboneNeck
    from boosted_detr_b200.layers import Layer
    from oracle import reference_path as R
    lib = _lib.load()
    lib.bdetr_set_mode(_lib.MODE_TF32)
    try:
        rng = np.random.default_rng(11)
        B, Rr, Cc, Cin, N = 4, 12, 12, 1792, 256                      # EfficientNetB4 top feature width (backbone.py)
        Layer._rng = np.random.default_rng(11)
        x = (rng.standard_normal((B, Rr, Cc, Cin)) * rng.uniform(0.5, 2.0, Cin) + rng.normal(0, 0.5, Cin)).astype(np.float32)
        go = rng.standard_normal((B, Rr, Cc, N)).astype(np.float32)
        neck = BackboneNeck(N)
        dx = torch.from_numpy(x).cuda()
        neck.forward([dx], training=False)                             # builds
        for n_, o, k in neck.named_weights():                          # non-trivial affine parameters / moving statistics
            if k.endswith(("gamma", "moving_variance")):
                o._weights[k].copy_(torch.from_numpy(rng.uniform(0.5, 1.5, o._weights[k].shape).astype(np.float32)))
            elif k.endswith(("beta", "bias", "moving_mean")):
                o._weights[k].copy_(torch.from_numpy(rng.normal(0, 0.2, o._weights[k].shape).astype(np.float32)))
        w = {n_: o._weights[k].cpu().numpy().copy() for n_, o, k in neck.named_weights()}
        out, ctx = neck.forward([dx], training=training, round_out=False)
        neck.backward(ctx, torch.from_numpy(go).cuda())
        torch.cuda.synchronize()
        p = R.params_to_torch(w, torch.float64, requires_grad=True)
        stats = {}
        ref = R.backbone_neck(torch.tensor(x, dtype=torch.float64), p, "BackboneNeck", training, stats)
        (ref * torch.tensor(go, dtype=torch.float64)).sum().backward()
        e = nerr(out.cpu().numpy(), ref.detach().numpy())
        print(f"neck forward (training={training}): {e:.2e}")
        assert e < 2e-3
        after = {n_: o._weights[k].cpu().numpy() for n_, o, k in neck.named_weights()}
        for k_, v in stats.items():
            assert nerr(after[k_], v.numpy() if hasattr(v, "numpy") else v) < 1e-4, k_
        if not training:
            assert all((after[k_] == w[k_]).all() for k_ in w if "moving" in k_)
        for n_, o, k in neck.named_weights():
            if k in o._grads:
                r = p[n_].grad.numpy()
                g = o._grads[k].cpu().numpy().reshape(r.shape)
                rel = float(np.sqrt(((g - r) ** 2).sum() / max((r ** 2).sum(), 1e-30)))
                print(f"  grad {k}: relative L2 {rel:.2e}")
                assert rel < 2e-2, k
    finally:
        lib.bdetr_set_mode(_lib.MODE_FP32)


def test_model_with_backbone_neck_trains_through_it():
    """BoostedDETR(backbone_neck=True): the model starts at the backbone output; loss and the neck's parameter gradients
    (through block 0's input gradient) against the fp64 oracle with the model's own assignments."""
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.boosted_model import BoostedDETR
    from boosted_detr_b200.parameters import baseline_params
    from oracle import reference_path as R
    from util import synth_targets
    lib = _lib.load()
    lib.bdetr_set_mode(_lib.MODE_TF32)
    try:
        N, B, T, Q = 2, 4, 8, 100
        p = baseline_params(1)
        p.update(image_size=(12 * 32, 12 * 32))
        model = BoostedDETR(**p, attribute_weight=1.0, seed=3, backbone_neck=True, backbone_channels=256 + 64).build()
        rng = np.random.default_rng(9)
        wq = model.get_weights_dict()
        wq["DecoderPrep/init_decoder_features"] = rng.normal(0, 0.5, wq["DecoderPrep/init_decoder_features"].shape).astype(np.float32)
        model.set_weights_dict(wq)
        w = model.get_weights_dict()
        cat, attr, box, n = synth_targets(rng, B, T, model.num_categories, model.num_attributes, attr_p=0.05)
        bf = rng.standard_normal((B, 12, 12, 320)).astype(np.float32)
        model.dropout_seed = None
        model.train_step({"backbone_features": bf, "category": cat, "attribute": attr, "bbox": box, "num_objects": n})
        torch.cuda.synchronize()
        masks = []
        for c in model.last_ctx_train["loss"]:
            c4r = c["col4row"].cpu().numpy()
            m = np.zeros((B, T, Q))
            bb, tt = np.nonzero(c4r >= 0)
            m[bb, tt, c4r[bb, tt]] = 1.0
            masks.append(torch.from_numpy(m))
        out, grads, _ = R.train_step_reference(w, bf, (cat, attr, box, n), N, 8, torch.float64, weights=R.model_weights(1.0), forced_masks=masks)
        assert nerr(model.metric_tensors["loss"].cpu().numpy(), out["loss"].detach().numpy()) < 2e-3
        g = model.get_grads_dict()
        for k in grads:
            if k.startswith("BackboneNeck/"):
                r = grads[k]
                rel = float(np.sqrt(((g[k].reshape(r.shape) - r) ** 2).sum() / max((r ** 2).sum(), 1e-30)))
                print(f"  {k}: relative L2 {rel:.2e}")
                assert rel < 5e-2, k
    finally:
        lib.bdetr_set_mode(_lib.MODE_FP32)
