"""GPU parity of the dense path (attention / FFN blocks, heads, full boosted model forward + loss +
gradients) against the fp64 oracle.  fp32 mode tolerance: 1e-5 normalised max error on predictions and
losses (north_star); gradients are checked per tensor with the tolerance written at the assert."""
import numpy as np
import pytest
import torch

from util import synth_targets

pytestmark = pytest.mark.gpu


def nerr(got, ref, floor=1e-30):
    """normalised max error |got-ref|_inf / max(|ref|_inf, floor)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), floor))


def _perturb(model, rng):
    """Randomise every weight so that no gradient path is hidden by zero biases / unit gammas."""
    d = model.get_weights_dict()
    for k, v in d.items():
        if k.endswith("/bias") or k.endswith("/beta"):
            d[k] = rng.normal(0, 0.1, v.shape).astype(np.float32)
        elif k.endswith("/gamma"):
            d[k] = (1.0 + rng.normal(0, 0.1, v.shape)).astype(np.float32)
        elif k.endswith("moving_mean"):
            d[k] = rng.normal(0, 0.2, v.shape).astype(np.float32)
        elif k.endswith("moving_variance"):
            d[k] = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
        elif k.endswith("init_decoder_features"):
            d[k] = rng.normal(0, 0.5, v.shape).astype(np.float32)
    model.set_weights_dict(d)
    return d


def test_gemm_variants():
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.device import ptr, stream_ptr
    rng = np.random.default_rng(0)
    for (M, N, K, ta, tb, bias, act, beta) in [(6400, 256, 256, 0, 0, 1, 1, 0), (1600, 82, 256, 0, 0, 1, 0, 0),
                                               (256, 256, 6400, 1, 0, 0, 0, 1), (1600, 256, 82, 0, 1, 0, 0, 0),
                                               (100, 3, 256, 0, 0, 1, 0, 1), (256, 4, 1600, 1, 0, 0, 0, 0),
                                               (77, 65, 33, 0, 1, 0, 0, 0), (130, 70, 1000, 1, 1, 1, 0, 0)]:
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        Bm = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        bv = rng.standard_normal(N).astype(np.float32)
        C0 = rng.standard_normal((M, N)).astype(np.float32)
        ref = (A.T if ta else A).astype(np.float64) @ (Bm.T if tb else Bm).astype(np.float64)
        if bias:
            ref = ref + bv
        if beta:
            ref = ref + C0
        if act:
            ref = np.maximum(ref, 0)
        dA, dB, db, dC = (torch.from_numpy(x).cuda() for x in (A, Bm, bv, C0.copy()))
        _lib.call("bdetr_gemm", M, N, K, ptr(dA), ta, ptr(dB), tb, ptr(db) if bias else None, act, beta, ptr(dC), stream_ptr())
        e = nerr(dC.cpu().numpy(), ref)
        print(f"gemm M{M} N{N} K{K} ta{ta} tb{tb}: {e:.2e}")
        assert e < 2e-6


@pytest.mark.parametrize("Lq,Lk,selfattn,dropout", [(400, 400, True, False), (100, 400, False, False),
                                                    (100, 100, True, True), (37, 70, False, True)])
def test_attention_block_vs_oracle(Lq, Lk, selfattn, dropout):
    from oracle import reference_path as R
    from boosted_detr_b200.layers import Layer, dropout_key
    from boosted_detr_b200.transformers import AttentionBlock
    rng = np.random.default_rng(Lq + Lk)
    Layer._rng = np.random.default_rng(1)
    B, D, H = 2, 256, 8
    q = rng.standard_normal((B, Lq, D)).astype(np.float32)
    k = q if selfattn else rng.standard_normal((B, Lk, D)).astype(np.float32)
    v = rng.standard_normal((B, Lk, D)).astype(np.float32)
    go = rng.standard_normal((B, Lq, D)).astype(np.float32)
    blk = AttentionBlock(H, name="blk")
    dq = torch.from_numpy(q).cuda()
    dk = dq if selfattn else torch.from_numpy(k).cuda()
    dv = torch.from_numpy(v).cuda()
    seed, site = 1234, 11
    key = dropout_key(seed, site) if dropout else 0
    assert key == (R.dropout_key(seed, site) if dropout else 0)
    out, ctx = blk.forward([dq, dk, dv], training=dropout, dropout_key=key)
    w = {n[len("blk/"):]: o._weights[kk].cpu().numpy() for n, o, kk in blk.named_weights()}
    # randomise biases / LN after build, then rerun
    for n, o, kk in blk.named_weights():
        if kk.endswith("bias") or kk.endswith("beta"):
            o._weights[kk].copy_(torch.from_numpy(rng.normal(0, 0.1, o._weights[kk].shape).astype(np.float32)))
        if kk.endswith("gamma"):
            o._weights[kk].copy_(torch.from_numpy((1 + rng.normal(0, 0.1, o._weights[kk].shape)).astype(np.float32)))
    out, ctx = blk.forward([dq, dk, dv], training=dropout, dropout_key=key)
    d_q, d_k, d_v = blk.backward(ctx, torch.from_numpy(go).cuda())
    w = {n[len("blk/"):]: o._weights[kk].cpu().numpy() for n, o, kk in blk.named_weights()}
    p = R.params_to_torch({"p/" + n: a for n, a in w.items()}, torch.float64, requires_grad=True)
    tq = torch.tensor(q, dtype=torch.float64, requires_grad=True)
    tk = tq if selfattn else torch.tensor(k, dtype=torch.float64, requires_grad=True)
    tv = torch.tensor(v, dtype=torch.float64, requires_grad=True)
    ref = R.attention_block(tq, tk, tv, p, "p", H, R.Dropout(seed if dropout else None), site, dropout)
    (ref * torch.tensor(go, dtype=torch.float64)).sum().backward()
    e = nerr(out.cpu().numpy(), ref.detach().numpy())
    print(f"attention block out: {e:.2e}")
    assert e < 1e-5
    errs = {"d_query": nerr(d_q.cpu().numpy(), tq.grad.numpy()), "d_value": nerr(d_v.cpu().numpy(), tv.grad.numpy())}
    if not selfattn:
        errs["d_key"] = nerr(d_k.cpu().numpy(), tk.grad.numpy())
    for n, o, kk in blk.named_weights():
        # the key-bias gradient is mathematically zero (softmax is shift invariant): use the query-bias scale
        floor = float(p["p/AttentionLayer/QueryProjection/bias"].grad.abs().max()) if "KeyProjection/bias" in n else 1e-30
        errs[n] = nerr(o._grads[kk].cpu().numpy(), p["p/" + n[len("blk/"):]].grad.numpy(), floor)
    for n, e in errs.items():
        print(f"  {n}: {e:.2e}")
    assert max(errs.values()) < 2e-5       # fp32 accumulation over up to 800 rows vs fp64


def test_ffn_block_vs_oracle():
    from oracle import reference_path as R
    from boosted_detr_b200.layers import Layer, dropout_key
    from boosted_detr_b200.transformers import FeedForwardBlock
    rng = np.random.default_rng(3)
    Layer._rng = np.random.default_rng(2)
    B, L, D = 3, 100, 256
    x = rng.standard_normal((B, L, D)).astype(np.float32)
    go = rng.standard_normal((B, L, D)).astype(np.float32)
    blk = FeedForwardBlock(name="ffn")
    seed, site = 99, 4
    key = dropout_key(seed, site)
    dx_in = torch.from_numpy(x).cuda()
    out, ctx = blk.forward([dx_in], training=True, dropout_key=key)
    for n, o, kk in blk.named_weights():
        if kk.endswith("bias") or kk.endswith("beta"):
            o._weights[kk].copy_(torch.from_numpy(rng.normal(0, 0.1, o._weights[kk].shape).astype(np.float32)))
    out, ctx = blk.forward([dx_in], training=True, dropout_key=key)
    d_x = blk.backward(ctx, torch.from_numpy(go).cuda())
    p = R.params_to_torch({"p/" + n[len("ffn/"):]: o._weights[kk].cpu().numpy() for n, o, kk in blk.named_weights()},
                          torch.float64, requires_grad=True)
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ref = R.feed_forward_block(tx, p, "p", R.Dropout(seed), site, True)
    (ref * torch.tensor(go, dtype=torch.float64)).sum().backward()
    assert nerr(out.cpu().numpy(), ref.detach().numpy()) < 1e-5
    errs = {"d_x": nerr(d_x.cpu().numpy(), tx.grad.numpy())}
    for n, o, kk in blk.named_weights():
        errs[n] = nerr(o._grads[kk].cpu().numpy(), p["p/" + n[len("ffn/"):]].grad.numpy())
    print(errs)
    assert max(errs.values()) < 2e-5


@pytest.mark.parametrize("kind,nout", [("cat", 82), ("attr", 3), ("box", 4), ("attr", 296)])
def test_head_vs_oracle(kind, nout):
    from oracle import reference_path as R
    from boosted_detr_b200.layers import Layer
    from boosted_detr_b200 import prediction_heads as PH
    rng = np.random.default_rng(5)
    Layer._rng = np.random.default_rng(3)
    B, Q, D = 4, 100, 256
    x = rng.standard_normal((B, Q, D)).astype(np.float32)
    gcum = rng.standard_normal((B, Q, nout)).astype(np.float32)
    prev = rng.random((B, Q, nout)).astype(np.float32)
    head = {"cat": lambda: PH.SingleClassPredictionHead(nout, D, Q, name="h"),
            "attr": lambda: PH.MultiClassPredictionHead(nout, D, Q, name="h"),
            "box": lambda: PH.BoxPredictionHead(D, Q, name="h")}[kind]()
    fn = {"cat": R.category_head, "attr": R.attribute_head, "box": R.box_head}[kind]
    dx_in = torch.from_numpy(x).cuda()
    head.forward([dx_in], training=False)          # build
    for n, o, kk in head.named_weights():
        if kk.endswith("bias") or kk.endswith("beta"):
            o._weights[kk].copy_(torch.from_numpy(rng.normal(0, 0.1, o._weights[kk].shape).astype(np.float32)))
    w0 = {"p/" + n[len("h/"):]: o._weights[kk].cpu().numpy().copy() for n, o, kk in head.named_weights()}
    # inference (moving statistics)
    act, _ = head.forward([dx_in], training=False)
    ref = fn(torch.tensor(x, dtype=torch.float64), R.params_to_torch(w0), "p", False)
    assert nerr(act.cpu().numpy(), ref.numpy()) < 1e-5
    # training (batch statistics), accumulate into a running prediction with mult 2
    cum = torch.from_numpy(prev.copy()).cuda()
    act, ctx = head.forward([dx_in], training=True, cum=cum, mult=2.0)
    d_x = head.backward(ctx, torch.from_numpy(gcum).cuda())
    p = R.params_to_torch(w0, torch.float64, requires_grad=True)
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    stats = {}
    ref = fn(tx, p, "p", True, stats)
    refcum = torch.tensor(prev, dtype=torch.float64) + 2.0 * ref
    (refcum * torch.tensor(gcum, dtype=torch.float64)).sum().backward()
    assert nerr(act.cpu().numpy(), ref.detach().numpy()) < 1e-5
    assert nerr(cum.cpu().numpy(), refcum.detach().numpy()) < 1e-5
    errs = {"d_x": nerr(d_x.cpu().numpy(), tx.grad.numpy())}
    for n, o, kk in head.named_weights():
        key = "p/" + n[len("h/"):]
        if kk in o._grads:
            errs[n] = nerr(o._grads[kk].cpu().numpy(), p[key].grad.numpy())
        else:
            errs[n + " (moving)"] = nerr(o._weights[kk].cpu().numpy(), stats[key])
    print(kind, errs)
    assert max(errs.values()) < 5e-5


def _model_and_data(N, B, rows, cols, C_extra=None, Q=100, T=20, seed=0, dataset="COCO", attribute_weight=1.0):
    from boosted_detr_b200.boosted_model import BoostedDETR
    from boosted_detr_b200.parameters import ModelParameters
    p = ModelParameters(dataset).default_params()
    p.pop("pad_value"); p.pop("oov_value")
    p.update(num_object_preds=Q, num_decoder_blocks=N, num_encoder_blocks=N, image_size=(rows * 32, cols * 32))
    model = BoostedDETR(**p, attribute_weight=attribute_weight, seed=seed).build()
    rng = np.random.default_rng(seed + 1)
    w = _perturb(model, rng)
    C, A = model.num_categories, model.num_attributes
    cat, attr, box, n = synth_targets(rng, B, T, C, A, attr_p=0.05)
    feats = np.tanh(rng.standard_normal((B, rows, cols, 256))).astype(np.float32)
    inputs = {"features": feats, "category": cat, "attribute": attr, "bbox": box, "num_objects": n}
    return model, model.get_weights_dict(), inputs


def test_model_inference_vs_oracle():
    from oracle import reference_path as R
    model, w, inputs = _model_and_data(N=2, B=2, rows=20, cols=20)
    got = model.call({"features": inputs["features"]}, training=False)
    out = R.boosted_detr_call(R.params_to_torch(w), torch.tensor(inputs["features"], dtype=torch.float64), None,
                              2, 8, training=False)
    for g, r, name in zip(got, out["preds"], ["cat", "attr", "box"]):
        e = nerr(g.cpu().numpy(), r.numpy())
        print(f"inference {name}: {e:.2e}")
        assert e < 1e-5


@pytest.mark.parametrize("N,B,rows,cols,dropout,dataset", [(2, 2, 20, 20, False, "COCO"), (3, 3, 6, 7, True, "COCO"),
                                                           (2, 2, 5, 5, False, "Fashionpedia")])
def test_model_train_step_vs_oracle(N, B, rows, cols, dropout, dataset):
    """BASELINE config 1 (2 block pairs, batch 2, 20x20 features, 100 queries): forward + Hungarian loss +
    gradients of the summed loss vector, against the fp64 oracle."""
    from oracle import reference_path as R
    model, w, inputs = _model_and_data(N=N, B=B, rows=rows, cols=cols, dataset=dataset)
    model.dropout_seed = 777 if dropout else None
    logs = model.train_step(inputs)
    ctx = model.last_ctx if hasattr(model, "last_ctx") else None
    tg = (inputs["category"], inputs["attribute"], inputs["bbox"], inputs["num_objects"])
    out, grads, stats = R.train_step_reference(w, inputs["features"], tg, N, 8, torch.float64,
                                               dropout_seed=777 if dropout else None,
                                               weights=R.model_weights(1.0))
    m = model.metric_tensors
    ref_loss = out["loss"].detach().numpy()
    e = nerr(m["loss"].cpu().numpy(), ref_loss)
    print(f"loss vector: {e:.2e}  (mean {ref_loss.mean():.4f}, keras-style logs {logs})")
    assert e < 1e-5
    for k in ["Category_Loss", "Attribute_Loss", "Box_Loss", "Existence_Loss"]:
        assert nerr(m[k].cpu().numpy(), out["metrics"][k].detach().numpy()) < 1e-5, k
    assert nerr(m["IOU"].cpu().numpy(), out["metrics"]["IOU"].detach().numpy()) < 1e-4
    g = model.get_grads_dict()
    assert set(g) == set(grads)
    # Tolerance: the loss gradient is ill-conditioned in fp32 wherever a cumulative probability sits next
    # to the .999 clip (d/dp ~ 1/(1-p)), so the yardstick is the oracle itself evaluated in fp32 with the
    # same assignment: GPU error <= max(2e-5, 5 x fp32-oracle error), both measured against fp64.
    _, grads32, _ = R.train_step_reference(w, inputs["features"], tg, N, 8, torch.float32,
                                           dropout_seed=777 if dropout else None, weights=R.model_weights(1.0),
                                           forced_masks=[m_.to(torch.float32) for m_ in out["masks"]])
    gmax = max(float(np.abs(v).max()) for v in grads.values())
    worst = []
    for k, ref in grads.items():
        floor = 1e-6 * gmax
        if k.endswith("KeyProjection/bias"):
            # mathematically zero (softmax is invariant to a key bias): only rounding noise, judged on the
            # scale of the same layer's query-bias gradient
            floor = float(np.abs(grads[k.replace("KeyProjection", "QueryProjection")]).max())
        worst.append((nerr(g[k], ref, floor), nerr(grads32[k], ref, floor), k, float(np.abs(ref).max())))
    worst.sort(reverse=True)
    for e, e32, k, mag in worst[:8]:
        print(f"  grad {k}: gpu {e:.2e} | fp32 oracle {e32:.2e} (max |ref| {mag:.2e})")
    # AttributePredictionHead_* and everything upstream of it inherit the clip ill-conditioning: a 2.5e-7 error
    # in a cumulative probability next to .999 (fp32 epsilon) is amplified ~1000x.  Layer-level tests above hold
    # the well-conditioned 2e-5 bar; here the end-to-end bar is 5e-4.
    for e, e32, k, mag in worst:
        assert e < max(5e-4, 5 * e32), k
    wd = model.get_weights_dict()
    for k, ref in stats.items():
        assert nerr(wd[k], ref) < 1e-5, k


def test_gradient_buckets_follow_backward_order():
    """Data-parallel overlap: the flat gradient buffer is one contiguous bucket per boosted block, in the order the
    backward finishes them, and the bucket hook fires once per block with exactly those ranges; at that point on the
    caller's stream + the passed events the bucket already holds its final values."""
    N = 3
    model, w, inputs = _model_and_data(N=N, B=2, rows=4, cols=4, Q=8, T=4)
    flat_g = model._flat[1]
    assert [b for b, _, _ in model._buckets] == list(range(N - 1, -1, -1))
    assert model._buckets[0][1] == 0 and model._buckets[-1][2] == flat_g.numel()
    assert all(model._buckets[k][2] == model._buckets[k + 1][1] for k in range(N - 1))
    for name, (off, cnt, _) in model._index.items():
        blk = int(name.split("/")[0].rsplit("_", 1)[1]) if name.split("/")[0].rsplit("_", 1)[-1].isdigit() else 0
        lo, hi = next((l, h) for b, l, h in model._buckets if b == blk)
        assert lo <= off and off + cnt <= hi, name
    calls, snaps = [], {}
    side = torch.cuda.Stream()

    def hook(block, lo, hi, events):
        calls.append((block, lo, hi))
        side.wait_stream(torch.cuda.current_stream())
        for ev in events:
            side.wait_event(ev)
        with torch.cuda.stream(side):
            snaps[block] = flat_g[lo:hi].clone()

    model.grad_bucket_hook = hook
    model.train_step(inputs)
    torch.cuda.synchronize()
    assert calls == [(b, lo, hi) for b, lo, hi in model._buckets]
    for b, lo, hi in model._buckets:
        assert torch.equal(snaps[b], flat_g[lo:hi]), f"bucket of block {b} was still being written when its hook fired"
        assert float(snaps[b].abs().sum()) > 0
