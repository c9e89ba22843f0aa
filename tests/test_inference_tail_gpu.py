"""Inference tail: the numeric half of InverseTokenization (reference tokenizers.py:126-137) and the
confidence early exit across boosted blocks (the reference's TODO, README.md:9)."""
import numpy as np
import pytest
import torch

from test_dense_gpu import _model_and_data, nerr

pytestmark = pytest.mark.gpu


def test_inverse_tokenization_matches_tf_semantics():
    from boosted_detr_b200.tokenizers import InverseTokenization
    rng = np.random.default_rng(3)
    B, Q, C, A = 5, 37, 82, 9
    cat = rng.random((B, Q, C)).astype(np.float32)
    cat[0, 0, :] = 0.25                       # all tied -> first index
    cat[0, 1, 40] = cat[0, 1, 70] = 2.0       # two maxima -> the first
    cat[1, 2, 81] = 3.0                       # last column
    attr = rng.random((B, Q, A)).astype(np.float32)
    attr[0, 0, 3] = 0.5                       # >= .5 is inclusive
    vocab = {"category": [f"c{i}" for i in range(C - 2)], "attribute": [f"a{i}" for i in range(A - 2)]}
    inv = InverseTokenization(vocab)
    tok_c, tok_a, conf, iconf = inv.tokens([torch.from_numpy(cat).cuda(), torch.from_numpy(attr).cuda()], conf_scale=0.5, want_confidence=True)
    torch.cuda.synchronize()
    assert (tok_c.squeeze(-1).cpu().numpy() == cat.argmax(-1)).all()            # numpy argmax = first maximum, like tf.argmax
    assert (tok_a.cpu().numpy() == (attr >= 0.5).astype(np.int32) * np.arange(A, dtype=np.int32)).all()
    assert np.array_equal(conf.cpu().numpy(), cat.max(-1) * np.float32(0.5))
    assert np.array_equal(iconf.cpu().numpy(), (cat.max(-1) * np.float32(0.5)).min(-1))
    cats, attrs = inv.sparce_to_strings(tok_c, tok_a)
    assert cats[0][0] == "<PAD>" and cats[1][2] == "c79" and attrs[0][0][3] == "a1" and attrs[0][0][0] == "<PAD>"


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_early_exit_returns_the_running_prediction_of_the_exit_block(mode):
    from boosted_detr_b200 import _lib
    from oracle import reference_path as R
    lib = _lib.load()
    lib.bdetr_set_mode(_lib.MODE_TF32 if mode == "tf32" else _lib.MODE_FP32)
    try:
        N = 4
        model, w, inputs = _model_and_data(N=N, B=2, rows=20, cols=20)
        feats = {"features": inputs["features"]}
        out = R.boosted_detr_call(R.params_to_torch(w), torch.tensor(inputs["features"], dtype=torch.float64), None, N, 8, training=False)
        tol = 1e-5 if mode == "fp32" else 2e-3
        # never confident enough: all blocks run
        model.early_exit_threshold, model.early_exit_min_blocks = 2.0, 1
        full = [t.cpu().numpy() for t in model.call(feats, training=False)]
        assert model.last_exit_block == N - 1
        for g, r in zip(full, out["preds"]):
            assert nerr(g, r.numpy()) < tol
        conf_ref = [(p[0].numpy().max(-1) / (i + 2)).min(-1) for i, p in enumerate(out["per_block_preds"])]      # [N][B]
        # always confident: stop right after the minimum number of blocks
        for min_blocks in (1, 3):
            model.early_exit_threshold, model.early_exit_min_blocks = 0.0, min_blocks
            early = [t.cpu().numpy() for t in model.call(feats, training=False)]
            assert model.last_exit_block == min_blocks - 1
            for g, r in zip(early, out["per_block_preds"][min_blocks - 1]):
                assert nerr(g, r.numpy()) < tol
            assert nerr(model.last_image_confidence.numpy(), conf_ref[min_blocks - 1]) < 10 * tol
        # a threshold between the confidences of two consecutive blocks picks the block the oracle picks
        mins = [float(c.min()) for c in conf_ref]
        order = [i for i in range(N - 1)]
        thr = 0.5 * (mins[0] + mins[1]) if abs(mins[0] - mins[1]) > 1e-3 else None
        if thr is not None:
            model.early_exit_threshold, model.early_exit_min_blocks = thr, 1
            model.call(feats, training=False)
            expect = next((i for i in order if mins[i] >= thr), N - 1)
            assert model.last_exit_block == expect
    finally:
        lib.bdetr_set_mode(_lib.MODE_FP32)


def test_checkpoint_round_trip_by_variable_name(tmp_path):
    """save_weights / load_weights: both name spellings restore every variable bit for bit, a
    checkpoint with extra entries (optimizer slots, backbone) loads, a missing variable or a wrong shape is an error."""
    from boosted_detr_b200.checkpoint import keras_to_object_graph
    model, w, inputs = _model_and_data(N=2, B=2, rows=4, cols=4, Q=8, T=4)
    ref = {k: v.copy() for k, v in model.get_weights_dict().items()}
    for naming in ("keras", "object_graph"):
        path = model.save_weights(str(tmp_path / f"ckpt_{naming}"), naming=naming)
        model.set_weights_dict({k: np.zeros_like(v) for k, v in ref.items()})
        names = model.load_weights(path)
        assert set(names) == set(ref)
        got = model.get_weights_dict()
        assert all((got[k] == ref[k]).all() for k in ref)
    extra = {keras_to_object_graph(k): v for k, v in ref.items()}
    extra["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"] = np.zeros(1)
    extra["EncoderBackbone/stem_conv/kernel/.ATTRIBUTES/VARIABLE_VALUE"] = np.zeros((3, 3, 3, 48), np.float32)
    model.load_weights(extra)
    k0 = next(iter(ref))
    with pytest.raises(KeyError):
        model.load_weights({k: v for k, v in ref.items() if k != k0})
    with pytest.raises(ValueError):
        model.load_weights({**ref, k0: np.zeros((3, 3), np.float32)})
    # the restored model computes what it computed before
    a = [t.cpu().numpy() for t in model.call({"features": inputs["features"]}, training=False)]
    model.load_weights(ref)
    b = [t.cpu().numpy() for t in model.call({"features": inputs["features"]}, training=False)]
    assert all((x == y).all() for x, y in zip(a, b))
