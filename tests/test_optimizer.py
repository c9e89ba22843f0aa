"""Optimizer + schedule: CosineDecayRestarts known answers on CPU; bdetr_sgd_step against the
float64 restatement on the GPU, with frozen blocks and variables that straddle several chunks."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200.optimizers import CHUNK, SGD, CosineDecayRestarts
from oracle import reference_path as R


def test_cosine_decay_restarts_known_answers():
    sched = CosineDecayRestarts(1e-3, 4000, m_mul=.95, alpha=0.1)          # the reference's schedule (notebook cell 26)
    # hand values: start of a period = lr0 * m_mul^i; end of period -> alpha floor; periods double (t_mul = 2)
    assert sched(0) == pytest.approx(1e-3, rel=1e-6)
    assert sched(2000) == pytest.approx(1e-3 * (0.9 * 0.5 + 0.1), rel=1e-5)           # half-way through period 0
    assert sched(3999) == pytest.approx(1e-4, rel=1e-3)
    assert sched(4000) == pytest.approx(1e-3 * (0.9 * 0.95 + 0.1), rel=1e-5)          # first restart, scaled by m_mul
    assert sched(4000 + 4000) == pytest.approx(1e-3 * (0.9 * 0.95 * 0.5 + 0.1), rel=1e-5)   # middle of period 1 (8000 long)
    assert sched(12000) == pytest.approx(1e-3 * (0.9 * 0.95 ** 2 + 0.1), rel=1e-5)    # second restart
    for step in (0, 1, 17, 3999, 4000, 4001, 11999, 12000, 27999, 28000, 123456):
        ref = R.cosine_decay_restarts(step, 1e-3, 4000, 2.0, .95, .1)
        assert sched(step) == pytest.approx(ref, rel=2e-4), step                     # float32 evaluation, like TF
    flat = CosineDecayRestarts(0.5, 10, t_mul=1.0, m_mul=1.0, alpha=0.0)
    assert flat(0) == pytest.approx(0.5) and flat(10) == pytest.approx(0.5) and flat(5) == pytest.approx(0.25, rel=1e-6)


def test_chunk_table_layout():
    slots = [("a", 0, 5), ("b", 8, CHUNK), ("c", 8 + CHUNK, 2 * CHUNK + 3)]
    tab = SGD.chunk_table(slots)
    assert tab.dtype.itemsize == 24                                              # matches bdetr_opt_chunk
    assert [int(r["len"]) for r in tab] == [5, CHUNK, CHUNK, CHUNK, 3]
    assert [int(r["var_first"]) for r in tab] == [0, 1, 2, 2, 2] and [int(r["var_chunks"]) for r in tab] == [1, 1, 3, 3, 3]
    assert int(tab[4]["offset"]) == 8 + CHUNK + 2 * CHUNK


def _small_model(seed=0, N=2, rows=5, cols=5, Q=100):
    from boosted_detr_b200.boosted_model import BoostedDETR
    from boosted_detr_b200.parameters import ModelParameters
    p = ModelParameters("COCO").default_params()
    p.pop("pad_value"); p.pop("oov_value")
    p.update(num_object_preds=Q, num_decoder_blocks=N, num_encoder_blocks=N, image_size=(rows * 32, cols * 32))
    return BoostedDETR(**p, attribute_weight=1.0, seed=seed).build()


@pytest.mark.gpu
@pytest.mark.parametrize("nesterov,clipnorm,momentum", [(True, 0.1, 0.9), (False, None, 0.5), (True, 5.0, 0.95)])
def test_sgd_step_vs_oracle(nesterov, clipnorm, momentum):
    import torch
    model = _small_model()
    # the reference's freezing schedule: only block 1 trains (notebook cell 30)
    for i in range(model.num_decoder_blocks):
        for group in (model.EncoderTransformerBlocks, model.DecoderBlocks, model.CategoryBlocks, model.AttributeBlocks, model.BoxBlocks):
            group[i].trainable = (i == 1)
    opt = SGD(learning_rate=CosineDecayRestarts(1e-2, 4, m_mul=.95, alpha=.1), momentum=momentum, nesterov=nesterov, clipnorm=clipnorm)
    model.compile(optimizer=opt)
    rng = np.random.default_rng(0)
    names = [n for n, o, k in model.named_weights() if n in model._index]
    trainable = {n for n, _, _ in SGD.trainable_slots(model)}
    assert trainable and all(("_1/" in n) or n.startswith("DecoderPrep") for n in trainable), sorted(trainable)[:5]
    assert any(model._index[n][1] > CHUNK for n in trainable)                    # a variable spanning several chunks
    w = {n: v.astype(np.float64) for n, v in model.get_weights_dict().items() if n in model._index}
    acc = {n: np.zeros_like(w[n]) for n in names}
    for step in range(3):
        scale = 10.0 ** rng.uniform(-3, 1)                                        # norms on both sides of clipnorm
        g = {n: (rng.standard_normal(w[n].shape) * scale / np.sqrt(w[n].size)).astype(np.float32) for n in names}
        for n, o, k in model.named_weights():
            if n in g:
                o._grads[k].copy_(torch.from_numpy(g[n]).cuda())
        lr = R.cosine_decay_restarts(step, 1e-2, 4, 2.0, .95, .1)
        new_w, new_a = R.sgd_step_reference({n: w[n] for n in trainable}, g, acc, lr, momentum, nesterov, clipnorm)
        w.update(new_w); acc.update(new_a)
        opt.apply(model)
        torch.cuda.synchronize()
        assert opt.last_lr == pytest.approx(lr, rel=2e-4)
        got = model.get_weights_dict()
        for n in names:
            if n in trainable:
                err = np.abs(got[n] - w[n]).max() / max(np.abs(w[n]).max(), 1e-12)
                assert err < 1e-6, (n, step, err)                                  # float32 arithmetic vs float64 restatement
            else:
                assert (got[n] == w[n].astype(np.float32)).all(), n                # frozen: bitwise untouched
    assert opt.iterations == 3


@pytest.mark.gpu
def test_fit_with_optimizer_reduces_loss():
    """compile(optimizer) + fit(): the loss of a fixed batch goes down over SGD steps (whole step: fwd, matcher, bwd,
    optimizer), through the eager train_step and through the CUDA-graph step."""
    from util import synth_targets
    from boosted_detr_b200.graph import GraphedTrainStep
    model = _small_model(seed=3)
    rng = np.random.default_rng(1)
    cat, attr, box, n = synth_targets(rng, 4, 6, model.num_categories, model.num_attributes, attr_p=0.05)
    feats = np.tanh(rng.standard_normal((4, 5, 5, 256))).astype(np.float32)
    batch = {"features": feats, "category": cat, "attribute": attr, "bbox": box, "num_objects": n}
    model.dropout_seed = None
    model.compile(optimizer=SGD(learning_rate=CosineDecayRestarts(1e-3, 4000, m_mul=.95, alpha=.1), momentum=.9, nesterov=True, clipnorm=0.1))
    hist = model.fit([batch] * 10, epochs=2)
    first, last = model.train_step(batch)["loss"], None
    assert np.isfinite(hist["loss"]).all()
    assert hist["loss"][1] < hist["loss"][0]
    step = GraphedTrainStep(model, batch)
    before = model.optimizer.iterations
    losses = [step(batch)["loss"] for _ in range(10)]
    assert model.optimizer.iterations == before + 10
    assert losses[-1] < first, (first, losses)
