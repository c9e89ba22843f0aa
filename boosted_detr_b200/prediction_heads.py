"""Per-block prediction heads — same class names / constructor arguments / call conventions as the
reference's ModelComponents/prediction_heads.py.  Each head is Dense(relu) -> BatchNorm -> Dense ->
activation; the post-activation output is optionally added into the boosted running prediction inside
the kernel (boosted_model.py:222-229).  The reference's Conv1D/Permute branch (:53-56) only runs when
the number of incoming predictions differs from `num_preds`, which never happens on this path; it is
rejected explicitly instead of being silently skipped."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .device import empty, f32, ptr, stream_ptr
from .layers import Layer, glorot_normal, he_normal
from .transformers import _struct

BN_EPS = 1e-3
BN_MOMENTUM = 0.99

KIND_SOFTMAX, KIND_SIGMOID, KIND_BOX = 0, 1, 2


class _Head(Layer):
    kind = None
    dense1 = "Dense"
    dense2 = None

    def __init__(self, num_out, hidden_dim, num_preds, name):
        super().__init__(name=name)
        self.num_out = num_out
        self.hidden_dim = hidden_dim
        self.num_preds = num_preds

    def build(self, input_shape):
        D = input_shape[0][-1]
        rng = Layer._rng
        self.add_weight(f"{self.dense1}/kernel", he_normal(rng, D, self.hidden_dim))
        self.add_weight(f"{self.dense1}/bias", np.zeros(self.hidden_dim, np.float32))
        self.add_weight("BatchNorm/gamma", np.ones(self.hidden_dim, np.float32))
        self.add_weight("BatchNorm/beta", np.zeros(self.hidden_dim, np.float32))
        self.add_weight("BatchNorm/moving_mean", np.zeros(self.hidden_dim, np.float32), trainable=False)
        self.add_weight("BatchNorm/moving_variance", np.ones(self.hidden_dim, np.float32), trainable=False)
        self.add_weight(f"{self.dense2}/kernel", glorot_normal(rng, self.hidden_dim, self.num_out))
        self.add_weight(f"{self.dense2}/bias", np.zeros(self.num_out, np.float32))

    def _structs(self):
        w, mode = self.gemm_weights()
        if self._struct_cache is None or self._struct_cache[2] != mode:
            def pack(s, stats):
                return _struct(_lib.HeadParams, {
                    "w1": s[f"{self.dense1}/kernel"], "b1": s[f"{self.dense1}/bias"],
                    "bn_gamma": s["BatchNorm/gamma"], "bn_beta": s["BatchNorm/beta"],
                    "bn_moving_mean": stats.get("BatchNorm/moving_mean"), "bn_moving_var": stats.get("BatchNorm/moving_variance"),
                    "w2": s[f"{self.dense2}/kernel"], "b2": s[f"{self.dense2}/bias"]})
            self._struct_cache = (pack(w, self._weights), pack(self._grads, {}), mode)
        return self._struct_cache[:2]

    def forward(self, inputs, training=False, cum=None, mult=1.0):
        """Returns (post-activation prediction of THIS head, ctx).  If `cum` is given the kernel also does
        cum += mult * prediction in place; with cum=None a fresh buffer receives mult * prediction."""
        x = f32(inputs[0])
        self.maybe_build([x])
        D = x.shape[-1]
        if x.dim() != 3 or x.shape[1] != self.num_preds:
            raise ValueError(f"{self.name}: expected [batch, {self.num_preds}, dim] features; the reference's "
                             "Conv1D re-projection of the prediction axis is not on the supported path")
        M = x.shape[0] * x.shape[1]
        Dh, N = self.hidden_dim, self.num_out
        sv = {"h": empty(M, Dh), "hn": empty(M, Dh), "bn_mean": empty(Dh), "bn_rstd": empty(Dh), "bn_acc": empty(2 * Dh * ((M + 127) // 128)),
              "act": empty(x.shape[0], x.shape[1], N)}
        init = cum is None
        if init:
            cum = empty(x.shape[0], x.shape[1], N)
        w, _ = self._structs()
        svs = _struct(_lib.HeadSaved, sv)
        # Keras: a BatchNormalization whose layer is frozen (trainable = False, reference notebook cell 30) runs in
        # inference mode -- moving statistics, no update -- even inside a training step
        bn_training = 1 if (training and self.trainable) else 0
        _lib.call("bdetr_head_fwd", M, D, Dh, N, self.kind, bn_training, float(mult), ptr(x), ctypes.byref(w),
                  BN_EPS, BN_MOMENTUM, ptr(cum), 1 if init else 0, ctypes.byref(svs), stream_ptr())
        ctx = {"x": x, "saved": sv, "saved_struct": svs, "dims": (M, D, Dh, N), "mult": float(mult), "cum": cum,
               "bn_training": bn_training}
        return sv["act"], ctx

    def backward(self, ctx, d_cum, d_x=None, acc=False):
        """d_cum: gradient w.r.t. the running prediction this head's output was added into."""
        M, D, Dh, N = ctx["dims"]
        if d_x is None:
            d_x, acc = torch.empty_like(ctx["x"]), False
        sc = {"d_logits": empty(M, N), "d_hn": empty(M, Dh), "d_h": empty(M, Dh)}
        w, gw = self._structs()
        scs = _struct(_lib.HeadScratch, sc)
        _lib.call("bdetr_head_bwd", M, D, Dh, N, self.kind, ctx["bn_training"], ctx["mult"], ptr(ctx["x"]), ctypes.byref(w), BN_EPS,
                  ctypes.byref(ctx["saved_struct"]), ptr(f32(d_cum)), ptr(d_x), 1 if acc else 0, ctypes.byref(gw),
                  ctypes.byref(scs), stream_ptr())
        return d_x


class BoxPredictionHead(_Head):
    """[batch, num_obj, 4] = 3*sigmoid(x/100) - 1  (reference :13-69)."""
    kind, dense1, dense2 = KIND_BOX, "Dense", "BoxCoords"

    def __init__(self, hidden_dim, num_preds, name="BoxPredictionHead", **kwargs):
        super().__init__(4, hidden_dim, num_preds, name)

    def get_config(self):
        return {**super().get_config(), "hidden_dim": self.hidden_dim, "num_preds": self.num_preds}


class SingleClassPredictionHead(_Head):
    """softmax class probabilities  (reference :72-137)."""
    kind, dense1, dense2 = KIND_SOFTMAX, "DenseCateg", "DenseLogits"

    def __init__(self, num_classes, hidden_dim, num_preds, name="SingleClassPredictionHead", **kwargs):
        super().__init__(num_classes, hidden_dim, num_preds, name)
        self.num_classes = num_classes

    def get_config(self):
        return {**super().get_config(), "num_classes": self.num_classes, "hidden_dim": self.hidden_dim,
                "num_preds": self.num_preds}


class MultiClassPredictionHead(_Head):
    """independent sigmoid per class  (reference :140-207)."""
    kind, dense1, dense2 = KIND_SIGMOID, "Dense", "DenseLinear"

    def __init__(self, num_classes, hidden_dim, num_preds, name="MultiClassPredictionHead", **kwargs):
        super().__init__(num_classes, hidden_dim, num_preds, name)
        self.num_classes = num_classes

    def get_config(self):
        return {**super().get_config(), "num_classes": self.num_classes, "hidden_dim": self.hidden_dim,
                "num_preds": self.num_preds}


# ---------------------------------------------------------------------------------------------------------------------
# Fused tensor-core path: the three heads of one boosted block in one call (bdetr_heads_fwd / bdetr_heads_bwd)
# ---------------------------------------------------------------------------------------------------------------------
def _params3(heads, which):
    arr = (_lib.HeadParams * 3)()
    for k, hd in enumerate(heads):
        src = hd._structs()[which]
        for f, _ in _lib.HeadParams._fields_:
            setattr(arr[k], f, getattr(src, f))
    return arr


def heads_forward_fused(heads, x, training, cum_in, mult):
    """heads = (category, attribute, box) layers; x [B,Q,D]; cum_in = three running predictions or None.
    Returns (new running predictions [3], ctx).  The running sum of THIS block gets fresh buffers (every block's loss
    keeps its own), written by the kernel as cum_in + mult * prediction (boosted_model.py:222-229)."""
    x = f32(x)
    for hd in heads:
        hd.maybe_build([x])
    B, Q, D = x.shape
    M, Dh = B * Q, heads[0].hidden_dim
    C, A = heads[0].num_out, heads[1].num_out
    ntot = C + A + 4
    sv = {"h": empty(3, M, Dh), "bn_mean": empty(3, Dh), "bn_rstd": empty(3, Dh), "w2f": empty(3, Dh), "b2f": empty(3, Dh),
          "bn_part": empty(((M + 127) // 128) * 2 * 3 * Dh), "act0": empty(B, Q, C), "act1": empty(B, Q, A), "act2": empty(B, Q, 4)}
    cum_out = [empty(B, Q, n) for n in (C, A, 4)]
    bn_training = _lib.INT3(*[1 if (training and hd.trainable) else 0 for hd in heads])
    ci, co = _lib.PTR3(), _lib.PTR3()
    for k in range(3):
        ci[k] = None if cum_in is None else cum_in[k].data_ptr()
        co[k] = cum_out[k].data_ptr()
    w3 = _params3(heads, 0)
    svs = _struct(_lib.HeadsSaved, sv)
    _lib.call("bdetr_heads_fwd", M, D, Dh, C, A, ptr(x), w3, ctypes.byref(bn_training), BN_EPS, BN_MOMENTUM, float(mult),
              ctypes.byref(ci), ctypes.byref(co), ctypes.byref(svs), stream_ptr())
    ctx = {"x": x, "saved": sv, "saved_struct": svs, "dims": (M, D, Dh, C, A), "mult": float(mult), "bn_training": bn_training,
           "w3": w3, "cum_in": cum_in, "cum": cum_out}
    return cum_out, ctx


def heads_backward_fused(heads, ctx, d_cum, d_x=None, acc=False, need_dx=True):
    """d_cum = gradients w.r.t. the three running predictions.  Frozen heads contribute no parameter gradients."""
    M, D, Dh, C, A = ctx["dims"]
    ntot = C + A + 4
    if need_dx and d_x is None:
        d_x, acc = torch.empty_like(ctx["x"]), False
    sc = {"d_logits": empty(M * ntot), "hTd": empty(Dh * ntot), "colsum_d": empty(ntot), "bn_s": empty(2 * 3 * Dh), "d_h": empty(3, M, Dh)}
    scs = _struct(_lib.HeadsScratch, sc)
    g3 = _params3(heads, 1)
    gw = _lib.PTR3()
    for k, hd in enumerate(heads):
        gw[k] = ctypes.addressof(g3[k]) if hd.trainable else None
    dc = _lib.PTR3(*[f32(t).data_ptr() for t in d_cum])
    _lib.call("bdetr_heads_bwd", M, D, Dh, C, A, ptr(ctx["x"]), ctx["w3"], ctypes.byref(ctx["bn_training"]), ctx["mult"],
              ctypes.byref(ctx["saved_struct"]), ctypes.byref(dc), ptr(d_x if need_dx else None), 1 if acc else 0,
              ctypes.byref(gw), ctypes.byref(scs), stream_ptr())
    ctx["_keep_bwd"] = (sc, g3)
    return d_x
