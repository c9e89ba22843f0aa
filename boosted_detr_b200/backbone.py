"""BackboneNeck — the step immediately before the hot path: same class name / constructor / call
convention as the reference's ModelComponents/backbone.py:66-95.  BatchNorm -> 1x1 Conv2D(encoder_dim, lecun_normal,
tanh) -> BatchNorm over the channels-last feature map of the EfficientNet backbone (which itself stays out of scope and,
in the reference's training runs, frozen: notebook cell 30).  Runs as one tcgen05 GEMM with both normalisations folded
around it (bdetr_backbone_neck_fwd / _bwd); tensor-core mode only."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .device import empty, f32, ptr, stream_ptr
from .layers import Layer, truncated_normal
from .prediction_heads import BN_EPS, BN_MOMENTUM
from .transformers import _struct


class BackboneNeck(Layer):
    def __init__(self, encoder_dim, name="BackboneNeck", **kwargs):
        super().__init__(name=name)
        self.encoder_dim = encoder_dim

    def get_config(self):
        return {**super().get_config(), "encoder_dim": self.encoder_dim}

    def build(self, input_shape):
        cin, n = input_shape[0][-1], self.encoder_dim
        self.features_shape = input_shape[0]
        rng = Layer._rng
        self.add_weight("batch_norm1/gamma", np.ones(cin, np.float32))
        self.add_weight("batch_norm1/beta", np.zeros(cin, np.float32))
        self.add_weight("batch_norm1/moving_mean", np.zeros(cin, np.float32), trainable=False)
        self.add_weight("batch_norm1/moving_variance", np.ones(cin, np.float32), trainable=False)
        # lecun_normal: truncated normal, stddev sqrt(1 / fan_in); Conv2D kernel [1, 1, Cin, N]
        self.add_weight("conv2d_downscaler/kernel", truncated_normal(rng, (1, 1, cin, n), np.sqrt(1.0 / cin) / 0.87962566103423978))
        self.add_weight("conv2d_downscaler/bias", np.zeros(n, np.float32))
        self.add_weight("batch_norm2/gamma", np.ones(n, np.float32))
        self.add_weight("batch_norm2/beta", np.zeros(n, np.float32))
        self.add_weight("batch_norm2/moving_mean", np.zeros(n, np.float32), trainable=False)
        self.add_weight("batch_norm2/moving_variance", np.ones(n, np.float32), trainable=False)

    def _structs(self):
        if self._struct_cache is None:
            names = {"bn1_gamma": "batch_norm1/gamma", "bn1_beta": "batch_norm1/beta", "bn1_moving_mean": "batch_norm1/moving_mean",
                     "bn1_moving_var": "batch_norm1/moving_variance", "conv_w": "conv2d_downscaler/kernel", "conv_b": "conv2d_downscaler/bias",
                     "bn2_gamma": "batch_norm2/gamma", "bn2_beta": "batch_norm2/beta", "bn2_moving_mean": "batch_norm2/moving_mean",
                     "bn2_moving_var": "batch_norm2/moving_variance"}
            w = _struct(_lib.NeckParams, {k: self._weights[v] for k, v in names.items()})
            g = _struct(_lib.NeckParams, {k: self._grads.get(v) for k, v in names.items()})
            self._struct_cache = (w, g)
        return self._struct_cache

    def forward(self, inputs, training=False, round_out=True):
        """inputs = [features [B, rows, cols, Cin]] -> [B, rows, cols, encoder_dim] (tf32-rounded: it is block 0's operand)."""
        x = f32(inputs[0])
        self.maybe_build([x])
        B, R, Cc, cin = x.shape
        M, N = B * R * Cc, self.encoder_dim
        x_tc = torch.empty_like(x)
        _lib.call("bdetr_round_tf32", x.numel(), ptr(x), ptr(x_tc), stream_ptr())
        chunks, parts = (M + 127) // 128, (cin + 31) // 32
        sv = {"t": empty(M, N), "wf": empty(cin, N), "bf": empty(N), "part": empty(max(chunks * 2 * cin, parts * N)),
              "stat1": empty(4, cin), "stat2": empty(4, N)}
        out = empty(B, R, Cc, N)
        bn_training = 1 if (training and self.trainable) else 0          # frozen -> inference-mode BatchNorm (Keras)
        w, _ = self._structs()
        svs = _struct(_lib.NeckSaved, sv)
        _lib.call("bdetr_backbone_neck_fwd", M, cin, N, ptr(x_tc), ctypes.byref(w), BN_EPS, BN_MOMENTUM, bn_training, ptr(out),
                  ctypes.byref(svs), 1 if round_out else 0, stream_ptr())
        return out, {"x_tc": x_tc, "saved": sv, "saved_struct": svs, "dims": (M, cin, N), "bn_training": bn_training}

    def backward(self, ctx, d_out):
        """Parameter gradients (accumulated); no input gradient: the backbone in front is frozen in the reference."""
        if not self.trainable:
            return None
        M, cin, N = ctx["dims"]
        d_u, gwf = empty(M, N), empty(cin * N + N)
        w, g = self._structs()
        _lib.call("bdetr_backbone_neck_bwd", M, cin, N, ptr(ctx["x_tc"]), ctypes.byref(w), ctx["bn_training"], ctypes.byref(ctx["saved_struct"]),
                  ptr(f32(d_out)), ctypes.byref(g), ptr(d_u), ptr(gwf), stream_ptr())
        ctx["_keep_bwd"] = (d_u, gwf)
        return None
