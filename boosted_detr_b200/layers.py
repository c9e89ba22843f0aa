"""Minimal Keras-shaped layer protocol for the host side (the reference's boundary is the Keras
layer protocol: __init__(**hparams), build(input_shape), call(list_of_tensors, training), get_config()).

Differences a reference user will notice: tensors are CUDA device buffers, and because there is no
autodiff tape every layer also has `forward(inputs, training) -> (outputs, ctx)` and
`backward(ctx, grads)`; `call` is `forward` with the context kept in `self.last_ctx`.
"""
from __future__ import annotations

import numpy as np
import torch

from .device import require_cuda


def lowbias32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def dropout_site_key(site: int) -> int:
    """Seed-independent half of the key: what the kernels combine with the device-resident step seed."""
    return lowbias32((site + 0x9E3779B9) & 0xFFFFFFFF)


def dropout_key(seed: int, site: int) -> int:
    """Key of the counter-based dropout mask of one Dropout layer instance (DESIGN.md "dropout")."""
    return lowbias32((seed & 0xFFFFFFFF) ^ dropout_site_key(site))


def truncated_normal(rng: np.random.Generator, shape, stddev: float) -> np.ndarray:
    """Keras TruncatedNormal semantics (resample beyond 2 sigma)."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * stddev).astype(np.float32)


def glorot_normal(rng, fan_in, fan_out):
    return truncated_normal(rng, (fan_in, fan_out), np.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978)


def he_normal(rng, fan_in, fan_out):
    return truncated_normal(rng, (fan_in, fan_out), np.sqrt(2.0 / fan_in) / 0.87962566103423978)


class Layer:
    _rng = np.random.default_rng(0)
    trainable_epoch = 0            # bumped whenever any layer's `trainable` flag is written (optimizer table cache key)

    def __init__(self, name=None, **kwargs):
        self.name = name or type(self).__name__
        self.built = False
        self._trainable = True
        self._weights: dict[str, torch.Tensor] = {}
        self._grads: dict[str, torch.Tensor] = {}
        self._non_trainable: set[str] = set()
        self._shadow: dict[str, torch.Tensor] = {}     # tf32-rounded copies of the Dense kernels (tensor-core mode)
        self._struct_cache = None
        self.last_ctx = None

    # Keras semantics: setting `trainable` on a layer sets it on all of its sublayers (the reference freezes whole
    # boosted blocks this way, Boosted_DETR_COCO.ipynb cell 30); the optimizer skips variables of frozen layers.
    @property
    def trainable(self):
        return self._trainable

    @trainable.setter
    def trainable(self, value):
        self._trainable = bool(value)
        Layer.trainable_epoch += 1
        for sub in self.sublayers():
            sub.trainable = value

    # -- weights ---------------------------------------------------------------------------
    def add_weight(self, name: str, value: np.ndarray, trainable: bool = True) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)).to(require_cuda())
        self._weights[name] = t
        if trainable:
            self._grads[name] = torch.zeros_like(t)
        else:
            self._non_trainable.add(name)
        return t

    def sublayers(self):
        for v in self.__dict__.values():
            if isinstance(v, Layer):
                yield v
            elif isinstance(v, (list, tuple)):
                for x in v:
                    if isinstance(x, Layer):
                        yield x

    def named_weights(self, prefix=""):
        """Yields (keras-style full name, owner layer, local key) for this layer and its children."""
        base = f"{prefix}{self.name}/"
        for k in self._weights:
            yield base + k, self, k
        for sub in self.sublayers():
            yield from sub.named_weights(base)

    def gemm_weights(self):
        """Weights as the GEMMs should read them: tf32-rounded shadows of the Dense kernels in tensor-core mode."""
        from . import _lib
        if self._shadow and _lib.tc_mode():
            return {**self._weights, **self._shadow}, 1
        return self._weights, 0

    def invalidate(self):
        self._struct_cache = None
        for sub in self.sublayers():
            sub.invalidate()

    # -- keras protocol ----------------------------------------------------------------------
    def build(self, input_shape):
        pass

    def maybe_build(self, inputs):
        if not self.built:
            self.build([tuple(t.shape) for t in inputs])
            self.built = True

    def get_config(self):
        return {"name": self.name}

    def call(self, inputs, training=False, **kwargs):
        out, self.last_ctx = self.forward(inputs, training=training, **kwargs)
        return out

    def __call__(self, inputs, training=False, **kwargs):
        return self.call(inputs, training=training, **kwargs)

    def show_summary(self):
        rows = [(n, tuple(o._weights[k].shape)) for n, o, k in self.named_weights()]
        for n, s in rows:
            print(f"{n:100s} {s}")
        print("total parameters:", sum(int(np.prod(s)) for _, s in rows))
