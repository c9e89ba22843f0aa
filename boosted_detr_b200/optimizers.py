"""Optimizer and learning-rate schedule of the reference's training runs, on the flat weight / gradient buffers.

Reference: /root/reference/Boosted_DETR_COCO.ipynb cell 26 / 30
    lr = tf.keras.optimizers.schedules.CosineDecayRestarts(initial_learning_rate=.001, first_decay_steps=4000,
                                                           m_mul=.95, alpha=0.1)
    tf.keras.optimizers.SGD(learning_rate=lr, momentum=.9, nesterov=True, clipnorm=0.1)
and the per-block `layer.trainable = True / False` freezing schedule of cell 30 (frozen variables are left out of the
chunk table, so they are neither clipped nor updated).  The update itself is bdetr_sgd_step (csrc/optimizer.cu).
"""
from __future__ import annotations

import ctypes
import os
import math

import numpy as np
import torch

from . import _lib
from .device import ptr, stream_ptr, zeros

CHUNK = int(os.environ.get("BDETR_OPT_CHUNK", "2048"))    # floats per CTA of the update kernels: a 256 x 256 kernel is 32 CTAs (16 K-element
                    # chunks left a 0.5 M-float range on ~50 CTAs: 24 us alone, 40-70 us beside the backward, all on the bucket pipeline's critical tail)


class CosineDecayRestarts:
    """tf.keras.optimizers.schedules.CosineDecayRestarts (SGDR), evaluated in float32 like TensorFlow does."""

    def __init__(self, initial_learning_rate, first_decay_steps, t_mul=2.0, m_mul=1.0, alpha=0.0, name=None):
        self.initial_learning_rate = float(initial_learning_rate)
        self.first_decay_steps = int(first_decay_steps)
        self.t_mul, self.m_mul, self.alpha = float(t_mul), float(m_mul), float(alpha)
        self.name = name

    def __call__(self, step: int) -> float:
        f = np.float32
        completed = f(step) / f(self.first_decay_steps)
        t_mul, m_mul, alpha = f(self.t_mul), f(self.m_mul), f(self.alpha)
        if self.t_mul == 1.0:
            i_restart = np.floor(completed)
            completed = completed - i_restart
        else:
            i_restart = np.floor(np.log(f(1.0) - completed * (f(1.0) - t_mul)) / np.log(t_mul))
            sum_r = (f(1.0) - t_mul ** i_restart) / (f(1.0) - t_mul)
            completed = (completed - sum_r) / t_mul ** i_restart
        m_fac = m_mul ** i_restart
        cosine_decayed = f(0.5) * m_fac * (f(1.0) + np.cos(f(math.pi) * completed))
        decayed = (f(1.0) - alpha) * cosine_decayed + alpha
        return float(f(self.initial_learning_rate) * decayed)

    def get_config(self):
        return {"initial_learning_rate": self.initial_learning_rate, "first_decay_steps": self.first_decay_steps,
                "t_mul": self.t_mul, "m_mul": self.m_mul, "alpha": self.alpha, "name": self.name}


class SGD:
    """tf.keras.optimizers.SGD(learning_rate, momentum, nesterov, clipnorm) for BoostedDETR.compile()."""

    def __init__(self, learning_rate=0.01, momentum=0.0, nesterov=False, clipnorm=None, name="SGD", **kwargs):
        if kwargs:
            raise TypeError(f"unsupported SGD arguments: {sorted(kwargs)}")
        self.learning_rate = learning_rate
        self.momentum = float(momentum)
        self.nesterov = bool(nesterov)
        self.clipnorm = None if clipnorm is None else float(clipnorm)
        self.name = name
        self.iterations = 0
        self._table_key = None
        self._table = None
        self._accum = None

    def current_lr(self) -> float:
        lr = self.learning_rate
        return float(lr(self.iterations)) if callable(lr) else float(lr)

    # -- chunk table over the trainable variables ------------------------------------------------
    @staticmethod
    def trainable_slots(model):
        """(name, offset, count) of every variable the optimizer updates, in flat-buffer order."""
        slots = []
        for name, owner, key in model.named_weights():
            if key in owner._non_trainable or not owner.trainable or name not in model._index:
                continue
            off, cnt, _ = model._index[name]
            slots.append((name, off, cnt))
        slots.sort(key=lambda s: s[1])
        return slots

    @staticmethod
    def chunk_table(slots) -> np.ndarray:
        rows = []
        for _, off, cnt in slots:
            n = max(1, -(-cnt // CHUNK))
            first = len(rows)
            for c in range(n):
                lo = c * CHUNK
                rows.append((off + lo, min(CHUNK, cnt - lo), first, n, 0))
        dt = np.dtype([("offset", "<i8"), ("len", "<i4"), ("var_first", "<i4"), ("var_chunks", "<i4"), ("reserved", "<i4")])
        return np.array(rows, dtype=dt) if rows else np.zeros(0, dtype=dt)

    def _ensure_table(self, model):
        from .layers import Layer
        key = (id(model._flat[0]), Layer.trainable_epoch)
        if key != self._table_key:
            tab = self.chunk_table(self.trainable_slots(model))
            dev = model._flat[0].device
            self._table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
            self._n_chunks = len(tab)
            self._partial = zeros(max(1, len(tab)))
            self._table_key = key
        if self._accum is None or self._accum.numel() != model._flat[0].numel():
            self._accum = zeros(model._flat[0].numel())       # Keras slot "momentum", one per variable, flat like the weights

    def _launch(self, model, lr, lr_dev):
        _lib.call("bdetr_sgd_step", self._n_chunks, ptr(self._table), ptr(model._flat[0]), ptr(model._flat[1]),
                  ptr(self._accum), ptr(self._partial), lr, lr_dev, self.momentum, int(self.nesterov),
                  0.0 if self.clipnorm is None else self.clipnorm, stream_ptr())

    # -- per-bucket form: the update of boosted block i's variables right behind its gradient all-reduce -------------
    def _ensure_bucket_tables(self, model):
        """One chunk table per gradient range (model._ranges: contiguous ranges of the flat buffers -- the decoder / heads
        part and the encoder part of every boosted block, in backward order), so the clip + update of a range can be queued as soon as its gradients are final
        (the optimizer step fused behind the all-reduce epilogue) instead of after the whole backward."""
        from .layers import Layer
        self._ensure_table(model)
        key = (id(model._flat[0]), Layer.trainable_epoch)
        if getattr(self, "_btab_key", None) == key:
            return
        slots = self.trainable_slots(model)
        tabs, counts = [], []
        ranges = list(getattr(model, "_ranges", None) or [(lo, hi) for _, lo, hi in model._buckets])
        self._brange = {r: k for k, r in enumerate(ranges)}      # (first float, one-past-last float) -> table index
        for lo, hi in ranges:
            t = self.chunk_table([sl for sl in slots if lo <= sl[1] < hi])
            tabs.append(t)
            counts.append(len(t))
        dev = model._flat[0].device
        allt = np.concatenate(tabs) if sum(counts) else np.zeros(0, dtype=tabs[0].dtype)
        self._btable = torch.from_numpy(allt.view(np.uint8).copy()).to(dev) if len(allt) else torch.zeros(32, dtype=torch.uint8, device=dev)
        self._bpartial = zeros(max(1, len(allt)))
        self._bcounts = counts
        self._bfirst = [int(x) for x in np.concatenate([[0], np.cumsum(counts)[:-1]])]
        self._btab_key = key

    def launch_bucket(self, model, lo_hi, lr=0.0, lr_dev=None):
        """Clip + update of the variables in the gradient range `lo_hi` = (first float, one-past-last float), one of
        model._ranges, on the current stream."""
        bucket_index = self._brange[tuple(lo_hi)]
        n, first = self._bcounts[bucket_index], self._bfirst[bucket_index]
        if n == 0:
            return
        rec = 24                                                  # sizeof(bdetr_opt_chunk)
        tab = ctypes.c_void_p(self._btable.data_ptr() + first * rec)
        part = ctypes.c_void_p(self._bpartial.data_ptr() + first * 4)
        _lib.call("bdetr_sgd_step", n, tab, ptr(model._flat[0]), ptr(model._flat[1]), ptr(self._accum), part, lr, lr_dev,
                  self.momentum, int(self.nesterov), 0.0 if self.clipnorm is None else self.clipnorm, stream_ptr())

    def finish_step(self, lr):
        """Book-keeping of a step whose update ran bucket by bucket."""
        self.iterations += 1
        self.last_lr = lr

    def apply(self, model):
        """One optimizer step on model's flat buffers with the gradients currently in them (after any all-reduce)."""
        if model._flat is None:
            raise RuntimeError("build the model before applying the optimizer")
        self._ensure_table(model)
        lr = self.current_lr()
        self._launch(model, lr, None)
        self.iterations += 1
        self.last_lr = lr

    # -- CUDA-graph form: the update kernels are captured once and read the step's rate from device memory -------
    def prepare(self, model):
        """Allocations and host->device copies of the chunk table: must happen OUTSIDE graph capture."""
        self._ensure_table(model)
        self._ensure_bucket_tables(model)
        if getattr(self, "_lr_dev", None) is None:
            self._lr_dev = zeros(1)
            self._lr_host = torch.zeros(16, dtype=torch.float32).pin_memory()    # ring: the host may run a few steps ahead

    def capture(self, model):
        """Call inside graph capture (after the gradients are final; `prepare` first).  The table is frozen into the
        graph: changing `trainable` flags afterwards needs a new capture."""
        self._launch(model, 0.0, ptr(self._lr_dev))

    def pre_replay(self):
        """Before every replay of a graph that contains `capture`: evaluates the schedule for this step and queues the
        4-byte pinned H2D copy of the rate on the current stream."""
        lr = self.current_lr()
        slot = self.iterations % 16
        self._lr_host[slot] = lr
        self._lr_dev.copy_(self._lr_host[slot:slot + 1], non_blocking=True)
        self.iterations += 1
        self.last_lr = lr

    def get_config(self):
        lr = self.learning_rate
        return {"name": self.name, "learning_rate": lr.get_config() if hasattr(lr, "get_config") else float(lr),
                "momentum": self.momentum, "nesterov": self.nesterov, "clipnorm": self.clipnorm}
