"""CUDA-graph replay of the training step: the whole fwd + matcher + bwd chain is ~600 small kernels, so at
BASELINE config 2 the step is launch-bound; one graph launch replaces the Python/ctypes launch sequence."""
from __future__ import annotations

import numpy as np
import torch

from .device import ptr
from .losses_and_metrics import raise_for_status


class GraphedTrainStep:
    """Captures model.forward + backward for fixed shapes.  __call__(inputs) copies the batch into the static
    input buffers (H2D from pinned staging when given numpy), replays, and returns the metric tensors."""

    def __init__(self, model, example_inputs, warmup=3):
        self.model = model
        if model._flat is None:
            model.build()
        self.static = {}
        feats, y_true = model._prepare(example_inputs, True)
        self.static = {"features": feats.clone(), "category": y_true[0].clone(), "attribute": y_true[1].clone(),
                       "bbox": y_true[2].clone(), "num_objects": y_true[3].clone()}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.static.values())
        if model.optimizer is not None:
            model.optimizer.prepare(model)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if getattr(model, "_trace", None) is not None:
            model._trace.clear()                   # keep only the marks of the captured step
        with torch.cuda.graph(self.graph):
            model._mark("step start (main)")
            self.metrics, self.status = self._step(capturing=True)
            model._mark("step end (main)")
        torch.cuda.synchronize()

    def _step(self, capturing=False):
        m = self.model
        s = self.static
        m.zero_grads()
        y_true = [s["category"], s["attribute"], s["bbox"], s["num_objects"]]
        _, ctx = m.forward(s["features"], y_true, True)
        metrics = m._collect_metrics(ctx)
        pipe = getattr(m, "bucket_pipeline", None)
        # the optimizer update belongs to the graph whenever the gradients are final inside it (single GPU, or the
        # bucketed all-reduce); its learning rate is read from device memory (see SGD.capture / pre_replay).  With a
        # bucket pipeline the update of block i's variables is queued right behind that block's all-reduce.
        self.optimizer_in_graph = m.optimizer is not None and (m.grad_allreduce is None or m.grad_bucket_hook is not None)
        self.optimizer_in_hook = bool(self.optimizer_in_graph and pipe is not None and pipe.bucket_optimizer and ctx.get("fused"))
        if self.optimizer_in_hook:
            pipe.opt_lr = (0.0, ptr(m.optimizer._lr_dev)) if capturing else None
        try:
            m.backward(ctx, gscale=1.0 / m.num_replicas)
        finally:
            if pipe is not None:
                pipe.opt_lr = None
        m._join_metrics()
        if m.grad_bucket_hook is not None and m.grad_allreduce is not None:
            m.grad_allreduce(m._flat[1])          # joins the bucketed all-reduces issued inside the backward (captured too)
        if self.optimizer_in_graph and capturing and not self.optimizer_in_hook:
            m.optimizer.capture(m)
        return metrics, m.status_all

    def load(self, inputs):
        """Host batch -> (pinned staging) -> static device buffers, async on the current stream."""
        m = self.model
        m.h2d_bytes = 0
        for k in ("features", "category", "attribute", "bbox", "num_objects"):
            x, dst = inputs[k], self.static[k]
            dtype = "i32" if k == "num_objects" else "f32"
            if isinstance(x, torch.Tensor) and x.is_cuda:
                if x.data_ptr() != dst.data_ptr():
                    dst.copy_(x.reshape(dst.shape), non_blocking=True)
            elif isinstance(x, torch.Tensor) and x.is_pinned() and x.dtype == dst.dtype and x.is_contiguous():
                dst.copy_(x.reshape(dst.shape), non_blocking=True)       # already in pinned host memory: no staging copy
                m.h2d_bytes += x.numel() * x.element_size()
            else:
                m._h2d(k, x, dtype, dst=dst)

    def _pre(self):
        """The per-step device words the captured kernels read (learning rate, dropout seed), queued on the current stream."""
        m = self.model
        if self.optimizer_in_graph:
            m.optimizer.pre_replay()
        m.push_dropout_seed()

    def replay(self):
        m = self.model
        if not getattr(self, "_primed", False):   # (already queued behind the previous launch by __call__(prefetch=...))
            self._pre()
        self._primed = False
        self.graph.replay()
        if m.grad_allreduce is not None and m.grad_bucket_hook is None:
            m.grad_allreduce(m._flat[1])          # non-overlapped mode: one all-reduce after the graph
        if m.optimizer is not None and not self.optimizer_in_graph:
            m.optimizer.apply(m)
        m.step_count += 1
        m.advance_dropout_seed()

    # -- input prefetch ---------------------------------------------------------------------------------------------
    def prefetch(self, inputs):
        """Starts the H2D copy of the NEXT batch into a second set of device buffers on a copy stream, so that it runs
        underneath the step that is executing; `__call__(inputs)` with the same object then only does device-to-device
        copies.  The caller keeps `inputs` (pinned host tensors) alive and unmodified until that call."""
        if not hasattr(self, "_staging"):
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged_ev, self._staging_free = torch.cuda.Event(), torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream())
            self._staged_for = None
        cs = self._copy_stream
        cs.wait_event(self._staging_free)                 # the previous staging -> static copy has run
        with torch.cuda.stream(cs):
            for k, dst in self._staging.items():
                x = inputs[k]
                if not (isinstance(x, torch.Tensor) and (x.is_cuda or x.is_pinned())):
                    self.model._h2d(k, x, "i32" if k == "num_objects" else "f32", dst=dst)
                else:
                    dst.copy_(x.reshape(dst.shape), non_blocking=True)
            self._staged_ev.record(cs)
        self._staged_for = inputs

    def _take_prefetched(self):
        main = torch.cuda.current_stream()
        main.wait_event(self._staged_ev)
        for k, dst in self.static.items():
            dst.copy_(self._staging[k], non_blocking=True)
        self._staging_free.record(main)
        self._staged_for = None

    def __call__(self, inputs, return_host=True, prefetch=None):
        """One training step on `inputs`.  `prefetch` = the batch of the FOLLOWING step (optional): its H2D copy runs on a
        copy stream underneath this step, and everything else the next launch needs -- the staging -> static input copies
        and the next step's learning-rate / dropout-seed words -- is queued on the main stream right BEHIND this step's
        graph while the host would otherwise sit in the logs' synchronisation.  The next call then only launches the graph
        (the host-side issue of ~8 small copies, ~50 us, no longer sits between two steps)."""
        if getattr(self, "_primed_for", None) is inputs and inputs is not None:
            pass                                  # inputs and per-step words are already in place
        elif getattr(self, "_staged_for", None) is inputs and inputs is not None:
            self._take_prefetched()
        else:
            self.load(inputs)
        self._primed_for = None
        self.replay()
        if prefetch is not None:
            self.prefetch(prefetch)
            if self.optimizer_in_graph or self.model.optimizer is None:
                # stream order keeps this behind the running graph, which still reads the current batch and words
                self._take_prefetched()
                self._pre()
                self._primed, self._primed_for = True, prefetch
        if not return_host:
            return self.metrics
        return self.model.host_logs()
