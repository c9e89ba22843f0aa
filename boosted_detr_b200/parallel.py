"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the gradient
all-reduce — the only collective on the path.  Per-replica BatchNorm statistics and loss
normaliser, loss scaled by 1/replicas, gradients summed: the semantics of the reference's
tf.distribute.MirroredStrategy (parameters.py:74)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # The gradient buckets are ~5 MB and hidden under the backward, so ring / tree over NVLink P2P is all the
            # bandwidth this path needs.  NVLS (in-switch reduction) is left off by default: its buffer registration
            # inside CUDA-graph capture could not be validated on 4 / 8 GPUs in round 1 (export NCCL_NVLS_ENABLE=1 to
            # try it).
            os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        import datetime
        dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
    return rank, world, local


def shard_batch(inputs: dict, rank: int, world: int) -> dict:
    """Splits the global batch evenly by image (partitioning of SURVEY.md §8e)."""
    out = {}
    for k, v in inputs.items():
        n = v.shape[0]
        assert n % world == 0, "global batch must divide evenly across replicas"
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


class Comm:
    """bdetr_comm: the NCCL communicator behind the C ABI (csrc/comm.cu).  torch.distributed is only the out-of-band
    channel that hands rank 0's 128-byte NCCL unique id to the other ranks; the gradient bytes never touch it."""

    def __init__(self):
        import ctypes
        from . import _lib
        assert dist.is_initialized()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        idbuf = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            _lib.call("bdetr_comm_unique_id", idbuf)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        idbuf = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
        self.handle = ctypes.c_void_p()
        _lib.call("bdetr_comm_init", ctypes.byref(self.handle), self.rank, self.world, idbuf)
        ver = ctypes.c_int(0)
        _lib.call("bdetr_comm_info", self.handle, None, None, ctypes.byref(ver))
        self.nccl_version = ver.value

    def allreduce(self, t: torch.Tensor):
        from . import _lib
        from .device import ptr, stream_ptr
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        _lib.call("bdetr_allreduce", self.handle, ptr(t), t.numel(), stream_ptr())

    def broadcast(self, t: torch.Tensor, root=0):
        from . import _lib
        from .device import ptr, stream_ptr
        _lib.call("bdetr_broadcast", self.handle, ptr(t), t.numel(), root, stream_ptr())


_comm = None


def get_comm():
    """The process's bdetr_comm (created on first use; needs torch.distributed initialised and a CUDA device)."""
    global _comm
    if _comm is None and dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available() \
            and os.environ.get("BDETR_COMM", "abi") == "abi":
        _comm = Comm()
    return _comm


def allreduce_gradients(flat_grads: torch.Tensor):
    """Sum of the per-replica gradients (each already scaled by 1/replicas): the only collective on the path.  CUDA
    buffers go through bdetr_allreduce (ncclAllReduce behind the C ABI); host tensors (the gloo tests) through
    torch.distributed."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        comm = get_comm() if flat_grads.is_cuda else None
        if comm is not None:
            comm.allreduce(flat_grads)
        else:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def broadcast_tensor(t: torch.Tensor, src: int = 0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        comm = get_comm() if t.is_cuda else None
        if comm is not None and t.dtype == torch.float32 and t.is_contiguous():
            comm.broadcast(t, src)
        else:
            dist.broadcast(t, src=src)
    return t


class DataParallel:
    """Attaches the gradient all-reduce to a BoostedDETR: model.train_step then computes
    sum over replicas of d(loss_replica / world)/d(theta).

    overlap=True (default): the flat gradient buffer is laid out block by block in backward order
    (BoostedDETR._flatten), and as soon as the backward has finished boosted block i its bucket (~5 MB) is
    all-reduced on a communication stream underneath the backward of blocks i-1 .. 0; only the last bucket is
    exposed (BDETR_SPLIT_RANGES=1 sends the decoder / heads part of a bucket ahead of its encoder part: no gain measured).
    The clip + SGD update of a bucket follows its all-reduce on a third stream, so the next all-reduce never waits for it.
    The collectives are issued in the same order on every rank and are captured into the CUDA graph of
    the step together with the kernels.  overlap=False: one all-reduce of the whole buffer after the backward."""

    def __init__(self, model, overlap=True, bucket_optimizer=True):
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.comm_stream = None
        self.opt_lr = None                # (lr, lr_dev) of the running step when the update runs bucket by bucket
        model.num_replicas = self.world
        if self.world > 1:
            model.grad_allreduce = self.finish
            if self.overlap:
                model.grad_bucket_hook = self.reduce_bucket
        elif bucket_optimizer:
            # one GPU: nothing to reduce, but the per-bucket optimizer update still leaves the critical path
            model.grad_allreduce = self.finish
            model.grad_bucket_hook = self.reduce_bucket
            self.overlap = True
        self.bucket_optimizer = bucket_optimizer and self.overlap
        model.bucket_pipeline = self
        self.broadcast_weights()

    def _comm(self):
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        return self.comm_stream

    def _opt(self):
        if getattr(self, "opt_stream", None) is None:
            self.opt_stream = torch.cuda.Stream()
        return self.opt_stream

    def reduce_bucket(self, block: int, lo: int, hi: int, events=(), join_from=None):
        """All-reduce flat_grads[lo:hi] (boosted block `block`) on the communication stream, ordered after the
        caller's stream and `events`."""
        comm = self._comm()
        comm.wait_stream(torch.cuda.current_stream())
        for ev in events:
            comm.wait_event(ev)
        if join_from is not None:          # deferred parameter-gradient chains forked from the caller's stream (bdetr_join_into)
            import ctypes
            from . import _lib
            _lib.call("bdetr_join_into", ctypes.c_void_p(join_from.cuda_stream), ctypes.c_void_p(comm.cuda_stream))
        with torch.cuda.stream(comm):
            self.model._mark(f"block {block} range [{lo}:{hi}) final (comm)")
            if self.world > 1:
                allreduce_gradients(self.model._flat[1][lo:hi])
                self.model._mark(f"block {block} range all-reduced (comm)")
        # clip + SGD update of this range's variables right behind its all-reduce -- on a stream of its
        # own, so the NEXT range's all-reduce does not queue behind this update (the communication stream was a serial
        # chain of all-reduce + update pairs that outlasted the backward: tools/trace_step.py under torchrun)
        if self.opt_lr is not None:
            m = self.model
            opt = self._opt() if self.world > 1 else comm
            opt.wait_stream(comm)
            with torch.cuda.stream(opt):
                m.optimizer.launch_bucket(m, (lo, hi), self.opt_lr[0], self.opt_lr[1])
                m._mark(f"block {block} range updated (opt)")

    def finish(self, flat_grads: torch.Tensor):
        """Called at the end of the backward: joins the bucketed all-reduces, or does the single one."""
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self._comm())
            if getattr(self, "opt_stream", None) is not None:
                torch.cuda.current_stream().wait_stream(self.opt_stream)
        elif self.world > 1:
            allreduce_gradients(flat_grads)

    # kept for callers that drive the collective themselves
    def allreduce(self, flat_grads: torch.Tensor):
        allreduce_gradients(flat_grads)

    def broadcast_weights(self):
        if self.world > 1:
            if self.model._flat is None:
                self.model.build()
            broadcast_tensor(self.model._flat[0])
            for _, o, k in self.model.named_weights():
                if k in o._non_trainable:
                    broadcast_tensor(o._weights[k])
