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


def allreduce_gradients(flat_grads: torch.Tensor):
    """Sum of the per-replica gradients (each already scaled by 1/replicas): the only collective on the path."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def broadcast_tensor(t: torch.Tensor, src: int = 0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


class DataParallel:
    """Attaches the gradient all-reduce to a BoostedDETR: model.train_step then computes
    sum over replicas of d(loss_replica / world)/d(theta).

    overlap=True (default): the flat gradient buffer is laid out block by block in backward order
    (BoostedDETR._flatten), and as soon as the backward has finished boosted block i its bucket (~5 MB) is
    all-reduced on a communication stream underneath the backward of blocks i-1 .. 0; only the last bucket is
    exposed.  The collectives are issued in the same order on every rank and are captured into the CUDA graph of
    the step together with the kernels.  overlap=False: one all-reduce of the whole buffer after the backward."""

    def __init__(self, model, overlap=True):
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.comm_stream = None
        model.num_replicas = self.world
        if self.world > 1:
            model.grad_allreduce = self.finish
            if self.overlap:
                model.grad_bucket_hook = self.reduce_bucket
        self.broadcast_weights()

    def _comm(self):
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        return self.comm_stream

    def reduce_bucket(self, block: int, lo: int, hi: int, events=()):
        """All-reduce flat_grads[lo:hi] (boosted block `block`) on the communication stream, ordered after the
        caller's stream and `events`."""
        comm = self._comm()
        comm.wait_stream(torch.cuda.current_stream())
        for ev in events:
            comm.wait_event(ev)
        with torch.cuda.stream(comm):
            allreduce_gradients(self.model._flat[1][lo:hi])

    def finish(self, flat_grads: torch.Tensor):
        """Called at the end of the backward: joins the bucketed all-reduces, or does the single one."""
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self._comm())
        else:
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
