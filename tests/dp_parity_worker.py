"""Worker of tests/test_parallel_gpu.py (launched by torchrun, one process per GPU; not a pytest file).

SURVEY.md §8e parity check on hardware: the G-GPU data-parallel step == a single-GPU emulation of G replicas (same shards,
per-replica BatchNorm statistics and loss normaliser, loss scaled by 1/G, gradients summed).  Every rank runs the emulation
of ALL shards on its own GPU and compares it with what the NCCL path left in its flat gradient buffer, for
DataParallel(overlap=True / False), eager and under the CUDA graph, in tensor-core mode (the benchmarked mode)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dense_gpu import _model_and_data                     # noqa: E402
from boosted_detr_b200 import _lib                              # noqa: E402
from boosted_detr_b200.graph import GraphedTrainStep            # noqa: E402
from boosted_detr_b200.parallel import DataParallel, init_from_env, shard_batch   # noqa: E402


def nerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    _lib.load().bdetr_set_mode(_lib.MODE_TF32)
    N, B = 3, 4 * world
    model, w0, inputs = _model_and_data(N=N, B=B, rows=10, cols=10, Q=50, T=10)       # same seed on every rank
    model.dropout_seed = None
    w0 = {k: v.copy() for k, v in w0.items()}
    # ---- emulation on this GPU: every shard through the un-hooked model, loss scaled by 1/world, gradients summed
    model.num_replicas = world
    emu = None
    losses = []
    for r in range(world):
        model.set_weights_dict(w0)
        model.train_step(shard_batch(inputs, r, world))
        torch.cuda.synchronize()
        g = model._flat[1].clone()
        emu = g if emu is None else emu + g
        losses.append(model.metric_tensors["loss"].clone())
    emu = emu.cpu().numpy()
    mine = shard_batch(inputs, rank, world)
    worst = 0.0
    for overlap in (True, False):
        model.set_weights_dict(w0)
        model.grad_allreduce = model.grad_bucket_hook = None
        DataParallel(model, overlap=overlap)
        model.train_step(mine)
        torch.cuda.synchronize()
        e = nerr(model._flat[1].cpu().numpy(), emu)
        assert torch.equal(model.metric_tensors["loss"], losses[rank]), "per-replica loss must not depend on the collective"
        print(f"rank {rank}: eager overlap={overlap}: all-reduced gradient vs {world}-replica emulation {e:.2e}", flush=True)
        worst = max(worst, e)
        gs = GraphedTrainStep(model, mine)
        model.set_weights_dict(w0)
        gs.load(mine)
        gs.replay()
        torch.cuda.synchronize()
        e = nerr(model._flat[1].cpu().numpy(), emu)
        print(f"rank {rank}: graph overlap={overlap}: {e:.2e}", flush=True)
        worst = max(worst, e)
        del gs
    # fp32 reduction-order noise only (atomics inside the wgrad kernels + the ring order of the all-reduce)
    assert worst < 1e-4, worst
    print(f"rank {rank}: DP PARITY OK (worst {worst:.2e})", flush=True)
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
