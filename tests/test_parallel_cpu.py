"""world_size-2 and -4 gloo tests of the data-parallel host logic: shard by image, per-replica
BatchNorm / loss normaliser, loss scaled by 1/replicas, gradients summed by all-reduce.  The oracle is the
checker: the all-reduced per-replica gradients must equal a single-process emulation of the replicas."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from util import make_weights, synth_targets  # noqa: E402

CFG = dict(N=2, rows=3, cols=3, D=128, H=4, Q=8, T=4, C=6, A=3, B=4)


def _data():
    c = CFG
    rng = np.random.default_rng(3)
    w = make_weights(rng, c["N"], c["rows"], c["cols"], c["D"], c["Q"], c["C"], c["A"])
    cat, attr, box, n = synth_targets(rng, c["B"], c["T"], c["C"], c["A"], attr_p=0.3)
    feats = np.tanh(rng.standard_normal((c["B"], c["rows"], c["cols"], c["D"]))).astype(np.float32)
    return w, {"features": feats, "category": cat, "attribute": attr, "bbox": box, "num_objects": n}


def _replica_grads(w, batch, world):
    """d(sum_b loss_b / world)/d(theta) of one replica, flattened in sorted-name order (oracle, fp64)."""
    from oracle import reference_path as R
    tg = (batch["category"], batch["attribute"], batch["bbox"], batch["num_objects"])
    _, grads, _ = R.train_step_reference(w, batch["features"], tg, CFG["N"], CFG["H"], torch.float64, weights=R.model_weights(1.0))
    return torch.cat([torch.from_numpy(grads[k] / world).reshape(-1) for k in sorted(grads)])


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from boosted_detr_b200 import parallel
    r, wsz, _ = parallel.init_from_env(backend="gloo")
    assert (r, wsz) == (rank, world)
    w, batch = _data()
    shard = parallel.shard_batch(batch, rank, world)
    assert shard["features"].shape[0] == CFG["B"] // world
    flat = _replica_grads(w, shard, world)
    parallel.allreduce_gradients(flat)                 # the collective on the path
    # weights broadcast from rank 0
    t = torch.full((5,), float(rank))
    parallel.broadcast_tensor(t)
    q.put((rank, flat.numpy(), t.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 4])
def test_replicas_match_single_process_emulation(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from boosted_detr_b200 import parallel
    w, batch = _data()
    emu = sum(_replica_grads(w, parallel.shard_batch(batch, r, world), world) for r in range(world)).numpy()
    for rank, flat, t in results:
        assert np.allclose(flat, emu, rtol=1e-12, atol=1e-14), "all-reduced gradients differ from the emulation"
        assert (t == 0).all()
    # per-replica semantics differ from one big batch (BatchNorm statistics + loss normaliser are per replica)
    full = _replica_grads(w, batch, 1).numpy()
    assert not np.allclose(full, emu, rtol=1e-3)


def test_shard_batch_rejects_uneven_batches():
    import pytest
    from boosted_detr_b200 import parallel
    with pytest.raises(AssertionError):
        parallel.shard_batch({"features": np.zeros((3, 2))}, 0, 2)
