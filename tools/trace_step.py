"""Timeline of one CUDA-graph-replayed training step at BASELINE config 2 (not a pytest file): the model drops timing
events (external event-record nodes) at the end of every phase of every boosted block; printed as milliseconds from
the start of the step, per stream.  usage: python tools/trace_step.py [tf32|fp32]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from boosted_detr_b200 import _lib
from boosted_detr_b200.graph import GraphedTrainStep
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32 if mode == "tf32" else _lib.MODE_FP32)
from boosted_detr_b200.parallel import DataParallel, init_from_env
rank, world, local = init_from_env()            # under torchrun: the data-parallel step (bucketed all-reduce inside the graph)
torch.cuda.set_device(local)
model = bench.make_model(bench.CFG)
DataParallel(model)
batch = bench.synth_batch(rank, bench.CFG["B"], 82, 3, bench.CFG)
model._trace = []
if "--fine" in sys.argv:
    from boosted_detr_b200 import transformers
    transformers.TRACE_HOOK = model._mark
gs = GraphedTrainStep(model, batch)
marks = list(model._trace)
model._trace = None
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
acc = np.zeros(len(marks))
reps = 10
for _ in range(3):
    gs.replay()
torch.cuda.synchronize()
for _ in range(reps):
    flush.zero_()
    gs.replay()
    torch.cuda.synchronize()
    t0 = marks[0][1]
    acc += np.array([t0.elapsed_time(ev) for _, ev in marks])
acc /= reps
if rank != 0:
    os._exit(0)
prev = {}
for (label, _), t in sorted(zip(marks, acc), key=lambda x: x[1]):
    stream = label[label.rfind("("):]
    d = t - prev.get(stream, 0.0)
    prev[stream] = t
    print(f"{t:8.3f} ms  (+{d:6.3f} on {stream:7s})  {label}")
