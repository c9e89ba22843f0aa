"""One eager training step of BASELINE config 2 inside a cudaProfilerStart/Stop range (for ncu launch lists)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from boosted_detr_b200 import _lib
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32 if mode == "tf32" else _lib.MODE_FP32)
model = bench.make_model(bench.CFG)
batch = bench.synth_batch(0, bench.CFG["B"], 82, 3, bench.CFG)
dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
for _ in range(2):
    model.train_step(dev, return_host=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.train_step(dev, return_host=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
