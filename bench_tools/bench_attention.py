"""Micro-benchmark of the attention core (not a pytest file): BASELINE config 2 and config 5 shapes."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr

lib = _lib.load()
out = {}
for mode_name, mode in (("tf32", _lib.MODE_TF32), ("fp32", _lib.MODE_FP32)):
    lib.bdetr_set_mode(mode)
    for (B, Lq, Lk) in [(16, 400, 400), (16, 100, 400), (4, 1050, 1050), (4, 20020, 20020)]:
        if mode_name == "fp32" and Lq > 2000:
            continue
        H, d = 8, 32
        D = H * d
        q = torch.randn(B, Lq, D, device="cuda"); k = torch.randn(B, Lk, D, device="cuda"); v = torch.randn(B, Lk, D, device="cuda")
        o = torch.empty(B, H, Lq, d, device="cuda"); lse = torch.empty(B, H, Lq, device="cuda")
        fn = lambda: _lib.call("bdetr_attention_core_fwd", B, H, Lq, Lk, d, ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), stream_ptr())
        # tensor-core mode, long sequences: time both schedules (1 = one tile per CTA, 2 = three streams per CTA)
        variants = (1, 2) if (mode_name == "tf32" and Lq >= 1000) else (0,)
        if mode_name == "tf32" and Lq >= 20000:
            variants = (1, 20, 21, 22, 23, 24)        # 20 + n: n of 8 exponential groups on the FMA pipe
        for which in variants:
            lib.bdetr_debug_force_attention_kernel(which)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            iters = 20 if Lq < 5000 else 5
            ts = []
            for _ in range(iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = float(np.median(ts)) * 1e-3
            flops = 4.0 * B * Lq * Lk * D
            out[f"{mode_name}_B{B}_Lq{Lq}_Lk{Lk}_kernel{which}"] = {"us": t * 1e6, "tflops": flops / t / 1e12}
            print(mode_name, B, Lq, Lk, "kernel", which, f"{t*1e6:.1f} us  {flops/t/1e12:.1f} TFLOP/s", flush=True)
        lib.bdetr_debug_force_attention_kernel(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_attention.json", "w"), indent=1)
