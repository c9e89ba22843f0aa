"""fp16 vs tf32 long-sequence attention core at BASELINE config 5's length (not a pytest file)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boosted_detr_b200 import _lib
from boosted_detr_b200.device import ptr, stream_ptr
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
B, H, d, L = 4, 8, 32, 20020
D = H * d
q = torch.randn(B, L, D, device="cuda"); k = torch.randn(B, L, D, device="cuda"); v = torch.randn(B, L, D, device="cuda")
o = torch.empty(B, H, L, d, device="cuda"); lse = torch.empty(B, H, L, device="cuda")
ws = torch.empty(lib.bdetr_attention_f16_workspace_bytes(B, H, L, L, d) // 2, dtype=torch.float16, device="cuda")
out = {}
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e-3
flops = 4.0 * B * L * L * D
for share in (0, 2, 4):
    lib.bdetr_debug_force_attention_kernel(20 + share)
    t16 = timeit(lambda: _lib.call("bdetr_attention_core_fwd_f16", B, H, L, L, d, ptr(q), ptr(k), ptr(v), ptr(ws), ptr(o), ptr(lse), stream_ptr()))
    t32 = timeit(lambda: _lib.call("bdetr_attention_core_fwd", B, H, L, L, d, ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), stream_ptr()))
    out[f"share{share}"] = {"fp16_us": t16 * 1e6, "fp16_tflops": flops / t16 / 1e12, "tf32_us": t32 * 1e6, "tf32_tflops": flops / t32 / 1e12}
    print(f"poly share {share}/8: fp16 (cast + kernel) {t16*1e6:.0f} us = {flops/t16/1e12:.0f} TFLOP/s | tf32 {t32*1e6:.0f} us = {flops/t32/1e12:.0f} TFLOP/s", flush=True)
lib.bdetr_debug_force_attention_kernel(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_attention_f16.json", "w"), indent=1)
