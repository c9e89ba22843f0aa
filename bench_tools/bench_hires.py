"""BASELINE config 5 (high-resolution inference) end to end (not a pytest file; written at the end of round 1 without
GPU access -- run it under a `timeout` first).  6 enc/dec pairs, 100 queries, 4 images per GPU, inference only:
  * the stress size of SURVEY 8: a 110 x 182 feature map = 20 020 encoder tokens per image (attention is ~95 % of the
    2.6 TFLOP per image), and
  * the reference-faithful size: 1333 x 800 at the backbone's stride 32 = 41 x 25 = 1 025 tokens.
Prints images/s, ms per batch and the algorithmic TFLOP/s (SURVEY 8d formulas), timed with CUDA events after warm-up.
usage: python bench_tools/bench_hires.py [B per GPU, default 4]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from boosted_detr_b200 import _lib
from boosted_detr_b200.boosted_model import BoostedDETR
from boosted_detr_b200.parameters import baseline_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32)
out = {}
for name, (rows, cols) in (("stress_20020_tokens", (110, 182)), ("reference_stride32_1025_tokens", (25, 41))):
    p = baseline_params(5)
    model = BoostedDETR(**p, attribute_weight=0.0, seed=0, feature_shape=(rows, cols))
    rng = np.random.default_rng(0)
    feats = torch.from_numpy(np.tanh(rng.standard_normal((B, rows, cols, 256))).astype(np.float32)).cuda()
    model.build(batch_size=1)
    # non-zero queries so the decoder does real work (zero-init queries make every query row identical)
    w = model.get_weights_dict()
    w["DecoderPrep/init_decoder_features"] = rng.normal(0, 0.02, w["DecoderPrep/init_decoder_features"].shape).astype(np.float32)
    model.set_weights_dict(w)
    fn = lambda: model.call({"features": feats}, training=False)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); preds = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    L, D, Q, N, C, A = rows * cols, 256, 100, 6, model.num_categories, model.num_attributes
    f_enc = 12 * L * D * D + 4 * L * L * D
    f_dec = 4 * L * D * D + 4 * Q * L * D + 8 * Q * D * D
    f_heads = 6 * Q * D * D + 2 * Q * D * (C + A + 4)
    f_self = 8 * Q * D * D + 4 * Q * Q * D
    flops = B * N * (f_enc + f_dec + f_heads) + (N - 1) * f_self
    assert all(torch.isfinite(t).all() for t in preds)
    out[name] = {"B": B, "L": L, "ms_per_batch": ms, "images_per_s": B / ms * 1e3, "algorithmic_tflops": flops / ms / 1e9}
    print(name, out[name], flush=True)
    del model, feats
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_hires.json", "w"), indent=1)
