"""Step-0 tie check at model level (not a pytest file yet: written at the end of round 1 without GPU access; turn it
into a test once it has run).  With the reference's initialisation -- DecoderPrep.init_decoder_features = zeros
(transformers.py:428-431) -- every query row is identical at step 0, so every block's predictions are identical across
queries, the cost matrix has identical columns, and scipy's tie rule assigns target row t to prediction column t.  The
product reproduces that only if identical rows stay BIT-identical through every kernel of the forward path (no
order-dependent atomics between rows).  usage: python tests/check_step0_ties.py [tf32|fp32]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_targets
from boosted_detr_b200 import _lib
from boosted_detr_b200.boosted_model import BoostedDETR
from boosted_detr_b200.parameters import baseline_params

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = _lib.load(); lib.bdetr_set_mode(_lib.MODE_TF32 if mode == "tf32" else _lib.MODE_FP32)
p = baseline_params(1)                                   # 2 enc/dec pairs, 100 queries, 20x20 features
model = BoostedDETR(**p, attribute_weight=1.0, seed=0).build()
model.dropout_seed = None                                # dropout masks differ per row: off for this check
assert float(np.abs(model.get_weights_dict()["DecoderPrep/init_decoder_features"]).max()) == 0.0
rng = np.random.default_rng(5)
B, T = 4, 20
cat, attr, box, n = synth_targets(rng, B, T, model.num_categories, model.num_attributes)
feats = np.tanh(rng.standard_normal((B, 20, 20, 256))).astype(np.float32)
model.train_step({"features": feats, "category": cat, "attribute": attr, "bbox": box, "num_objects": n})
ok = True
for i, c in enumerate(model.last_ctx_train["loss"]):
    cost = c["cost"].cpu().numpy(); c4r = c["col4row"].cpu().numpy()
    same_cols = bool((cost == cost[:, :, :1]).all())
    ident = all((c4r[b, :n[b]] == np.arange(n[b])).all() for b in range(B))
    print(f"block {i}: identical cost columns {same_cols}, identity assignment {ident}")
    ok = ok and same_cols and ident
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
