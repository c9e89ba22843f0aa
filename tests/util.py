"""Seeded synthetic inputs shaped like SURVEY.md §8d (shared by CPU and GPU tests and bench.py)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synth_targets(rng, B, T, C, A, attr_p=0.02, min_n=1):
    """category one-hot [B,T,C], attribute multi-hot [B,T,A], bbox [B,T,4] (pad -10), num_objects [B]."""
    n = rng.integers(min_n, T + 1, size=B).astype(np.int32)
    cat = np.zeros((B, T, C), np.float32)
    attr = np.zeros((B, T, A), np.float32)
    box = np.full((B, T, 4), -10.0, np.float32)
    for b in range(B):
        k = int(n[b])
        cls = rng.integers(2, C, size=k)
        cat[b, np.arange(k), cls] = 1.0
        cat[b, k:, 0] = 1.0                       # <PAD> -> class 0
        attr[b, :k] = (rng.random((k, A)) < attr_p).astype(np.float32)
        attr[b, k:, 0] = 1.0
        box[b, :k, 0:2] = rng.uniform(0.0, 0.8, size=(k, 2))
        box[b, :k, 2:4] = rng.uniform(0.02, 0.2, size=(k, 2))
    return cat, attr, box, n


def synth_preds(rng, B, Q, C, A, k_sum=None):
    """cumulative predictions: softmax(N(0,1)) * k, sigmoid attrs * k, boxes in the 3*sigmoid-1 range."""
    k = rng.integers(1, 8) if k_sum is None else k_sum
    z = rng.standard_normal((B, Q, C)).astype(np.float32)
    p = np.exp(z - z.max(-1, keepdims=True))
    cat = (p / p.sum(-1, keepdims=True) * k).astype(np.float32)
    attr = (1.0 / (1.0 + np.exp(-rng.standard_normal((B, Q, A)))) * rng.uniform(0.2, 1.5)).astype(np.float32)
    box = np.concatenate([rng.uniform(-0.1, 0.9, size=(B, Q, 2)), rng.uniform(-0.05, 0.4, size=(B, Q, 2))],
                         -1).astype(np.float32)
    return cat, attr, box


_lsap_c = None


def lsap_c():
    """The C restatement in oracle/ (built on demand with gcc)."""
    global _lsap_c
    if _lsap_c is None:
        path = os.path.join(ROOT, "oracle", "liblsap_ref.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _lsap_c = ctypes.CDLL(path)
        _lsap_c.lsap_ref_f32.restype = ctypes.c_int
        _lsap_c.lsap_ref_batch_mask.restype = ctypes.c_int
    return _lsap_c


def lsap_c_solve(c):
    c = np.ascontiguousarray(c, np.float32)
    nr, nc = c.shape
    a = np.zeros(max(nr, nc, 1), np.int64)
    b = np.zeros(max(nr, nc, 1), np.int64)
    k = lsap_c().lsap_ref_f32(c.ctypes.data_as(ctypes.c_void_p), nr, nc, nc,
                              a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p))
    return k, a[:max(k, 0)], b[:max(k, 0)]


def make_weights(rng, N, rows, cols, D, Q, C, A, perturb=True):
    """All hot-path weights with the reference's Keras variable names, built on the
    host with numpy only (usable by the oracle without a GPU)."""
    def tn(shape, std):
        x = rng.standard_normal(shape)
        return (np.clip(x, -2, 2) * std).astype(np.float32)
    w = {}
    b = (lambda n: rng.normal(0, 0.1, n).astype(np.float32)) if perturb else (lambda n: np.zeros(n, np.float32))
    g = (lambda n: (1 + rng.normal(0, 0.1, n)).astype(np.float32)) if perturb else (lambda n: np.ones(n, np.float32))

    def attn(pfx):
        for nm in ("QueryProjection", "KeyProjection", "ValueProjection", "OutputProjection"):
            w[f"{pfx}/AttentionLayer/{nm}/kernel"] = tn((D, D), np.sqrt(1.0 / D))
            w[f"{pfx}/AttentionLayer/{nm}/bias"] = b(D)
        w[f"{pfx}/LayerNorm/gamma"], w[f"{pfx}/LayerNorm/beta"] = g(D), b(D)

    def ffn(pfx):
        for nm in ("DenseRelu", "DenseLinear"):
            w[f"{pfx}/{nm}/kernel"] = tn((D, D), np.sqrt(1.0 / D))
            w[f"{pfx}/{nm}/bias"] = b(D)
        w[f"{pfx}/LayerNorm/gamma"], w[f"{pfx}/LayerNorm/beta"] = g(D), b(D)

    def head(pfx, d1, d2, nout):
        w[f"{pfx}/{d1}/kernel"], w[f"{pfx}/{d1}/bias"] = tn((D, D), np.sqrt(2.0 / D)), b(D)
        w[f"{pfx}/BatchNorm/gamma"], w[f"{pfx}/BatchNorm/beta"] = g(D), b(D)
        w[f"{pfx}/BatchNorm/moving_mean"] = b(D)
        w[f"{pfx}/BatchNorm/moving_variance"] = rng.uniform(0.5, 1.5, D).astype(np.float32) if perturb else np.ones(D, np.float32)
        w[f"{pfx}/{d2}/kernel"], w[f"{pfx}/{d2}/bias"] = tn((D, nout), np.sqrt(2.0 / (D + nout))), b(nout)

    k = np.arange(rows * cols, dtype=np.float64)[:, None]
    den = 2.0 * (1.0 + np.arange(D, dtype=np.float64))[None, :] / D
    pos = np.where((k % 2) == 1, np.sin(k / den), np.cos(k / den)).reshape(rows, cols, D).astype(np.float32)
    for i in range(N):
        w[f"ImageEncoderAttention_{i}/positional_encoding"] = pos.copy()
        attn(f"ImageEncoderAttention_{i}/EncoderBlock_0/SelfAttentionBlock")
        ffn(f"ImageEncoderAttention_{i}/EncoderBlock_0/FeedForwardBlock")
        if i >= 1:
            attn(f"DecoderBlock_{i}/SelfAttentionBlock")
        attn(f"DecoderBlock_{i}/JointAttentionBlock")
        ffn(f"DecoderBlock_{i}/FeedForwardBlock")
        head(f"CategoryPredictionHead_{i}", "DenseCateg", "DenseLogits", C)
        head(f"AttributePredictionHead_{i}", "Dense", "DenseLinear", A)
        head(f"BoxPredictionHead_{i}", "Dense", "BoxCoords", 4)
    w["DecoderPrep/init_decoder_features"] = rng.normal(0, 0.5 if perturb else 0.02, (Q, D)).astype(np.float32)
    return w
