"""fp16-operand long-sequence attention (BDETR_MODE_FP16, attention_umma_ms_f16.cu; reference policy: Keras mixed_float16,
/root/reference/ModelComponents/parameters.py:73) against an fp64 softmax, through the C ABI."""
import numpy as np
import pytest
import torch

from test_dense_gpu import _model_and_data, nerr

pytestmark = pytest.mark.gpu


def _ref(q, k, v, B, Lq, Lk, H, d, rows=None):
    qh = q.astype(np.float64).reshape(B, Lq, H, d).transpose(0, 2, 1, 3)
    kh = k.astype(np.float64).reshape(B, Lk, H, d).transpose(0, 2, 1, 3)
    vh = v.astype(np.float64).reshape(B, Lk, H, d).transpose(0, 2, 1, 3)
    if rows is not None:
        qh = qh[:, :, rows]
    s = qh @ kh.transpose(0, 1, 3, 2) / np.sqrt(d)
    mx = s.max(-1, keepdims=True)
    p = np.exp(s - mx)
    return (p / p.sum(-1, keepdims=True)) @ vh, (np.log(p.sum(-1)) + mx[..., 0]) / np.log(2.0)


@pytest.mark.parametrize("B,Lq,Lk,qscale", [(2, 2100, 2100, 1.0), (1, 500, 1300, 1.0), (3, 129, 64, 1.0), (1, 4096, 4096, 1.0),
                                             (2, 700, 2100, 12.0)])
def test_attention_core_f16(B, Lq, Lk, qscale):
    """Same shapes as the tf32 multi-stream test (ragged query blocks, ragged last key tile, a single key tile, peaky
    scores that raise the softmax reference maximum in the middle of tiles -> in-TMEM rescale of the packed fp16 P).
    Bars: 1e-3 normalised max error against the fp64 softmax of the fp16-rounded operands (the kernel's arithmetic:
    P and the products in fp16, accumulation fp32), 2e-3 against the fp64 softmax of the UNROUNDED fp32 operands
    (adds the operand rounding itself, 2^-12 per element)."""
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.device import ptr, stream_ptr
    lib = _lib.load()
    H, d = 8, 32
    D = H * d
    rng = np.random.default_rng(Lq + Lk)
    q, k, v = (rng.standard_normal((B, L, D)).astype(np.float32) for L in (Lq, Lk, Lk))
    q = (q * qscale).astype(np.float32)
    dq, dk, dv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    o = torch.full((B, H, Lq, d), float("nan"), device="cuda"); lse = torch.full((B, H, Lq), float("nan"), device="cuda")
    try:
        assert lib.bdetr_debug_force_attention_kernel(2) == 0          # serve shapes below the production threshold too
        n = lib.bdetr_attention_f16_workspace_bytes(B, H, Lq, Lk, d)
        assert n == 2 * (B * Lq + 2 * B * Lk) * D
        ws = torch.empty(n // 2, dtype=torch.float16, device="cuda")
        _lib.call("bdetr_attention_core_fwd_f16", B, H, Lq, Lk, d, ptr(dq), ptr(dk), ptr(dv), ptr(ws), ptr(o), ptr(lse), stream_ptr())
        torch.cuda.synchronize()
    finally:
        lib.bdetr_debug_force_attention_kernel(0)
    o, lse = o.cpu().numpy(), lse.cpu().numpy()
    assert np.isfinite(o).all() and np.isfinite(lse).all()
    # the workspace holds the fp16 copies in q | k | v order
    w = ws.cpu().numpy()
    assert np.array_equal(w[:B * Lq * D], q.astype(np.float16).ravel())
    h = lambda x: x.astype(np.float16).astype(np.float32)
    ref_o, ref_l = _ref(h(q), h(k), h(v), B, Lq, Lk, H, d)
    e_o, e_l = nerr(o, ref_o), nerr(lse, ref_l)
    raw_o, raw_l = _ref(q, k, v, B, Lq, Lk, H, d)
    r_o, r_l = nerr(o, raw_o), nerr(lse, raw_l)
    print(f"fp16 attention core B{B} Lq{Lq} Lk{Lk} x{qscale}: vs fp16-rounded operands o {e_o:.2e} lse {e_l:.2e}; vs fp32 operands o {r_o:.2e} lse {r_l:.2e}")
    assert e_o < 1e-3 and e_l < 1e-4
    if qscale == 1.0:                                  # 12x scores: the operand rounding alone moves exp(s) by |s| 2^-12 ~ 2 %
        assert r_o < 2e-3 and r_l < 1e-3


def test_attention_core_f16_at_config5_length():
    """L = 20 020 (BASELINE config 5), all heads, one image: a strided subsample of rows against an fp64 softmax over all keys."""
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.device import ptr, stream_ptr
    lib = _lib.load()
    B, H, d, L = 1, 8, 32, 20020
    D = H * d
    rng = np.random.default_rng(20020)
    q = (rng.standard_normal((B, L, D)) * 1.5).astype(np.float32)
    k = rng.standard_normal((B, L, D)).astype(np.float32)
    v = rng.standard_normal((B, L, D)).astype(np.float32)
    dq, dk, dv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    o = torch.full((B, H, L, d), float("nan"), device="cuda"); lse = torch.full((B, H, L), float("nan"), device="cuda")
    n = lib.bdetr_attention_f16_workspace_bytes(B, H, L, L, d)
    assert n > 0, "config-5 length must be served by the fp16 kernel without forcing"
    ws = torch.empty(n // 2, dtype=torch.float16, device="cuda")
    _lib.call("bdetr_attention_core_fwd_f16", B, H, L, L, d, ptr(dq), ptr(dk), ptr(dv), ptr(ws), ptr(o), ptr(lse), stream_ptr())
    torch.cuda.synchronize()
    o, lse = o.cpu().numpy(), lse.cpu().numpy()
    assert np.isfinite(o).all() and np.isfinite(lse).all()
    rows = np.unique(np.concatenate([np.arange(0, L, 257), np.arange(L - 140, L), np.arange(0, 130)]))
    h = lambda x: x.astype(np.float16).astype(np.float32)
    ref_o, ref_l = _ref(h(q), h(k), h(v), B, L, L, H, d, rows=rows)
    e_o, e_l = nerr(o[:, :, rows], ref_o), nerr(lse[:, :, rows], ref_l)
    print(f"fp16 attention core L={L}: {len(rows)} sampled rows x {H} heads: o {e_o:.2e} lse {e_l:.2e}")
    assert e_o < 1e-3 and e_l < 1e-4


def test_f16_entry_refuses_unserved_shapes():
    from boosted_detr_b200 import _lib
    lib = _lib.load()
    assert lib.bdetr_attention_f16_workspace_bytes(16, 8, 400, 400, 32) == 0      # config 2: short sequences keep the TF32 kernels
    assert lib.bdetr_attention_f16_workspace_bytes(4, 8, 20020, 20020, 32) == 2 * 3 * 4 * 20020 * 256
    assert lib.bdetr_attention_f16_workspace_bytes(4, 8, 20020, 20020, 64) == 0


def test_model_inference_fp16_mode_vs_fp64_oracle():
    """BoostedDETR.call(training=False) in BDETR_MODE_FP16 at a sequence length the fp16 kernel serves without forcing
    (4 images x 48 x 40 = 1 920 encoder tokens, 160 CTAs), two boosted blocks, against the fp64 oracle: predictions within
    north_star's reduced-precision bar (1e-3 normalised max error), and the fp16 kernel must really have run
    (the launch counter differs from TF32 mode by the cast kernel of each encoder block)."""
    from oracle import reference_path as R
    from boosted_detr_b200 import _lib
    lib = _lib.load()
    N, B, rows, cols = 2, 4, 48, 40
    assert lib.bdetr_attention_f16_workspace_bytes(B, 8, rows * cols, rows * cols, 32) > 0
    model, w, inputs = _model_and_data(N=N, B=B, rows=rows, cols=cols)
    got, launches = {}, {}
    try:
        for name, mode in (("tf32", _lib.MODE_TF32), ("fp16", _lib.MODE_FP16)):
            lib.bdetr_set_mode(mode)
            assert lib.bdetr_get_mode() == mode
            model.call({"features": inputs["features"]}, training=False)
            lib.bdetr_reset_launch_count()
            got[name] = [t.cpu().numpy() for t in model.call({"features": inputs["features"]}, training=False)]
            torch.cuda.synchronize()
            launches[name] = int(lib.bdetr_launch_count())
    finally:
        lib.bdetr_set_mode(_lib.MODE_FP32)
    assert launches["fp16"] == launches["tf32"] + N, launches      # one cast launch per encoder self-attention
    saved = R.ATTENTION_QUERY_CHUNK
    R.ATTENTION_QUERY_CHUNK = 256
    try:
        out = R.boosted_detr_call(R.params_to_torch(w), torch.tensor(inputs["features"], dtype=torch.float64), None, N, 8, training=False)
    finally:
        R.ATTENTION_QUERY_CHUNK = saved
    for i, nm in enumerate(["cat", "attr", "box"]):
        ref = out["preds"][i].numpy()
        e16, e32 = nerr(got["fp16"][i], ref), nerr(got["tf32"][i], ref)
        print(f"inference {nm}: fp16 mode {e16:.2e}, tf32 mode {e32:.2e}")
        assert e16 < 1e-3, nm
