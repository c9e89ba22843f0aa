"""Boosted encoder / decoder transformer layers — same class names, constructor arguments and call
conventions as the reference's ModelComponents/transformers.py, executed by libbdetr.so."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .device import empty, f32, ptr, stream_ptr
from .layers import Layer, glorot_normal

DROPOUT_RATE = 0.1   # reference transformers.py:135,179
LN_EPS = 1e-3        # reference :137 (explicit) / Keras default (:180)


def _struct(cls, tensors: dict):
    s = cls()
    for k, _ in cls._fields_:
        t = tensors.get(k)
        setattr(s, k, None if t is None else t.data_ptr())
    return s


class MultiheadAttention(Layer):
    """Weight holder for the four Dense projections (reference :18-109).  The computation runs inside
    AttentionBlock's fused entry point; heads are column slices, and the [B,H,Lq,d] output is re-read
    as [B,Lq,H*d] without a permute, exactly like reference line :100."""

    def __init__(self, num_attention_heads, dim, name="MultiheadAttention", **kwargs):
        super().__init__(name=name)
        self.num_attention_heads = num_attention_heads
        self.dim = dim

    def get_config(self):
        return {**super().get_config(), "num_attention_heads": self.num_attention_heads, "dim": self.dim}

    def build(self, input_shape):
        query_dim = input_shape[0][-1]
        proj = self.num_attention_heads * self.dim
        rng = Layer._rng
        for nm, (fi, fo) in {"QueryProjection": (input_shape[0][-1], proj), "KeyProjection": (input_shape[1][-1], proj),
                             "ValueProjection": (input_shape[2][-1], proj), "OutputProjection": (proj, query_dim)}.items():
            self.add_weight(f"{nm}/kernel", glorot_normal(rng, fi, fo))
            self.add_weight(f"{nm}/bias", np.zeros(fo, np.float32))

    def forward(self, inputs, training=False, attention_mask=None):
        """Stand-alone MultiheadAttention.call (reference :68-102), inference only: the three projections, the attention
        core and the output projection through bdetr_gemm / bdetr_attention_core_fwd.  Inside the model the layer runs
        fused in AttentionBlock's entry points (which also provide the backward).  `attention_mask` multiplies the
        probabilities AFTER the softmax in the reference (:92-94) and is all ones on this path: only None is accepted."""
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is always ones on the reference's path; pass None")
        query, key, value = (f32(t) for t in inputs)
        self.maybe_build([query, key, value])
        B, Lq, D = query.shape
        Lk, H, d = key.shape[1], self.num_attention_heads, self.dim
        P = H * d
        w, _ = self.gemm_weights()
        rnd = 1 if _lib.tc_mode() else 0

        def dense(x, nm, rows, n_in, n_out):
            y = empty(rows, n_out)
            _lib.call("bdetr_gemm", rows, n_out, n_in, ptr(x), 0, ptr(w[f"{nm}/kernel"]), 0, ptr(self._weights[f"{nm}/bias"]), 0, 0,
                      ptr(y), stream_ptr())
            if rnd:
                _lib.call("bdetr_round_tf32", y.numel(), ptr(y), ptr(y), stream_ptr())
            return y
        qp = dense(query, "QueryProjection", B * Lq, D, P)
        kp = dense(key, "KeyProjection", B * Lk, key.shape[-1], P)
        vp = dense(value, "ValueProjection", B * Lk, value.shape[-1], P)
        o, lse = empty(B, H, Lq, d), empty(B, H, Lq)
        _lib.call("bdetr_attention_core_fwd", B, H, Lq, Lk, d, ptr(qp), ptr(kp), ptr(vp), ptr(o), ptr(lse), stream_ptr())
        out = empty(B, Lq, w["OutputProjection/kernel"].shape[1])
        # [B,H,Lq,d] re-read as [B*Lq, H*d] without a permute (reference :100, quirk Q1)
        _lib.call("bdetr_gemm", B * Lq, out.shape[-1], P, ptr(o), 0, ptr(w["OutputProjection/kernel"]), 0,
                  ptr(self._weights["OutputProjection/bias"]), 0, 0, ptr(out), stream_ptr())
        return out, {"qp": qp, "kp": kp, "vp": vp, "o": o, "lse": lse}


class AttentionBlock(Layer):
    """LayerNorm(query + Dropout(MHA(query, key, value)))  (reference :112-158)."""

    def __init__(self, num_attention_heads, name="AttentionBlock", **kwargs):
        super().__init__(name=name)
        self.num_attention_heads = num_attention_heads
        self.AttentionLayer = None
        self.rate = DROPOUT_RATE

    def get_config(self):
        return {**super().get_config(), "num_attention_heads": self.num_attention_heads}

    def build(self, input_shape):
        D = input_shape[0][-1]
        assert D % self.num_attention_heads == 0
        self.AttentionLayer = MultiheadAttention(self.num_attention_heads, D // self.num_attention_heads, name="AttentionLayer")
        self.AttentionLayer.build(input_shape)
        self.AttentionLayer.built = True
        self.add_weight("LayerNorm/gamma", np.ones(D, np.float32))
        self.add_weight("LayerNorm/beta", np.zeros(D, np.float32))

    def _structs(self):
        a = self.AttentionLayer
        wa, mode = a.gemm_weights()
        if self._struct_cache is None or self._struct_cache[2] != mode:
            def pack(src_a, src_s):
                return _struct(_lib.AttnParams, {
                    "wq": src_a["QueryProjection/kernel"], "bq": src_a["QueryProjection/bias"],
                    "wk": src_a["KeyProjection/kernel"], "bk": src_a["KeyProjection/bias"],
                    "wv": src_a["ValueProjection/kernel"], "bv": src_a["ValueProjection/bias"],
                    "wo": src_a["OutputProjection/kernel"], "bo": src_a["OutputProjection/bias"],
                    "ln_gamma": src_s["LayerNorm/gamma"], "ln_beta": src_s["LayerNorm/beta"]})
            self._struct_cache = (pack(wa, self._weights), pack(a._grads, self._grads), mode)
        return self._struct_cache[:2]

    def forward(self, inputs, training=False, dropout_key=0, seed_dev=None):
        """`seed_dev` (optional device uint32 word): the kernels then derive the mask key as
        lowbias32(*seed_dev ^ dropout_key), so a captured CUDA graph draws a fresh mask per replay."""
        query, key, value = (f32(t) for t in inputs)
        self.maybe_build([query, key, value])
        B, Lq, D = query.shape
        Lk, H = key.shape[1], self.num_attention_heads
        sv = {"qp": empty(B, Lq, D), "kp": empty(B, Lk, D), "vp": empty(B, Lk, D), "o": empty(B, H, Lq, D // H),
              "lse": empty(B, H, Lq), "z": empty(B, Lq, D), "mean": empty(B * Lq), "rstd": empty(B * Lq)}
        out = empty(B, Lq, D)
        rate = self.rate if training else 0.0
        w, _ = self._structs()
        svs = _struct(_lib.AttnSaved, sv)
        _lib.call("bdetr_attention_block_fwd", B, Lq, Lk, D, H, ptr(query), ptr(key), ptr(value), ctypes.byref(w),
                  rate, dropout_key, ptr(seed_dev), LN_EPS, ptr(out), ctypes.byref(svs), stream_ptr())
        ctx = {"inputs": (query, key, value), "saved": sv, "saved_struct": svs, "rate": rate, "key": dropout_key,
               "seed_dev": seed_dev, "dims": (B, Lq, Lk, D, H)}
        return out, ctx

    def backward(self, ctx, d_out, d_query=None, d_key=None, d_value=None, acc=(False, False, False)):
        """Returns (d_query, d_key, d_value); parameter gradients are accumulated into self._grads.
        Pass existing buffers (+acc flags) to accumulate into them.  query-is-key inputs share one buffer."""
        query, key, value = ctx["inputs"]
        B, Lq, Lk, D, H = ctx["dims"]
        acc = list(acc)
        if d_query is None:
            d_query, acc[0] = empty(B, Lq, D), False
        if d_key is None:
            if key is query:
                d_key, acc[1] = d_query, True
            else:
                d_key, acc[1] = empty(B, Lk, D), False
        if d_value is None:
            if value is query:
                d_value, acc[2] = d_query, True
            elif value is key:
                d_value, acc[2] = d_key, True
            else:
                d_value, acc[2] = empty(B, Lk, D), False
        sc = {"d_qp": empty(B, Lq, D), "d_kp": empty(B, Lk, D), "d_vp": empty(B, Lk, D), "d_o": empty(B, H, Lq, D // H),
              "d_z": empty(B, Lq, D), "delta": empty(B, H, Lq)}
        w, gw = self._structs()
        scs = _struct(_lib.AttnScratch, sc)
        flags = (1 if acc[0] else 0) | (2 if acc[1] else 0) | (4 if acc[2] else 0)
        _lib.call("bdetr_attention_block_bwd", B, Lq, Lk, D, H, ptr(query), ptr(key), ptr(value), ctypes.byref(w),
                  ctx["rate"], ctx["key"], ptr(ctx["seed_dev"]), ctypes.byref(ctx["saved_struct"]), ptr(f32(d_out)),
                  ptr(d_query), ptr(d_key), ptr(d_value), flags, ctypes.byref(gw), ctypes.byref(scs), stream_ptr())
        return d_query, d_key, d_value


def fused_path(D=256):
    """The fused tensor-core entry points (include/bdetr.h) serve tensor-core mode at the model width 256."""
    return D == 256 and _lib.tc_mode()


def make_fold(pos=None, tab_q=None, tab_k=None, resid_pos=False, pos_tc=None):
    f = _lib.PosFold()
    f.pos = None if pos is None else pos.data_ptr()
    f.tab_q = None if tab_q is None else tab_q.data_ptr()
    f.tab_k = None if tab_k is None else tab_k.data_ptr()
    f.resid_pos = 1 if resid_pos else 0
    f.pos_tc = None if pos_tc is None else pos_tc.data_ptr()
    f._keep = (pos, tab_q, tab_k, pos_tc)
    return f


def pos_projection(pos_tc, layers_and_names):
    """tab[g] = pos W[g] + b[g] for up to three (AttentionLayer, 'QueryProjection' | 'KeyProjection') pairs: the
    positional term of a projection, batch-invariant, one grouped GEMM (bdetr_pos_projection)."""
    L, D = pos_tc.shape
    n = len(layers_and_names)
    Ws, bs, tabs = _lib.PTR3(), _lib.PTR3(), _lib.PTR3()
    out = []
    for g, (layer, nm) in enumerate(layers_and_names):
        w, _ = layer.gemm_weights()
        t = empty(L, D)
        Ws[g], bs[g], tabs[g] = w[f"{nm}/kernel"].data_ptr(), layer._weights[f"{nm}/bias"].data_ptr(), t.data_ptr()
        out.append(t)
    _lib.call("bdetr_pos_projection", L, D, ptr(pos_tc), n, ctypes.byref(Ws), ctypes.byref(bs), ctypes.byref(tabs), stream_ptr())
    return out


def _attn_forward_fused(self, query, memory, fold, training=False, dropout_key=0, seed_dev=None):
    """AttentionBlock with the positional adds folded into the projections (bdetr_attention_fused_fwd): `memory` feeds both
    the key and the value projection; fold = make_fold(...)."""
    query, memory = f32(query), (query if memory is query else f32(memory))
    self.maybe_build([query, memory, memory])
    B, Lq, D = query.shape
    Lk, H = memory.shape[1], self.num_attention_heads
    sv = {"qp": empty(B, Lq, D), "kp": empty(B, Lk, D), "vp": empty(B, Lk, D), "o": empty(B, H, Lq, D // H),
          "lse": empty(B, H, Lq), "z": empty(B, Lq, D) if training else None, "mean": empty(B * Lq), "rstd": empty(B * Lq)}
    out = empty(B, Lq, D)
    rate = self.rate if training else 0.0
    w, _ = self._structs()
    lib = _lib.load()
    if lib.bdetr_get_mode() == _lib.MODE_FP16 and not training:      # fp16 copies of q / k / v for the long-sequence attention kernel (inference
        # forward only: the backward kernels recompute S from the TF32 operands, and that mixed pairing is not tested)
        n16 = lib.bdetr_attention_f16_workspace_bytes(B, H, Lq, Lk, D // H)
        if n16:
            sv["ws16"] = torch.empty(n16 // 2, dtype=torch.float16, device=query.device)
    svs = _struct(_lib.AttnSaved, sv)
    _lib.call("bdetr_attention_fused_fwd", B, Lq, Lk, D, H, ptr(query), ptr(memory), ctypes.byref(fold) if fold is not None else None,
              ctypes.byref(w), rate, dropout_key, ptr(seed_dev), LN_EPS, 1 if training else 0, ptr(out), ctypes.byref(svs), stream_ptr())
    ctx = {"fused": True, "query": query, "memory": memory, "fold": fold, "saved": sv, "saved_struct": svs, "rate": rate,
           "key": dropout_key, "seed_dev": seed_dev, "dims": (B, Lq, Lk, D, H)}
    return out, ctx


def _attn_backward_fused(self, ctx, d_out, d_query=None, d_memory=None, acc=(False, False), d_pos=None, need_dq=True, need_dm=True):
    """Returns (d_query, d_memory) (None where skipped).  Parameter gradients are skipped for a frozen layer."""
    query, memory = ctx["query"], ctx["memory"]
    B, Lq, Lk, D, H = ctx["dims"]
    acc = list(acc)
    self_attn = memory is query
    if need_dq and d_query is None:
        d_query, acc[0] = empty(B, Lq, D), False
    if not self_attn and need_dm and d_memory is None:
        d_memory, acc[1] = empty(B, Lk, D), False
    sc = {"d_qp": empty(B, Lq, D), "d_kp": empty(B, Lk, D), "d_vp": empty(B, Lk, D), "d_o": empty(B, H, Lq, D // H),
          "d_z": empty(B, Lq, D), "delta": empty(B, H, Lq)}
    d_resid = empty(B, Lq, D)
    sums = empty(4, max(Lq, Lk), D)
    w, gw = self._structs()
    scs = _struct(_lib.AttnScratch, sc)
    flags = (1 if acc[0] else 0) | (2 if acc[1] else 0)
    train = self.trainable
    _lib.call("bdetr_attention_fused_bwd", B, Lq, Lk, D, H, ptr(query), ptr(memory),
              ctypes.byref(ctx["fold"]) if ctx["fold"] is not None else None, ctypes.byref(w), ctx["rate"], ctx["key"],
              ptr(ctx["seed_dev"]), ctypes.byref(ctx["saved_struct"]), ptr(f32(d_out)),
              ptr(d_query if need_dq else None), ptr(d_memory if (need_dm and not self_attn) else None), flags,
              ptr(d_pos if train else None), ctypes.byref(gw) if train else None, ctypes.byref(scs), ptr(d_resid), ptr(sums), stream_ptr())
    ctx["_keep_bwd"] = (sc, d_resid, sums)
    return d_query, d_memory


AttentionBlock.forward_fused = _attn_forward_fused
AttentionBlock.backward_fused = _attn_backward_fused


def _self_forward_hoisted(self, q0, q0_tc, B, training=False, dropout_key=0, seed_dev=None):
    """Decoder self-attention block on the shared [Q,D] queries, hoisted out of the batch (bdetr_decoder_self_fwd)."""
    Q, D = q0.shape
    self.maybe_build([q0.view(1, Q, D)] * 3)
    H = self.num_attention_heads
    sv = {"qp": empty(Q, D), "kp": empty(Q, D), "vp": empty(Q, D), "o": empty(H, Q, D // H), "lse": empty(H, Q),
          "z": empty(B, Q, D) if training else None, "mean": empty(B * Q), "rstd": empty(B * Q)}
    mha, out = empty(Q, D), empty(B, Q, D)
    rate = self.rate if training else 0.0
    w, _ = self._structs()
    svs = _struct(_lib.AttnSaved, sv)
    _lib.call("bdetr_decoder_self_fwd", B, Q, D, H, ptr(q0), ptr(q0_tc), ctypes.byref(w), rate, dropout_key, ptr(seed_dev), LN_EPS,
              1 if training else 0, ptr(out), ctypes.byref(svs), ptr(mha), stream_ptr())
    ctx = {"hoisted": True, "q0": q0, "q0_tc": q0_tc, "saved": sv, "saved_struct": svs, "mha": mha, "rate": rate, "key": dropout_key,
           "seed_dev": seed_dev, "dims": (B, Q, D, H)}
    return out, ctx


def _self_backward_hoisted(self, ctx, d_out, d_q0):
    """Accumulates into d_q0 [Q,D] (the shared query parameter's gradient) and, unless frozen, the layer's gradients."""
    B, Q, D, H = ctx["dims"]
    sc = {"d_qp": empty(Q, D), "d_kp": empty(Q, D), "d_vp": empty(Q, D), "d_o": empty(H, Q, D // H), "d_z": empty(B, Q, D),
          "delta": empty(H, Q)}
    d_resid, sums = empty(B, Q, D), empty(2, Q, D)
    w, gw = self._structs()
    scs = _struct(_lib.AttnScratch, sc)
    _lib.call("bdetr_decoder_self_bwd", B, Q, D, H, ptr(ctx["q0"]), ptr(ctx["q0_tc"]), ctypes.byref(w), ctx["rate"], ctx["key"],
              ptr(ctx["seed_dev"]), ctypes.byref(ctx["saved_struct"]), ptr(f32(d_out)), ptr(d_q0),
              ctypes.byref(gw) if self.trainable else None, ctypes.byref(scs), ptr(d_resid), ptr(sums), stream_ptr())
    ctx["_keep_bwd"] = (sc, d_resid, sums)


AttentionBlock.forward_hoisted = _self_forward_hoisted
AttentionBlock.backward_hoisted = _self_backward_hoisted


class FeedForwardBlock(Layer):
    """LayerNorm(x + Dropout(DenseLinear(DenseRelu(x))))  (reference :161-198); hidden width = feature dim."""

    def __init__(self, name="FeedForwardBlock", **kwargs):
        super().__init__(name=name)
        self.rate = DROPOUT_RATE

    def build(self, input_shape):
        D = input_shape[0][-1]
        rng = Layer._rng
        self.add_weight("DenseRelu/kernel", glorot_normal(rng, D, D))
        self.add_weight("DenseRelu/bias", np.zeros(D, np.float32))
        self.add_weight("DenseLinear/kernel", glorot_normal(rng, D, D))
        self.add_weight("DenseLinear/bias", np.zeros(D, np.float32))
        self.add_weight("LayerNorm/gamma", np.ones(D, np.float32))
        self.add_weight("LayerNorm/beta", np.zeros(D, np.float32))

    def _structs(self):
        w, mode = self.gemm_weights()
        if self._struct_cache is None or self._struct_cache[2] != mode:
            pack = lambda s: _struct(_lib.FfnParams, {"w1": s["DenseRelu/kernel"], "b1": s["DenseRelu/bias"],
                                                      "w2": s["DenseLinear/kernel"], "b2": s["DenseLinear/bias"],
                                                      "ln_gamma": s["LayerNorm/gamma"], "ln_beta": s["LayerNorm/beta"]})
            self._struct_cache = (pack(w), pack(self._grads), mode)
        return self._struct_cache[:2]

    def forward(self, inputs, training=False, dropout_key=0, seed_dev=None):
        x = f32(inputs[0])
        self.maybe_build([x])
        D = x.shape[-1]
        M = x.numel() // D
        fused = fused_path(D) and M >= 128
        sv = {"h": empty(M, D), "z": empty(M, D) if (training or not fused) else None, "mean": empty(M), "rstd": empty(M)}
        out = torch.empty_like(x)
        rate = self.rate if training else 0.0
        w, _ = self._structs()
        svs = _struct(_lib.FfnSaved, sv)
        if fused:
            _lib.call("bdetr_ffn_fused_fwd", M, D, ptr(x), ctypes.byref(w), rate, dropout_key, ptr(seed_dev), LN_EPS,
                      1 if training else 0, ptr(out), ctypes.byref(svs), stream_ptr())
        else:
            _lib.call("bdetr_ffn_block_fwd", M, D, ptr(x), ctypes.byref(w), rate, dropout_key, ptr(seed_dev), LN_EPS,
                      ptr(out), ctypes.byref(svs), stream_ptr())
        return out, {"x": x, "saved": sv, "saved_struct": svs, "rate": rate, "key": dropout_key, "seed_dev": seed_dev,
                     "dims": (M, D), "fused": fused}

    def backward(self, ctx, d_out, d_x=None, acc=False):
        M, D = ctx["dims"]
        if d_x is None:
            d_x, acc = torch.empty_like(ctx["x"]), False
        sc = {"d_z": empty(M, D), "d_h": empty(M, D)}
        w, gw = self._structs()
        scs = _struct(_lib.FfnScratch, sc)
        if ctx.get("fused"):
            _lib.call("bdetr_ffn_fused_bwd", M, D, ptr(ctx["x"]), ctypes.byref(w), ctx["rate"], ctx["key"],
                      ptr(ctx["seed_dev"]), ctypes.byref(ctx["saved_struct"]), ptr(f32(d_out)), ptr(d_x), 1 if acc else 0,
                      ctypes.byref(gw) if self.trainable else None, ctypes.byref(scs), stream_ptr())
        else:
            _lib.call("bdetr_ffn_block_bwd", M, D, ptr(ctx["x"]), ctypes.byref(w), ctx["rate"], ctx["key"],
                      ptr(ctx["seed_dev"]), ctypes.byref(ctx["saved_struct"]), ptr(f32(d_out)), ptr(d_x), 1 if acc else 0, ctypes.byref(gw),
                      ctypes.byref(scs), stream_ptr())
        ctx["_keep_bwd"] = sc
        return d_x


TRACE_HOOK = None          # tests/trace_step.py: callable(label) dropping a timing event on the current stream


def add_positional(x, pos):
    B, L, D = x.shape
    out = torch.empty_like(x)
    _lib.call("bdetr_add_positional_fwd", B, L, D, ptr(x), ptr(pos), ptr(out), stream_ptr())
    return out


def batch_sum_into(d_out, d_param):
    B, L, D = d_out.shape
    _lib.call("bdetr_add_positional_bwd", B, L, D, ptr(d_out), ptr(d_param), stream_ptr())


def accumulate(x, y):
    _lib.call("bdetr_accumulate", x.numel(), ptr(x), ptr(y), stream_ptr())


class EncoderBlock(Layer):
    """q = k = x + pos, v = x -> AttentionBlock -> FeedForwardBlock  (reference :200-241).
    `encoder_positional` is the un-tiled [L,D] table (batch-invariant; the reference tiles it)."""

    def __init__(self, num_attention_heads, name="EncoderBlock", **kwargs):
        super().__init__(name=name)
        self.num_attention_heads = num_attention_heads
        self.SelfAttentionBlock = AttentionBlock(num_attention_heads, name="SelfAttentionBlock")
        self.FeedForwardBlock = FeedForwardBlock(name="FeedForwardBlock")

    def get_config(self):
        return {**super().get_config(), "num_attention_heads": self.num_attention_heads}

    def forward(self, inputs, training=False, dropout_keys=(0, 0), seed_dev=None, pos_tc=None, tabs=None):
        """Tensor-core mode: q = (x+pos) Wq = x Wq + pos Wq -- the positional add is folded into the grouped q/k/v
        projection as a row table (`tabs` = (pos Wq + bq, pos Wk + bk), computed here when the caller did not), and the
        residual x + pos is formed inside the output-projection + LayerNorm kernel.  `pos_tc`: tf32-rounded pos."""
        x, pos = inputs
        x = f32(x)
        if fused_path(x.shape[-1]) and x.shape[0] * x.shape[1] >= 128:
            att = self.SelfAttentionBlock
            att.maybe_build([x, x, x])
            pos_tc = pos if pos_tc is None else pos_tc
            if tabs is None:
                tabs = pos_projection(pos_tc, [(att.AttentionLayer, "QueryProjection"), (att.AttentionLayer, "KeyProjection")])
            fold = make_fold(pos=pos, tab_q=tabs[0], tab_k=tabs[1], resid_pos=True, pos_tc=pos_tc)
            a, c1 = att.forward_fused(x, x, fold, training, dropout_keys[0], seed_dev)
            y, c2 = self.FeedForwardBlock.forward([a], training, dropout_keys[1], seed_dev)
            return y, {"attn": c1, "ffn": c2, "fused": True}
        xp = add_positional(x, pos)                                   # Add1 == Add2 (:226-227)
        a, c1 = self.SelfAttentionBlock.forward([xp, xp, x], training, dropout_keys[0], seed_dev)
        y, c2 = self.FeedForwardBlock.forward([a], training, dropout_keys[1], seed_dev)
        return y, {"attn": c1, "ffn": c2}

    def backward(self, ctx, d_out, d_pos, d_x=None, acc=False, need_dx=True):
        """Returns d_x; accumulates the positional gradient (summed over the batch) into d_pos.  Fused path: d_x may be
        an existing buffer to accumulate into (acc=True); need_dx=False skips the input gradient."""
        if ctx.get("fused"):
            if not need_dx and not self.trainable:
                return None                                               # frozen and nothing trainable upstream
            d_a = self.FeedForwardBlock.backward(ctx["ffn"], d_out)
            d_x, _ = self.SelfAttentionBlock.backward_fused(ctx["attn"], d_a, d_query=d_x, acc=(acc, False), d_pos=d_pos, need_dq=need_dx)
            return d_x
        d_a = self.FeedForwardBlock.backward(ctx["ffn"], d_out)
        if TRACE_HOOK is not None:
            TRACE_HOOK(f"  {self.name} ffn bwd done (main)")
        d_xp, _, d_x = self.SelfAttentionBlock.backward(ctx["attn"], d_a)   # d_key aliases d_query (same tensor)
        if TRACE_HOOK is not None:
            TRACE_HOOK(f"  {self.name} attn bwd done (main)")
        batch_sum_into(d_xp, d_pos)
        accumulate(d_xp, d_x)
        return d_x


def positional_table(rows, cols, dim):
    """Sine/cosine initial value of the trainable table (reference :282-291): odd flattened position ->
    sin, even -> cos, denominator 2(1+dim)/D.  Vectorised (the reference uses a Python double loop)."""
    k = np.arange(rows * cols, dtype=np.float64)[:, None]
    den = 2.0 * (1.0 + np.arange(dim, dtype=np.float64))[None, :] / dim
    tab = np.where((k % 2) == 1, np.sin(k / den), np.cos(k / den))
    return tab.reshape(rows, cols, dim).astype(np.float32)


class ImageEncoderAttention(Layer):
    """call([x4d]) -> (x4d, positional4d)  (reference :244-321).  The returned positional encoding is
    the [rows, cols, D] table itself (not tiled over the batch; DecoderPrep accepts both)."""

    def __init__(self, num_blocks, num_attention_heads, name="ImageEncoderAttention", **kwargs):
        super().__init__(name=name)
        self.num_blocks = num_blocks
        self.num_attention_heads = num_attention_heads
        self.EncoderBlocks = [EncoderBlock(num_attention_heads, name=f"EncoderBlock_{i}") for i in range(num_blocks)]

    def get_config(self):
        return {**super().get_config(), "num_blocks": self.num_blocks, "num_attention_heads": self.num_attention_heads}

    def build(self, input_shape):
        _, R, Cc, D = input_shape[0]
        self.add_weight("positional_encoding", positional_table(R, Cc, D))

    def pos_tc(self):
        """[L,D] positional table as a tensor-core operand: its tf32-rounded shadow when the model keeps one."""
        pos = self._weights["positional_encoding"]
        sh = self._shadow.get("positional_encoding")
        t = sh if (sh is not None and _lib.tc_mode()) else pos
        return t.view(-1, pos.shape[-1])

    def forward(self, inputs, training=False, dropout_keys=None, seed_dev=None, tabs=None):
        x4 = f32(inputs[0])
        self.maybe_build([x4])
        B, R, Cc, D = x4.shape
        pos = self._weights["positional_encoding"]
        x = x4.view(B, R * Cc, D)
        ctxs = []
        for i, blk in enumerate(self.EncoderBlocks):
            keys = (0, 0) if dropout_keys is None else dropout_keys[i]
            x, c = blk.forward([x, pos.view(R * Cc, D)], training, keys, seed_dev, pos_tc=self.pos_tc(),
                               tabs=tabs if (tabs is not None and self.num_blocks == 1) else None)
            ctxs.append(c)
        return (x.view(B, R, Cc, D), pos), {"blocks": ctxs, "shape": (B, R, Cc, D)}

    def backward(self, ctx, d_x4, d_pos_extra=None, d_x=None, acc=False, need_dx=True):
        """d_x4: gradient of the encoder output; d_pos_extra: gradient arriving at the returned table.  Fused path
        (one encoder block, tensor-core mode): the input gradient may be accumulated into an existing buffer `d_x`
        (acc=True) or skipped (need_dx=False)."""
        B, R, Cc, D = ctx["shape"]
        g_pos = self._grads["positional_encoding"].view(R * Cc, D)
        if d_pos_extra is not None:
            accumulate(d_pos_extra.reshape(R * Cc, D), g_pos)
        d = d_x4.view(B, R * Cc, D)
        if len(self.EncoderBlocks) == 1 and ctx["blocks"][0].get("fused"):
            d = self.EncoderBlocks[0].backward(ctx["blocks"][0], d, g_pos, d_x=None if d_x is None else d_x.view(B, R * Cc, D), acc=acc, need_dx=need_dx)
            return None if d is None else d.view(B, R, Cc, D)
        for blk, c in zip(reversed(self.EncoderBlocks), reversed(ctx["blocks"])):
            d = blk.backward(c, d, g_pos)
        return d.view(B, R, Cc, D)


class DecoderPrep(Layer):
    """call([x4d, pos]) -> (encoder_value, decoder_features, encoder_key, decoder_positional)  (reference :397-456)."""

    def __init__(self, num_object_preds, decoder_dim, name="DecoderPrep", **kwargs):
        super().__init__(name=name)
        self.num_object_preds = num_object_preds
        self.decoder_dim = decoder_dim

    def get_config(self):
        return {**super().get_config(), "num_object_preds": self.num_object_preds, "decoder_dim": self.decoder_dim}

    def build(self, input_shape):
        # zeros initialiser, trainable (reference :427-431)
        self.add_weight("init_decoder_features", np.zeros((self.num_object_preds, self.decoder_dim), np.float32))

    def tile_queries(self, B, like=None):
        """The learned queries tiled over the batch (:445-447).  Callable ahead of `forward` (the queries do not
        depend on the image) so that the decoder self-attention can start early."""
        if like is not None:
            self.maybe_build([like])
        q0 = self._weights["init_decoder_features"]
        dec = empty(B, *q0.shape)
        _lib.call("bdetr_tile_queries_fwd", B, q0.shape[0], q0.shape[1], ptr(q0), ptr(dec), stream_ptr())
        return dec

    def forward(self, inputs, training=False, dec=None):
        x4, pos = inputs
        x4 = f32(x4)
        self.maybe_build([x4])
        B, R, Cc, D = x4.shape
        enc_value = x4.view(B, R * Cc, D)
        enc_key = add_positional(enc_value, pos.reshape(-1, D)[: R * Cc])      # Add (:441)
        if dec is None:
            dec = self.tile_queries(B)
        return (enc_value, dec, enc_key, dec), {"shape": (B, R, Cc, D)}

    def backward_queries(self, d_dec):
        """Gradient of the tiled queries -> the shared query parameter (sum over the batch)."""
        batch_sum_into(d_dec, self._grads["init_decoder_features"])

    def backward(self, ctx, d_enc_value, d_dec, d_enc_key, d_pos):
        """Returns d_x4; accumulates into d_pos ([L,D]) and (when d_dec is given) into the query parameter's gradient."""
        B, R, Cc, D = ctx["shape"]
        if d_dec is not None:
            self.backward_queries(d_dec)
        batch_sum_into(d_enc_key, d_pos)
        accumulate(d_enc_key, d_enc_value)
        return d_enc_value.view(B, R, Cc, D)


class DecoderBlock_NoSelfAttention(Layer):
    """cross-attention -> FFN  (reference :324-353)."""

    def __init__(self, num_attention_heads, name="DecoderBlock_NoSelfAttention", **kwargs):
        super().__init__(name=name)
        self.num_attention_heads = num_attention_heads
        self.JointAttentionBlock = AttentionBlock(num_attention_heads, name="JointAttentionBlock")
        self.FeedForwardBlock = FeedForwardBlock(name="FeedForwardBlock")

    def get_config(self):
        return {**super().get_config(), "num_attention_heads": self.num_attention_heads}

    def forward(self, inputs, training=False, dropout_keys=(0, 0, 0), pre_self=None, seed_dev=None):
        enc_value, dec, enc_key, _ = inputs
        a, c1 = self.JointAttentionBlock.forward([dec, enc_key, enc_value], training, dropout_keys[1], seed_dev)
        y, c2 = self.FeedForwardBlock.forward([a], training, dropout_keys[2], seed_dev)
        return y, {"joint": c1, "ffn": c2}

    def backward(self, ctx, d_out, defer_self=False):
        """Returns (d_enc_value, d_dec, d_enc_key)."""
        d_a = self.FeedForwardBlock.backward(ctx["ffn"], d_out)
        d_dec, d_key, d_val = self.JointAttentionBlock.backward(ctx["joint"], d_a)
        return d_val, d_dec, d_key

    def backward_self(self, ctx, d_s):
        return d_s                                # no self-attention in block 0: d_s already is the query gradient

    def forward_fused(self, enc, fold, dec_in, training=False, dropout_keys=(0, 0, 0), seed_dev=None):
        """Tensor-core path: cross-attention with the key's positional add folded into the k projection
        (fold = make_fold(pos, tab_k=pos Wk + bk, pos_tc)), then the FFN block.  dec_in [B,Q,D] = the tiled queries
        (block 0) or the hoisted self-attention block's output."""
        a, c1 = self.JointAttentionBlock.forward_fused(dec_in, enc, fold, training, dropout_keys[1], seed_dev)
        y, c2 = self.FeedForwardBlock.forward([a], training, dropout_keys[2], seed_dev)
        return y, {"joint": c1, "ffn": c2, "fused": True}

    def backward_fused(self, ctx, d_out, d_enc=None, acc_enc=False, d_pos=None, need_d_dec=True, need_d_enc=True):
        """Returns (d_dec_in, d_enc)."""
        d_a = self.FeedForwardBlock.backward(ctx["ffn"], d_out)
        return self.JointAttentionBlock.backward_fused(ctx["joint"], d_a, d_memory=d_enc, acc=(False, acc_enc), d_pos=d_pos,
                                                       need_dq=need_d_dec, need_dm=need_d_enc)


class DecoderBlock(Layer):
    """self-attention (no positional, reference :378-380) -> cross-attention -> FFN  (reference :356-394)."""

    def __init__(self, num_attention_heads, name="DecoderBlock", **kwargs):
        super().__init__(name=name)
        self.num_attention_heads = num_attention_heads
        self.SelfAttentionBlock = AttentionBlock(num_attention_heads, name="SelfAttentionBlock")
        self.JointAttentionBlock = AttentionBlock(num_attention_heads, name="JointAttentionBlock")
        self.FeedForwardBlock = FeedForwardBlock(name="FeedForwardBlock")

    def get_config(self):
        return {**super().get_config(), "num_attention_heads": self.num_attention_heads}

    def forward(self, inputs, training=False, dropout_keys=(0, 0, 0), pre_self=None, seed_dev=None):
        """pre_self: (output, ctx) of SelfAttentionBlock.forward([dec, dec, dec]) when the caller already ran it
        (it depends on the queries only, so the model overlaps it with the encoder block)."""
        enc_value, dec, enc_key, _ = inputs
        s, c0 = pre_self if pre_self is not None else self.SelfAttentionBlock.forward([dec, dec, dec], training, dropout_keys[0], seed_dev)
        a, c1 = self.JointAttentionBlock.forward([s, enc_key, enc_value], training, dropout_keys[1], seed_dev)
        y, c2 = self.FeedForwardBlock.forward([a], training, dropout_keys[2], seed_dev)
        return y, {"self": c0, "joint": c1, "ffn": c2}

    def backward(self, ctx, d_out, defer_self=False):
        """Returns (d_enc_value, d_dec, d_enc_key); with defer_self the middle entry is the gradient of the
        self-attention OUTPUT and the caller finishes with backward_self (on another stream)."""
        d_a = self.FeedForwardBlock.backward(ctx["ffn"], d_out)
        d_s, d_key, d_val = self.JointAttentionBlock.backward(ctx["joint"], d_a)
        if defer_self:
            return d_val, d_s, d_key
        return d_val, self.backward_self(ctx, d_s), d_key

    def backward_self(self, ctx, d_s):
        d_dec, _, _ = self.SelfAttentionBlock.backward(ctx["self"], d_s)     # q = k = v share one buffer
        return d_dec

    forward_fused = DecoderBlock_NoSelfAttention.forward_fused
    backward_fused = DecoderBlock_NoSelfAttention.backward_fused
