"""Plain `DETR` (the reference's ModelComponents/model.py:17-250) on the same layers as BoostedDETR:
ONE ImageEncoderAttention with `num_encoder_blocks` encoder blocks, a CHAIN of decoder blocks (block i's output is block
i+1's decoder input, model.py:180-184), one set of prediction heads (hidden width 4 x decoder_dim for category /
attribute, decoder_dim for boxes, :104-120) applied to the last decoder output, the Hungarian matching loss at the last
block only (`use_intermediate_losses = False`, :179,186).  Inputs as for BoostedDETR: the BackboneNeck output
`inputs['features']` (or `inputs['backbone_features']` with backbone_neck=True) and tokenised targets.  Sequential
schedule (no multi-stream tricks): this model is the compatibility surface, BoostedDETR is the optimised hot path."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .boosted_model import BACKBONE_STRIDE, BoostedDETR
from .device import zeros
from .layers import Layer
from .losses_and_metrics import MatchingLoss
from .prediction_heads import BoxPredictionHead, MultiClassPredictionHead, SingleClassPredictionHead
from .transformers import DecoderBlock, DecoderBlock_NoSelfAttention, DecoderPrep, ImageEncoderAttention, accumulate


class DETR(BoostedDETR):
    def __init__(self, num_object_preds, image_size, num_encoder_blocks, num_encoder_heads, encoder_dim,
                 num_decoder_blocks, num_decoder_heads, decoder_dim, num_panoptic_heads=1, panoptic_dim=32,
                 vocab_dict=None, classification_only=False, attribute_weight=1.0, name="DETR", feature_shape=None, seed=0,
                 backbone_neck=False, backbone_channels=1792, **kwargs):
        # attributes / tokenizer sizes / neck exactly as BoostedDETR; the layer lists are replaced below
        super().__init__(num_object_preds, image_size, num_encoder_blocks, num_encoder_heads, encoder_dim, 0, num_decoder_heads,
                         decoder_dim, num_panoptic_heads, panoptic_dim, vocab_dict, classification_only, attribute_weight, name,
                         feature_shape, seed, backbone_neck, backbone_channels)
        self.num_decoder_blocks = num_decoder_blocks
        self.ImageEncoderAttention = ImageEncoderAttention(num_encoder_blocks, num_encoder_heads, name="ImageEncoderAttention")
        self.DecoderPrep = DecoderPrep(num_object_preds, decoder_dim, name="DecoderPrep")
        self.DecoderBlocks = [DecoderBlock_NoSelfAttention(num_decoder_heads, name="DecoderBlock_0")]
        self.DecoderBlocks += [DecoderBlock(num_decoder_heads, name=f"DecoderBlock_{i}") for i in range(1, num_decoder_blocks)]
        self.CategoryPredictionHead = SingleClassPredictionHead(self.num_categories, 4 * decoder_dim, num_object_preds, name="CategoryPredictionHead")
        self.AttributePredictionHead = MultiClassPredictionHead(self.num_attributes, 4 * decoder_dim, num_object_preds, name="AttributePredictionHead")
        self.BoxPredictionHead = BoxPredictionHead(decoder_dim, num_object_preds, name="BoxPredictionHead")
        self.EncoderTransformerBlocks, self.CategoryBlocks, self.AttributeBlocks, self.BoxBlocks = [], [], [], []

    def layers(self):
        neck = [self.BackboneNeck] if self.BackboneNeck is not None else []
        return [*neck, self.ImageEncoderAttention, self.DecoderPrep, *self.DecoderBlocks, self.CategoryPredictionHead,
                self.AttributePredictionHead, self.BoxPredictionHead]

    def _flatten(self):
        """One flat weight / gradient buffer (no per-block buckets: this model has no boosted block structure)."""
        named = [(n, o, k) for n, o, k in self.named_weights() if k not in o._non_trainable]
        self._index, off = {}, 0
        for n, o, k in named:
            w = o._weights[k]
            self._index[n] = (off, w.numel(), tuple(w.shape))
            off += (w.numel() + 3) // 4 * 4
        self._buckets = [(0, 0, off)]
        flat_w, flat_g, flat_tc = zeros(off), zeros(off), zeros(off)
        for n, o, k in named:
            o0, cnt, shp = self._index[n]
            flat_w[o0:o0 + cnt].copy_(o._weights[k].reshape(-1))
            o._weights[k] = flat_w[o0:o0 + cnt].view(shp)
            o._grads[k] = flat_g[o0:o0 + cnt].view(shp)
            if k.endswith("/kernel") or k in ("positional_encoding", "init_decoder_features"):
                o._shadow[k] = flat_tc[o0:o0 + cnt].view(shp)
        self._flat = (flat_w, flat_g)
        self._flat_tc = flat_tc
        for layer in self.layers():
            layer.invalidate()

    def forward(self, feats, y_true, training):
        use_dropout = training and self.dropout_seed is not None
        if use_dropout and not torch.cuda.is_current_stream_capturing():
            self.push_dropout_seed()
        seed_dev = self._seed_dev if use_dropout else None
        rate = 0.1 if use_dropout else 0.0
        x = feats
        if self.tensor_core_mode():
            self.refresh_shadow()
            x = torch.empty_like(feats)
            _lib.call("bdetr_round_tf32", feats.numel(), feats.data_ptr(), x.data_ptr(), torch.cuda.current_stream().cuda_stream)
        from .layers import dropout_site_key
        key = (lambda i, s: dropout_site_key(8 * i + s)) if use_dropout else (lambda i, s: 0)
        for j, blk in enumerate(self.ImageEncoderAttention.EncoderBlocks):
            blk.SelfAttentionBlock.rate = blk.FeedForwardBlock.rate = rate
        enc_keys = [(key(j, 0), key(j, 1)) for j in range(self.num_encoder_blocks)]
        (x4, pos), c_enc = self.ImageEncoderAttention.forward([x], training, enc_keys, seed_dev)
        prep, c_prep = self.DecoderPrep.forward([x4, pos], training)
        enc_value, dec, enc_key, _ = prep
        c_decs = []
        for i, blk in enumerate(self.DecoderBlocks):                       # decoder chain (model.py:180-184)
            for nm in ("SelfAttentionBlock", "JointAttentionBlock", "FeedForwardBlock"):
                if hasattr(blk, nm):
                    getattr(blk, nm).rate = rate
            dec, c = blk.forward([enc_value, dec, enc_key, None], training, (key(i, 2), key(i, 3), key(i, 4)), seed_dev=seed_dev)
            c_decs.append(c)
        heads = (self.CategoryPredictionHead, self.AttributePredictionHead, self.BoxPredictionHead)
        preds, c_heads = [], []
        for h in heads:
            _, c = h.forward([dec], training)
            preds.append(c["cum"])
            c_heads.append(c)
        loss_ctxs = [self.loss_fn.forward(y_true, preds)] if training else []
        return preds, {"enc": c_enc, "prep": c_prep, "decs": c_decs, "heads": c_heads, "loss": loss_ctxs, "y_true": y_true}

    def backward(self, ctx, gscale=1.0):
        lc = ctx["loss"][0]
        B, T, Q, C, A = lc["dims"]
        r = [zeros(B, Q, C), zeros(B, Q, A), zeros(B, Q, 4)]
        self.loss_fn.backward(lc, *r, gscale)
        heads = (self.CategoryPredictionHead, self.AttributePredictionHead, self.BoxPredictionHead)
        d_dec = None
        for h, c, g in zip(heads, ctx["heads"], r):
            d = h.backward(c, g)
            if d_dec is None:
                d_dec = d
            else:
                accumulate(d, d_dec)
        d_ev = d_ek = None
        for blk, c in zip(reversed(self.DecoderBlocks), reversed(ctx["decs"])):
            dv, d_dec, dk = blk.backward(c, d_dec)
            if d_ev is None:
                d_ev, d_ek = dv, dk
            else:
                accumulate(dv, d_ev)
                accumulate(dk, d_ek)
        Bf, L, D = d_ev.shape
        g_pos = self.ImageEncoderAttention._grads["positional_encoding"].view(L, D)
        d_x4 = self.DecoderPrep.backward(ctx["prep"], d_ev, d_dec, d_ek, g_pos)
        d_x = self.ImageEncoderAttention.backward(ctx["enc"], d_x4)
        if getattr(self, "_neck_ctx", None) is not None and self.BackboneNeck.trainable:
            self.BackboneNeck.backward(self._neck_ctx, d_x)
        return d_x

    def _collect_metrics(self, ctx):
        lc = ctx["loss"][0]
        tot, iou = lc["losses"], lc["iou"]
        self.metric_means = torch.cat([tot.mean(dim=1), iou.mean().reshape(1)])
        self.status_all = lc["status"].reshape(-1)
        self.losses = [tot[0]]
        self.metric_tensors = {"loss": tot[0], "Category_Loss": tot[1], "Attribute_Loss": tot[2], "Box_Loss": tot[3],
                               "Existence_Loss": tot[4], "IOU": iou.unsqueeze(0)}
        return self.metric_tensors

    def _join_metrics(self):
        pass
