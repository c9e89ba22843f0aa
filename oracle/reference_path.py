"""CPU ORACLE (test infrastructure, NOT product code) for the Boosted_DETR per-step hot path.

PARITY UNPINNED at every TensorFlow / Keras / tensorflow_addons boundary: the reference
has no tests, golden vectors or fixtures, and TF cannot be installed in this image, so
the TF-side formulas below are restated from the published behaviour of those libraries
(Keras Dense / LayerNormalization / BatchNormalization / BinaryCrossentropy, tfa
giou_loss and SigmoidFocalCrossEntropy).  PINNED: the assignment step calls the very
function the reference calls, scipy.optimize.linear_sum_assignment
(/root/reference/ModelComponents/losses_and_metrics.py:242).  CROSS-CHECKED (tests/test_oracle_crosscheck.py):
GIoU / IoU, focal loss, LayerNorm / BatchNorm / attention core, SGD-Nesterov + clipnorm and the cosine-restart
schedule agree with the independent torchvision / torch implementations of the same published algorithms.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product (boosted_detr_b200/) never does.

Everything is written with torch CPU ops (float32 or float64) so that torch autograd
supplies the reference gradients ("tape.gradient of the summed loss vector").
All citations are file:line under /root/reference/ModelComponents/.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

# losses_and_metrics.py:8-11
DEFAULT_CATEGORY_WEIGHT = 1000.0
DEFAULT_BOX_WEIGHT = 1.0
DEFAULT_ATTRIBUTE_WEIGHT = 100.0
DEFAULT_EXIST_WEIGHT = 100.0

KERAS_EPS = 1e-7          # tf.keras.backend.epsilon()
LN_EPS = 1e-3             # transformers.py:137 (explicit) and Keras default (:180)
BN_EPS = 1e-3             # Keras BatchNormalization default epsilon
BN_MOMENTUM = 0.99        # Keras BatchNormalization default momentum
DROPOUT_RATE = 0.1        # transformers.py:135,179


# ----------------------------------------------------------------------------------------
# Dropout masks.  TF's RNG stream cannot be reproduced, so the oracle takes the keep-mask
# from a counter-based hash that the CUDA kernels implement identically (DESIGN.md
# "dropout").  keep(idx) = lowbias32(idx ^ key) >= rate * 2^32.
# ----------------------------------------------------------------------------------------
def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def dropout_key(seed: int, site: int) -> int:
    a = int(_lowbias32(np.array([(site + 0x9E3779B9) & 0xFFFFFFFF]))[0])
    return int(_lowbias32(np.array([(seed ^ a) & 0xFFFFFFFF]))[0])


def dropout_keep_mask(n: int, seed: int, site: int, rate: float = DROPOUT_RATE) -> np.ndarray:
    idx = np.arange(n, dtype=np.uint64)
    h = _lowbias32(idx ^ np.uint64(dropout_key(seed, site)))
    thresh = np.uint64(int(float(np.float32(rate)) * 4294967296.0))   # the kernels take `rate` as fp32
    return h >= thresh


class Dropout:
    """training-mode dropout with hash masks; `seed=None` disables it (rate 0)."""

    def __init__(self, seed=None, rate=DROPOUT_RATE):
        self.seed, self.rate = seed, rate

    def __call__(self, x: torch.Tensor, site: int, training: bool) -> torch.Tensor:
        if not training or self.seed is None or self.rate == 0.0:
            return x
        keep = dropout_keep_mask(x.numel(), self.seed, site, self.rate).reshape(tuple(x.shape))
        scale = np.float32(1.0) / (np.float32(1.0) - np.float32(self.rate))
        return x * torch.from_numpy(keep).to(x.dtype) * float(scale)


# dropout site ids (one per Keras Dropout layer instance), block i -> 8*i + k
SITE_ENC_ATTN, SITE_ENC_FFN, SITE_DEC_SELF, SITE_DEC_CROSS, SITE_DEC_FFN = 0, 1, 2, 3, 4


def site(block: int, which: int) -> int:
    return 8 * block + which


# ----------------------------------------------------------------------------------------
# Keras primitives
# ----------------------------------------------------------------------------------------
def dense(x, p, prefix):
    """tf.keras.layers.Dense: x @ kernel + bias, kernel is [in, out]."""
    return x @ p[prefix + "/kernel"] + p[prefix + "/bias"]


def layer_norm(x, p, prefix, eps=LN_EPS):
    """tf.keras.layers.LayerNormalization over the last axis, biased variance."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * p[prefix + "/gamma"] + p[prefix + "/beta"]


def batch_norm(x, p, prefix, training, new_stats=None):
    """tf.keras.layers.BatchNormalization on [B,Q,F] (non-fused path): batch statistics over
    all axes but the last, biased variance, eps 1e-3; moving stats updated with momentum .99."""
    gamma, beta = p[prefix + "/gamma"], p[prefix + "/beta"]
    if training:
        flat = x.reshape(-1, x.shape[-1])
        mean = flat.mean(dim=0)
        var = ((flat - mean) ** 2).mean(dim=0)
        if new_stats is not None:
            mm, mv = p[prefix + "/moving_mean"], p[prefix + "/moving_variance"]
            new_stats[prefix + "/moving_mean"] = (mm * BN_MOMENTUM + mean * (1 - BN_MOMENTUM)).detach()
            new_stats[prefix + "/moving_variance"] = (mv * BN_MOMENTUM + var * (1 - BN_MOMENTUM)).detach()
    else:
        mean, var = p[prefix + "/moving_mean"], p[prefix + "/moving_variance"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta


def backbone_neck(x4d, p, prefix="BackboneNeck", training=False, new_stats=None):
    """BackboneNeck.call, backbone.py:90-95: batch_norm1 -> Conv2D(1x1, tanh) -> batch_norm2 on a channels-last map
    (a 1x1 convolution is a Dense over the pixels; kernel [1,1,Cin,N])."""
    x = batch_norm(x4d, p, prefix + "/batch_norm1", training, new_stats)
    w = p[prefix + "/conv2d_downscaler/kernel"]
    x = torch.tanh(x @ w.reshape(w.shape[-2], w.shape[-1]) + p[prefix + "/conv2d_downscaler/bias"])
    return batch_norm(x, p, prefix + "/batch_norm2", training, new_stats)


# ----------------------------------------------------------------------------------------
# transformers.py
# ----------------------------------------------------------------------------------------
ATTENTION_QUERY_CHUNK = 0      # bench.py's config-5 CPU leg sets this (query rows per slice); 0 = the reference's one-shot form


def multihead_attention(query, key, value, p, prefix, num_heads):
    """MultiheadAttention.call, transformers.py:68-102 (mask is always ones on this path)."""
    B, Lq, _ = query.shape
    Lk = key.shape[1]
    q = dense(query, p, prefix + "/QueryProjection")            # :72
    k = dense(key, p, prefix + "/KeyProjection")                # :73
    v = dense(value, p, prefix + "/ValueProjection")            # :74
    hd = q.shape[-1]
    d = hd // num_heads
    q = q.reshape(B, Lq, num_heads, d).permute(0, 2, 1, 3)      # :77,82  [B,H,Lq,d]
    k = k.reshape(B, Lk, num_heads, d).permute(0, 2, 3, 1)      # :78,83  [B,H,d,Lk]
    v = v.reshape(B, Lk, num_heads, d).permute(0, 2, 1, 3)      # :79,84  [B,H,Lk,d]
    if ATTENTION_QUERY_CHUNK and Lq > ATTENTION_QUERY_CHUNK:
        # identical arithmetic row by row (every softmax row still sees all Lk keys); only bounds the host memory the
        # [B,H,Lq,Lk] score tensor needs at BASELINE config 5's 20 020-token sequences (51 GB otherwise)
        x = torch.cat([torch.softmax((q[:, :, i:i + ATTENTION_QUERY_CHUNK] @ k) * (1.0 / math.sqrt(float(d))), dim=-1) @ v
                       for i in range(0, Lq, ATTENTION_QUERY_CHUNK)], dim=2)
    else:
        x = q @ k                                               # :87
        x = x * (1.0 / math.sqrt(float(d)))                     # :88 Rescaling(scale=1/sqrt(dim))
        x = torch.softmax(x, dim=-1)                            # :89
        x = x @ v                                               # :97  [B,H,Lq,d]
    x = x.contiguous().reshape(B, Lq, hd)                       # :100 raw reshape, NO permute back (Q1)
    return dense(x, p, prefix + "/OutputProjection")            # :101


def attention_block(query, key, value, p, prefix, num_heads, drop, site_id, training):
    """AttentionBlock.call, transformers.py:139-151: LN(query + Dropout(MHA))."""
    a = multihead_attention(query, key, value, p, prefix + "/AttentionLayer", num_heads)
    a = drop(a, site_id, training)
    return layer_norm(query + a, p, prefix + "/LayerNorm")


def feed_forward_block(x, p, prefix, drop, site_id, training):
    """FeedForwardBlock.call, transformers.py:182-193 (hidden width = feature dim)."""
    h = torch.relu(dense(x, p, prefix + "/DenseRelu"))
    h = dense(h, p, prefix + "/DenseLinear")
    h = drop(h, site_id, training)
    return layer_norm(x + h, p, prefix + "/LayerNorm")


def encoder_block(x, pos, p, prefix, num_heads, drop, block, training):
    """EncoderBlock.call, transformers.py:222-235: q = k = x+pos, v = x (residual on x+pos, Q4)."""
    q = x + pos
    k = x + pos
    x = attention_block(q, k, x, p, prefix + "/SelfAttentionBlock", num_heads, drop,
                        site(block, SITE_ENC_ATTN), training)
    return feed_forward_block(x, p, prefix + "/FeedForwardBlock", drop,
                              site(block, SITE_ENC_FFN), training)


def positional_table(rows, cols, dim, dtype=np.float64):
    """ImageEncoderAttention.build, transformers.py:282-291: value depends on the parity of the
    flattened POSITION k (odd -> sin, even -> cos) and on den = 2(1+dim)/D."""
    k = np.arange(rows * cols, dtype=np.float64)[:, None]
    den = 2.0 * (1.0 + np.arange(dim, dtype=np.float64))[None, :] / dim
    tab = np.where((k % 2) == 1, np.sin(k / den), np.cos(k / den))
    return tab.reshape(rows, cols, dim).astype(dtype)


def image_encoder_attention(x4d, p, prefix, num_heads, drop, block, training):
    """ImageEncoderAttention.call (num_blocks=1), transformers.py:294-315."""
    B, R, Cc, D = x4d.shape
    pos = p[prefix + "/positional_encoding"].reshape(1, R * Cc, D).expand(B, R * Cc, D)
    x = x4d.reshape(B, R * Cc, D)
    x = encoder_block(x, pos, p, prefix + "/EncoderBlock_0", num_heads, drop, block, training)
    return x.reshape(B, R, Cc, D), pos.reshape(B, R, Cc, D)


def decoder_prep(x4d, pos4d, p, prefix="DecoderPrep"):
    """DecoderPrep.call, transformers.py:433-450."""
    B, R, Cc, D = x4d.shape
    enc_value = x4d.reshape(B, R * Cc, D)
    enc_key = enc_value + pos4d.reshape(B, R * Cc, D)           # :441
    q0 = p[prefix + "/init_decoder_features"]
    dec = q0.unsqueeze(0).expand(B, *q0.shape)                  # :445-447
    return enc_value, dec, enc_key, dec


def decoder_block(enc_value, dec, enc_key, p, prefix, num_heads, drop, block, training, self_attention):
    """DecoderBlock_NoSelfAttention.call :340-353 (block 0) / DecoderBlock.call :374-394 (>=1)."""
    if self_attention:
        dec = attention_block(dec, dec, dec, p, prefix + "/SelfAttentionBlock", num_heads, drop,
                              site(block, SITE_DEC_SELF), training)
    dec = attention_block(dec, enc_key, enc_value, p, prefix + "/JointAttentionBlock", num_heads, drop,
                          site(block, SITE_DEC_CROSS), training)
    return feed_forward_block(dec, p, prefix + "/FeedForwardBlock", drop,
                              site(block, SITE_DEC_FFN), training)


# ----------------------------------------------------------------------------------------
# prediction_heads.py  (the Conv1D/Permute branch is dead: num_preds == Q always)
# ----------------------------------------------------------------------------------------
def category_head(dec, p, prefix, training, new_stats=None):
    """SingleClassPredictionHead.call, prediction_heads.py:113-131."""
    h = torch.relu(dense(dec, p, prefix + "/DenseCateg"))
    h = batch_norm(h, p, prefix + "/BatchNorm", training, new_stats)
    return torch.softmax(dense(h, p, prefix + "/DenseLogits"), dim=-1)


def attribute_head(dec, p, prefix, training, new_stats=None):
    """MultiClassPredictionHead.call, prediction_heads.py:182-201."""
    h = torch.relu(dense(dec, p, prefix + "/Dense"))
    h = batch_norm(h, p, prefix + "/BatchNorm", training, new_stats)
    return torch.sigmoid(dense(h, p, prefix + "/DenseLinear"))


def box_head(dec, p, prefix, training, new_stats=None):
    """BoxPredictionHead.call, prediction_heads.py:46-63: 3*sigmoid(x/100) - 1."""
    h = torch.relu(dense(dec, p, prefix + "/Dense"))
    h = batch_norm(h, p, prefix + "/BatchNorm", training, new_stats)
    return 3.0 * torch.sigmoid(dense(h, p, prefix + "/BoxCoords") / 100.0) - 1.0


# ----------------------------------------------------------------------------------------
# losses_and_metrics.py
# ----------------------------------------------------------------------------------------
def safe_clip(x):
    """:26-27.  Gradient passes where .001 <= x <= .999 (tf.clip_by_value)."""
    return torch.clamp(x, 0.001, 0.999)


def keras_bce_elem(y, p):
    """tf.keras.backend.binary_crossentropy (probabilities): clip to [eps,1-eps], eps inside logs."""
    p = torch.clamp(p, KERAS_EPS, 1.0 - KERAS_EPS)
    return -(y * torch.log(p + KERAS_EPS) + (1.0 - y) * torch.log(1.0 - p + KERAS_EPS))


def keras_bce(y, p):
    """tf.keras.losses.BinaryCrossentropy(reduction=NONE): mean over the last axis."""
    return keras_bce_elem(y, p).mean(dim=-1)


def category_loss(y_true, y_pred):
    """CategoryLoss :44-49: BCE(y_true, clip(y_pred) * y_true)."""
    return keras_bce(y_true, safe_clip(y_pred) * y_true)


def attribute_loss(y_true, y_pred):
    """AttributeLoss :51-57: tfa SigmoidFocalCrossEntropy(alpha .25, gamma 2) on a trailing
    singleton axis (its reduce_sum is over that axis), then mean over attributes."""
    y_true = y_true.unsqueeze(-1)
    p = safe_clip(y_pred).unsqueeze(-1)
    ce = keras_bce_elem(y_true, p)
    p_t = y_true * p + (1.0 - y_true) * (1.0 - p)
    alpha_factor = y_true * 0.25 + (1.0 - y_true) * 0.75
    mod = (1.0 - p_t) ** 2.0
    focal = (alpha_factor * mod * ce).sum(dim=-1)
    return focal.mean(dim=-1)


def coco_to_tf(box):
    """:59-66  [x,y,w,h] -> [ymin, xmin, ymin+h, xmin+w]."""
    xmin, ymin, w, h = box[..., 0:1], box[..., 1:2], box[..., 2:3], box[..., 3:4]
    return torch.cat([ymin, xmin, ymin + h, xmin + w], dim=-1)


def _tf_max(a, b):
    """tf.maximum with TF's gradient convention (ties go to the first argument)."""
    a, b = torch.broadcast_tensors(a, b)
    return torch.where(a >= b, a, b)


def _tf_min(a, b):
    a, b = torch.broadcast_tensors(a, b)
    return torch.where(a <= b, a, b)


def _div_no_nan(x, y):
    safe = torch.where(y == 0, torch.ones_like(y), y)
    return torch.where(y == 0, torch.zeros_like(y), x / safe)


def tfa_giou(b1, b2, mode="giou"):
    """tensorflow_addons.losses.giou_loss._calculate_giou, boxes [ymin,xmin,ymax,xmax]."""
    zero = torch.zeros((), dtype=b1.dtype)
    b1_ymin, b1_xmin, b1_ymax, b1_xmax = b1.unbind(-1)
    b2_ymin, b2_xmin, b2_ymax, b2_xmax = b2.unbind(-1)
    b1_w = _tf_max(zero, b1_xmax - b1_xmin)
    b1_h = _tf_max(zero, b1_ymax - b1_ymin)
    b2_w = _tf_max(zero, b2_xmax - b2_xmin)
    b2_h = _tf_max(zero, b2_ymax - b2_ymin)
    b1_area = b1_w * b1_h
    b2_area = b2_w * b2_h
    i_ymin = _tf_max(b1_ymin, b2_ymin)
    i_xmin = _tf_max(b1_xmin, b2_xmin)
    i_ymax = _tf_min(b1_ymax, b2_ymax)
    i_xmax = _tf_min(b1_xmax, b2_xmax)
    i_w = _tf_max(zero, i_xmax - i_xmin)
    i_h = _tf_max(zero, i_ymax - i_ymin)
    i_area = i_w * i_h
    union = b1_area + b2_area - i_area
    iou = _div_no_nan(i_area, union)
    if mode == "iou":
        return iou
    e_ymin = _tf_min(b1_ymin, b2_ymin)
    e_xmin = _tf_min(b1_xmin, b2_xmin)
    e_ymax = _tf_max(b1_ymax, b2_ymax)
    e_xmax = _tf_max(b1_xmax, b2_xmax)
    e_w = _tf_max(zero, e_xmax - e_xmin)
    e_h = _tf_max(zero, e_ymax - e_ymin)
    e_area = e_w * e_h
    return iou - _div_no_nan(e_area - union, e_area)


def box_loss(y_true, y_pred, giou_weight=2.0, l2_weight=5.0):
    """BoxLoss :68-72 (intended [B,T,Q] semantics; see SURVEY Q10 for tfa's squeeze)."""
    t, q = coco_to_tf(y_true), coco_to_tf(y_pred)
    giou_l = 1.0 - tfa_giou(t, q, "giou")
    l2 = ((10.0 * t - 10.0 * q) ** 2).mean(dim=-1)
    return giou_weight * giou_l + l2_weight * l2


def iou_metric(y_true, y_pred):
    """IOU_Metric :17-18 = 1 - (1 - iou)."""
    return 1.0 - (1.0 - tfa_giou(coco_to_tf(y_true), coco_to_tf(y_pred), "iou"))


def cost_array(y_true, y_pred, func):
    """CostArray.call :222-225: targets on axis -3 (rows), predictions on axis -2 (columns)."""
    return func(y_true.unsqueeze(-2), y_pred.unsqueeze(-3))


def matching_assignment(cost: np.ndarray, num_objects: np.ndarray) -> np.ndarray:
    """MatchingAssignment.scipy_linear_assignment_mask :234-245 (the reference's literal loop)."""
    masks = np.zeros_like(cost)
    n = np.asarray(num_objects).reshape(-1, 1)
    for i in range(cost.shape[0]):
        ni = int(n[i][0])
        rows, cols = linear_sum_assignment(cost[i, :ni, :])
        masks[i][rows, cols] = 1.0
    return masks


def weighted_cost(y_true, y_pred, weights):
    """MatchingLoss.call :119-130 up to the matrix handed to the matcher."""
    category, attribute, bbox, _ = y_true
    cat_preds, attr_preds, box_preds = y_pred
    w_cat, w_box, w_attr, _ = weights
    cat_c = w_cat * cost_array(category, cat_preds, category_loss)
    attr_c = w_attr * cost_array(attribute, attr_preds, attribute_loss)
    box_c = w_box * cost_array(bbox, box_preds, box_loss)
    return cat_c, attr_c, box_c, cat_c + box_c + attr_c          # :130 summation order


def matching_loss(y_true, y_pred, weights, mask=None):
    """MatchingLoss.call :111-161.  Returns (losses[5] each [B], iou metric [1,Q], mask, total cost).
    The mask is a constant for the gradient (tf.numpy_function has none)."""
    category, attribute, bbox, num_objects = y_true
    cat_preds, attr_preds, box_preds = y_pred
    w_exist = weights[3]
    cat_c, attr_c, box_c, total_c = weighted_cost(y_true, y_pred, weights)
    if mask is None:
        mask_np = matching_assignment(total_c.detach().to(torch.float32).numpy(),
                                      np.asarray(num_objects))
        mask = torch.from_numpy(mask_np).to(total_c.dtype)
    assigned = mask.max(dim=-2).values.unsqueeze(-1)             # :207-208
    cat_c, attr_c, box_c = mask * cat_c, mask * attr_c, mask * box_c
    exist = w_exist * keras_bce(1.0 - assigned, safe_clip(cat_preds[..., 0:1]))   # :139-140 [B,Q]
    total_n = 1.0 + float(np.asarray(num_objects).sum())          # :144
    n_preds = 1.0 + float(cat_preds.shape[1])                     # :145
    cat_l = cat_c.sum(dim=(-2, -1)) / total_n
    attr_l = attr_c.sum(dim=(-2, -1)) / total_n
    box_l = box_c.sum(dim=(-2, -1)) / total_n
    exist_l = exist.mean(dim=-1) / n_preds
    total = cat_l + attr_l + box_l + exist_l                      # :153
    miou = mask * cost_array(bbox, box_preds, iou_metric)         # MatchingMetric :187-188
    miou = miou.unsqueeze(0).sum(dim=(1, 2)) / total_n            # :158 list -> [1,B,T,Q], axes [1,2] (Q7)
    return [total, cat_l, attr_l, box_l, exist_l], miou, mask, total_c


# ----------------------------------------------------------------------------------------
# boosted_model.py
# ----------------------------------------------------------------------------------------
def model_weights(attribute_weight=1.0, classification_only=False):
    """BoostedDETR.__init__ :40-50,147-152 -> (w_cat, w_box, w_attr, w_exist)."""
    w_attr = DEFAULT_ATTRIBUTE_WEIGHT if attribute_weight is None else float(attribute_weight)
    w_box = 0.0 if classification_only else DEFAULT_BOX_WEIGHT
    return (DEFAULT_CATEGORY_WEIGHT, w_box, w_attr, DEFAULT_EXIST_WEIGHT)


def boosted_detr_call(p, features, targets, num_blocks, num_heads, training,
                      drop=None, weights=None, new_stats=None, forced_masks=None, frozen_blocks=()):
    """BoostedDETR.call :170-267 starting at the BackboneNeck output `features` [B,R,Cc,D].

    targets = (category one-hot [B,T,C], attribute multi-hot [B,T,A], bbox [B,T,4], num_objects [B])
    Returns dict(preds=[cat,attr,box], loss=[B] (sum over blocks), metrics{...}, masks=[per block]).
    """
    drop = drop or Dropout(None)
    weights = weights or model_weights()
    x = features
    if "BackboneNeck/conv2d_downscaler/kernel" in p and x.shape[-1] == p["BackboneNeck/conv2d_downscaler/kernel"].shape[-2]:
        x = backbone_neck(x, p, "BackboneNeck", training, new_stats)      # boosted_model.py:195-196
    B, R, Cc, D = x.shape
    loss = cat_l = att_l = box_l = exist_l = 0.0
    iou = None
    masks, costs, per_block = [], [], []
    cat_preds = attr_preds = box_preds = None
    for i in range(num_blocks):                                   # :199
        x = x.reshape(B, R, Cc, D)                                # :204
        x, pos = image_encoder_attention(x, p, f"ImageEncoderAttention_{i}", num_heads, drop, i, training)
        enc_value, dec, enc_key, _ = decoder_prep(x, pos, p)      # :210
        dec = decoder_block(enc_value, dec, enc_key, p, f"DecoderBlock_{i}", num_heads, drop, i,
                            training, self_attention=(i >= 1))    # :213
        # Keras: BatchNormalization inside a layer with trainable = False (the freezing schedule of
        # Boosted_DETR_COCO.ipynb cell 30) runs in inference mode -- moving statistics, no update -- even when the
        # model is called with training=True.  Dropout of frozen blocks still follows `training`.
        bn_training = training and i not in frozen_blocks
        cat_i = category_head(dec, p, f"CategoryPredictionHead_{i}", bn_training, new_stats)
        attr_i = attribute_head(dec, p, f"AttributePredictionHead_{i}", bn_training, new_stats)
        box_i = box_head(dec, p, f"BoxPredictionHead_{i}", bn_training, new_stats)
        if i == 0:                                                # :222-225
            cat_preds, attr_preds, box_preds = cat_i, attr_i, box_i
        cat_preds = cat_preds + cat_i                             # :227-229 (block 0 twice, Q2)
        attr_preds = attr_preds + attr_i
        box_preds = box_preds + box_i
        per_block.append((cat_preds, attr_preds, box_preds))
        if training:                                              # :232
            fm = None if forced_masks is None else forced_masks[i]
            losses_i, metrics_i, mask_i, cost_i = matching_loss(
                targets, [cat_preds, attr_preds, box_preds], weights, mask=fm)
            loss = loss + losses_i[0]
            cat_l = cat_l + losses_i[1]
            att_l = att_l + losses_i[2]
            box_l = box_l + losses_i[3]
            exist_l = exist_l + losses_i[4]
            iou = metrics_i
            masks.append(mask_i)
            costs.append(cost_i)
    out = {"preds": [cat_preds, attr_preds, box_preds], "per_block_preds": per_block,
           "encoder_features": x}
    if training:
        out.update(loss=loss, metrics={"Category_Loss": cat_l, "Attribute_Loss": att_l,
                                       "Box_Loss": box_l, "Existence_Loss": exist_l, "IOU": iou},
                   masks=masks, costs=costs)
    return out


def image_encoder_attention_n(x4d, p, prefix, num_blocks, num_heads, drop, training):
    """ImageEncoderAttention.call with num_blocks encoder blocks (transformers.py:294-315); block j's dropout sites are 8j, 8j+1."""
    B, R, Cc, D = x4d.shape
    pos = p[prefix + "/positional_encoding"].reshape(1, R * Cc, D).expand(B, R * Cc, D)
    x = x4d.reshape(B, R * Cc, D)
    for j in range(num_blocks):
        x = encoder_block(x, pos, p, f"{prefix}/EncoderBlock_{j}", num_heads, drop, j, training)
    return x.reshape(B, R, Cc, D), pos.reshape(B, R, Cc, D)


def detr_call(p, features, targets, num_encoder_blocks, num_decoder_blocks, num_heads, training, drop=None, weights=None,
              new_stats=None, forced_mask=None):
    """Plain DETR.call, /root/reference/ModelComponents/model.py:153-236, from the BackboneNeck output: one encoder
    stack, a CHAIN of decoder blocks, one set of heads on the last decoder output, the matching loss at the last block."""
    drop = drop or Dropout(None)
    weights = weights or model_weights()
    x = features
    if "BackboneNeck/conv2d_downscaler/kernel" in p and x.shape[-1] == p["BackboneNeck/conv2d_downscaler/kernel"].shape[-2]:
        x = backbone_neck(x, p, "BackboneNeck", training, new_stats)
    x, pos = image_encoder_attention_n(x, p, "ImageEncoderAttention", num_encoder_blocks, num_heads, drop, training)
    enc_value, dec, enc_key, _ = decoder_prep(x, pos, p)
    for i in range(num_decoder_blocks):
        dec = decoder_block(enc_value, dec, enc_key, p, f"DecoderBlock_{i}", num_heads, drop, i, training, self_attention=(i >= 1))
    cat = category_head(dec, p, "CategoryPredictionHead", training, new_stats)
    attr = attribute_head(dec, p, "AttributePredictionHead", training, new_stats)
    box = box_head(dec, p, "BoxPredictionHead", training, new_stats)
    out = {"preds": [cat, attr, box]}
    if training:
        losses, metrics, mask, cost = matching_loss(targets, [cat, attr, box], weights, mask=forced_mask)
        out.update(loss=losses[0], losses=losses, iou=metrics, mask=mask, cost=cost)
    return out


def params_to_torch(params: dict, dtype=torch.float64, requires_grad=False) -> dict:
    out = {}
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad and not k.endswith(("moving_mean", "moving_variance")):
            t.requires_grad_(True)
        out[k] = t
    return out


def train_step_reference(params, features, targets, num_blocks, num_heads, dtype=torch.float64,
                         dropout_seed=None, weights=None, forced_masks=None, frozen_blocks=()):
    """Forward (training=True) + gradient of the summed loss vector (Keras train_step semantics:
    tape.gradient of a [B] vector = gradient of its sum).  Returns (out, grads dict, new BN stats)."""
    p = params_to_torch(params, dtype, requires_grad=True)
    feats = torch.tensor(np.asarray(features), dtype=dtype)
    tg = (torch.tensor(np.asarray(targets[0]), dtype=dtype), torch.tensor(np.asarray(targets[1]), dtype=dtype),
          torch.tensor(np.asarray(targets[2]), dtype=dtype), np.asarray(targets[3]))
    new_stats = {}
    out = boosted_detr_call(p, feats, tg, num_blocks, num_heads, True, Dropout(dropout_seed),
                            weights, new_stats, forced_masks, frozen_blocks)
    out["loss"].sum().backward()
    grads = {k: (v.grad.detach().numpy() if v.grad is not None else np.zeros(v.shape))
             for k, v in p.items() if v.requires_grad}
    return out, grads, {k: v.numpy() for k, v in new_stats.items()}


# ---------------------------------------------------------------------------------------------
# Optimizer of the reference's training runs (/root/reference/Boosted_DETR_COCO.ipynb cells 26, 30):
#   tf.keras.optimizers.SGD(learning_rate=CosineDecayRestarts(.001, 4000, m_mul=.95, alpha=.1),
#                           momentum=.9 | .95, nesterov=True, clipnorm=0.1)
# TensorFlow is not installable here: restated from the published TF 2.x semantics (PARITY UNPINNED), in float64 so
# that the product's float32 arithmetic is checked against something more accurate than itself.
# ---------------------------------------------------------------------------------------------
def cosine_decay_restarts(step, initial_learning_rate, first_decay_steps, t_mul=2.0, m_mul=1.0, alpha=0.0):
    """tf.keras.optimizers.schedules.CosineDecayRestarts.__call__ (SGDR, Loshchilov & Hutter)."""
    completed = step / first_decay_steps
    if t_mul == 1.0:
        i_restart = math.floor(completed)
        completed -= i_restart
    else:
        i_restart = math.floor(math.log(1.0 - completed * (1.0 - t_mul)) / math.log(t_mul))
        sum_r = (1.0 - t_mul ** i_restart) / (1.0 - t_mul)
        completed = (completed - sum_r) / t_mul ** i_restart
    m_fac = m_mul ** i_restart
    cosine_decayed = 0.5 * m_fac * (1.0 + math.cos(math.pi * completed))
    return initial_learning_rate * ((1.0 - alpha) * cosine_decayed + alpha)


def sgd_step_reference(variables: dict, grads: dict, accums: dict, lr, momentum, nesterov, clipnorm, dtype=np.float64):
    """One Keras SGD step over named numpy arrays (updated copies returned).  clipnorm is PER VARIABLE
    (tf.clip_by_norm: g * c / max(||g||, c)); momentum form of ResourceApplyKerasMomentum.  dtype float64 for parity
    checks, float32 (TensorFlow's own arithmetic) when the step is being timed."""
    new_v, new_a = {}, {}
    lr, momentum = dtype(lr), dtype(momentum)
    for k, w in variables.items():
        g = np.asarray(grads[k], dtype)
        if clipnorm is not None:
            g = g * dtype(clipnorm / max(float(np.sqrt(np.dot(g.ravel(), g.ravel()))), clipnorm))
        a = np.asarray(accums[k], dtype) * momentum - lr * g
        new_a[k] = a
        new_v[k] = np.asarray(w, dtype) + (a * momentum - lr * g if nesterov else a)
    return new_v, new_a
