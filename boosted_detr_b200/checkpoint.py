"""Keras checkpoint import / export by variable name.

The reference saves `ModelCheckpoint(..., save_weights_only=True)` TensorFlow checkpoints and restores them with
`detection_model.load_weights(tf.train.latest_checkpoint(dir))` (Boosted_DETR_COCO.ipynb cells 19, 26).  TensorFlow is
not installable here, so the exchange format is a plain `.npz` of {name: array}; three spellings of a variable's name
are accepted on import and can be produced on export:

  keras        ImageEncoderAttention_0/EncoderBlock_0/SelfAttentionBlock/AttentionLayer/QueryProjection/kernel
               (the layer-name path: what `model.variables[i].name` shows in the reference, with or without the model
               prefix `DETR/` and the `:0` suffix)
  object_graph EncoderTransformerBlocks/0/EncoderBlocks/0/SelfAttentionBlock/AttentionLayer/QueryProjection/kernel/.ATTRIBUTES/VARIABLE_VALUE
               (the key of a TF2 object-based checkpoint: Python attribute path of the reference's classes,
               boosted_model.py:85-116, transformers.py:41-48,137,174-180,252,288, prediction_heads.py:40-43,106-109,175-178)

A maintainer with TensorFlow exports a reference checkpoint with five lines:
    r = tf.train.load_checkpoint(path); np.savez(out, **{k: r.get_tensor(k) for k, _ in tf.train.list_variables(path)
                                                        if k.endswith('VARIABLE_VALUE') and 'OPTIMIZER_SLOT' not in k})
and `load_weights(model, out)` here takes it as is; `save_weights(model, path, naming='object_graph')` goes the other way.
Shapes are the reference's own ([in, out] Dense kernels, [rows, cols, D] positional tables), so no transposition happens.
"""
from __future__ import annotations

import re

import numpy as np

_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_LIST_OWNERS = {"ImageEncoderAttention": "EncoderTransformerBlocks", "DecoderBlock": "DecoderBlocks",
                "CategoryPredictionHead": "CategoryBlocks", "AttributePredictionHead": "AttributeBlocks",
                "BoxPredictionHead": "BoxBlocks"}


def keras_to_object_graph(name: str) -> str:
    """Layer-name path -> key of a TF2 object-based checkpoint of the reference model."""
    parts = name.split("/")
    m = re.fullmatch(r"([A-Za-z]+)_(\d+)", parts[0])
    if m and m.group(1) in _LIST_OWNERS:
        head = [_LIST_OWNERS[m.group(1)], m.group(2)]
    else:
        head = [parts[0]]                                   # DecoderPrep
    rest = []
    for p in parts[1:]:
        m2 = re.fullmatch(r"EncoderBlock_(\d+)", p)
        rest += ["EncoderBlocks", m2.group(1)] if m2 else [p]
    return "/".join(head + rest) + _SUFFIX


def object_graph_to_keras(key: str) -> str:
    k = key[:-len(_SUFFIX)] if key.endswith(_SUFFIX) else key
    parts = k.split("/")
    inv = {v: n for n, v in _LIST_OWNERS.items()}
    out, i = [], 0
    if parts[0] in inv and len(parts) > 1 and parts[1].isdigit():
        out.append(f"{inv[parts[0]]}_{parts[1]}")
        i = 2
    while i < len(parts):
        if parts[i] == "EncoderBlocks" and i + 1 < len(parts) and parts[i + 1].isdigit():
            out.append(f"EncoderBlock_{parts[i + 1]}")
            i += 2
        else:
            out.append(parts[i])
            i += 1
    return "/".join(out)


def normalise_name(name: str, model_name: str = "DETR") -> str:
    """Any accepted spelling -> this package's variable name."""
    n = name
    if n.endswith(_SUFFIX) or n.split("/")[0] in _LIST_OWNERS.values():
        n = object_graph_to_keras(n)
    if n.endswith(":0"):
        n = n[:-2]
    if n.startswith(model_name + "/"):
        n = n[len(model_name) + 1:]
    return n


def save_weights(model, path: str, naming: str = "keras") -> str:
    """Writes every variable of the path (trainable + BatchNorm moving statistics) to `path` (.npz)."""
    if naming not in ("keras", "object_graph"):
        raise ValueError("naming must be 'keras' or 'object_graph'")
    w = model.get_weights_dict()
    out = {(keras_to_object_graph(n) if naming == "object_graph" else n): a for n, a in w.items()}
    if not path.endswith(".npz"):
        path += ".npz"
    np.savez(path, **out)
    return path


def load_weights(model, source, strict: bool = True) -> list[str]:
    """`source`: path of an .npz or a {name: array} mapping in any accepted spelling.  Returns the names that were set.
    strict: every variable of the model must be present with the reference's shape; unknown entries (optimizer slots,
    backbone variables of a full reference checkpoint) are ignored."""
    data = dict(np.load(source)) if isinstance(source, str) else dict(source)
    have = {normalise_name(k, model.name): np.asarray(v) for k, v in data.items()}
    want = {n: tuple(o._weights[k].shape) for n, o, k in model.named_weights()}
    missing = [n for n in want if n not in have]
    if strict and missing:
        raise KeyError(f"{len(missing)} variables missing from the checkpoint, e.g. {missing[:3]}")
    bad = [(n, have[n].shape, want[n]) for n in want if n in have and tuple(have[n].shape) != want[n]]
    if bad:
        raise ValueError(f"shape mismatch for {bad[0][0]}: checkpoint {bad[0][1]} vs model {bad[0][2]}")
    model.set_weights_dict({n: have[n] for n in want if n in have})
    return [n for n in want if n in have]
