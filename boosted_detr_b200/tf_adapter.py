"""Optional TensorFlow adapter ("optional shim only if TF is importable").  TensorFlow cannot be installed in
this image, so this module has never been executed here; it shows the binding a TF-equipped maintainer would use and
raises ImportError with a clear message otherwise.

Each adapter is a `tf.keras.layers.Layer` whose `call` hands DLPack views of its (GPU-resident) inputs to the matching
layer of this package and wraps the result back -- no host copy.  Gradients: `tf.custom_gradient` around the package's
explicit forward / backward pair."""
from __future__ import annotations

try:  # pragma: no cover - TensorFlow is absent from the build image
    import tensorflow as tf
except ImportError as _e:  # pragma: no cover
    tf = None
    _IMPORT_ERROR = _e


def _require_tf():
    if tf is None:
        raise ImportError("tensorflow is not installed: boosted_detr_b200.tf_adapter needs it "
                          f"(the package itself runs without TensorFlow); original error: {_IMPORT_ERROR}")


def to_torch(x):  # pragma: no cover
    import torch
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(x))


def to_tf(t):  # pragma: no cover
    import torch
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


def keras_layer(layer):  # pragma: no cover
    """Wraps a boosted_detr_b200 layer (AttentionBlock, FeedForwardBlock, EncoderBlock, heads, MatchingLoss, ...) as a
    tf.keras.layers.Layer with the reference's call convention (lists of tensors)."""
    _require_tf()

    class _Adapter(tf.keras.layers.Layer):
        def __init__(self):
            super().__init__(name=layer.name)
            self.inner = layer

        def call(self, inputs, training=False):
            @tf.custom_gradient
            def op(*xs):
                out, ctx = self.inner.forward([to_torch(x) for x in xs], training=bool(training))
                outs = out if isinstance(out, (list, tuple)) else [out]

                def grad(*dys):
                    g = self.inner.backward(ctx, *[to_torch(d) for d in dys])
                    g = g if isinstance(g, (list, tuple)) else [g]
                    return [None if t is None else to_tf(t) for t in g]
                return [to_tf(o) for o in outs], grad
            return op(*inputs)

    return _Adapter()
