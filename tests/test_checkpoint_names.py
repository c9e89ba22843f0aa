"""Checkpoint variable-name mapping: layer-name paths <-> TF2 object-graph checkpoint keys of the
reference model (attribute paths of boosted_model.py:85-116, transformers.py, prediction_heads.py).  Host logic only."""
from boosted_detr_b200.checkpoint import keras_to_object_graph, normalise_name, object_graph_to_keras

CASES = [
    ("ImageEncoderAttention_3/EncoderBlock_0/SelfAttentionBlock/AttentionLayer/QueryProjection/kernel",
     "EncoderTransformerBlocks/3/EncoderBlocks/0/SelfAttentionBlock/AttentionLayer/QueryProjection/kernel/.ATTRIBUTES/VARIABLE_VALUE"),
    ("ImageEncoderAttention_0/positional_encoding", "EncoderTransformerBlocks/0/positional_encoding/.ATTRIBUTES/VARIABLE_VALUE"),
    ("DecoderPrep/init_decoder_features", "DecoderPrep/init_decoder_features/.ATTRIBUTES/VARIABLE_VALUE"),
    ("DecoderBlock_2/JointAttentionBlock/LayerNorm/gamma", "DecoderBlocks/2/JointAttentionBlock/LayerNorm/gamma/.ATTRIBUTES/VARIABLE_VALUE"),
    ("DecoderBlock_0/FeedForwardBlock/DenseRelu/bias", "DecoderBlocks/0/FeedForwardBlock/DenseRelu/bias/.ATTRIBUTES/VARIABLE_VALUE"),
    ("CategoryPredictionHead_5/BatchNorm/moving_variance", "CategoryBlocks/5/BatchNorm/moving_variance/.ATTRIBUTES/VARIABLE_VALUE"),
    ("AttributePredictionHead_1/DenseLinear/kernel", "AttributeBlocks/1/DenseLinear/kernel/.ATTRIBUTES/VARIABLE_VALUE"),
    ("BoxPredictionHead_4/BoxCoords/bias", "BoxBlocks/4/BoxCoords/bias/.ATTRIBUTES/VARIABLE_VALUE"),
]


def test_round_trip():
    for keras, og in CASES:
        assert keras_to_object_graph(keras) == og
        assert object_graph_to_keras(og) == keras
        assert normalise_name(og) == keras
        assert normalise_name("DETR/" + keras + ":0") == keras
        assert normalise_name(keras) == keras
