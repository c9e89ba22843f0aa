"""Model hyper-parameters — the kwargs contract of BoostedDETR(**params), mirroring the reference's
ModelComponents/parameters.py:98-178 (vocabularies are represented by their sizes only: the string
tokenizers are outside the hot path)."""
from __future__ import annotations


class ModelParameters:
    def __init__(self, dataset_name="COCO"):
        self._num_object_preds = 96          # reference :103
        self._image_size = (560, 560)        # reference :104
        self._pad = "<PAD>"
        self._oov = "<OOV>"
        self._dataset_name = dataset_name

    def dataset_name(self):
        return self._dataset_name

    def vocab_dict(self, name=None):
        """Placeholder vocabularies with the reference's sizes (COCO: 80 categories, 1 attribute;
        Fashionpedia: 46 categories, 294 attributes; +2 special tokens each -> C=82/A=3, C=48/A=296)."""
        d = {"COCO": {"category": [f"coco_cat_{i}" for i in range(80)], "attribute": ["<none>"]},
             "Fashionpedia": {"category": [f"fp_cat_{i}" for i in range(46)],
                              "attribute": [f"fp_attr_{i}" for i in range(294)]}}
        return d[name] if name else d

    def default_vocab(self):
        return self.vocab_dict(self._dataset_name)

    def default_params(self, value=None):
        parameters = {"image_size": self._image_size, "encoder_dim": 256, "num_encoder_blocks": 4,
                      "num_encoder_heads": 8, "num_decoder_blocks": 4, "num_decoder_heads": 8, "decoder_dim": 256,
                      "num_panoptic_heads": 1, "panoptic_dim": 32, "num_object_preds": self._num_object_preds,
                      "vocab_dict": self.default_vocab(), "pad_value": self._pad, "oov_value": self._oov}
        return parameters if value is None else parameters[value]


def baseline_params(config: int):
    """The concrete BASELINE.json configs of SURVEY.md §8 (sizes for measurement)."""
    mp = ModelParameters("Fashionpedia" if config == 3 else "COCO")
    p = mp.default_params()
    p.pop("pad_value"); p.pop("oov_value")
    p.update(num_object_preds=100, image_size=(640, 640))
    if config == 1:
        p.update(num_encoder_blocks=2, num_decoder_blocks=2)
    elif config in (2, 3):
        p.update(num_encoder_blocks=6, num_decoder_blocks=6)
    elif config == 5:
        p.update(num_encoder_blocks=6, num_decoder_blocks=6, image_size=(1333, 800))
    else:
        raise ValueError("config must be 1, 2, 3 or 5")
    return p
