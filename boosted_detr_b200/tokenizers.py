"""Inference tail of the path — the numeric half of the reference's ModelComponents/tokenizers.py `InverseTokenization`
(:91-141): argmax category token, thresholded multi-hot attribute tokens, on the GPU (bdetr_inverse_tokenize); the
token -> string lookup (StringLookup(invert=True), :102-111) is a host-side list index and kept optional.

Vocabulary layout follows the reference's StringLookup(mask_token='<PAD>', oov_token='<OOV>'): index 0 = '<PAD>',
1 = '<OOV>', then the vocabulary (tokenizers.py:32-33, 102-111)."""
from __future__ import annotations

import torch

from . import _lib
from .device import empty, f32, ptr, stream_ptr


class InverseTokenization:
    def __init__(self, vocab_dict, name="Tokenization", **kwargs):
        self.name = name
        self.vocab_dict = vocab_dict
        self.mask_token = "<PAD>"
        self.out_of_vocab_token = "<OOV>"
        self._category_vocab = [self.mask_token, self.out_of_vocab_token, *vocab_dict["category"]]
        self._attribute_vocab = [self.mask_token, self.out_of_vocab_token, *vocab_dict["attribute"]]
        self._vocab_size_category = len(self._category_vocab)
        self._vocab_size_attributes = len(self._attribute_vocab)

    def get_config(self):
        return {"name": self.name, "vocab_dict": self.vocab_dict}

    def tokens(self, inputs, conf_scale=1.0, want_confidence=False):
        """(tokens_categories [B,Q,1] int32, tokens_attributes [B,Q,A] int32[, confidence [B,Q], image_confidence [B]])."""
        cat_preds, attribute_preds = (f32(t) for t in inputs)
        B, Q, C = cat_preds.shape
        A = attribute_preds.shape[-1]
        tok_c = empty(B, Q, dtype=torch.int32)
        tok_a = empty(B, Q, A, dtype=torch.int32)
        conf = empty(B, Q) if want_confidence else None
        iconf = empty(B) if want_confidence else None
        _lib.call("bdetr_inverse_tokenize", B, Q, C, A, ptr(cat_preds), ptr(attribute_preds), ptr(tok_c), ptr(tok_a), ptr(conf),
                  ptr(iconf), float(conf_scale), stream_ptr())
        if want_confidence:
            return tok_c.unsqueeze(-1), tok_a, conf, iconf
        return tok_c.unsqueeze(-1), tok_a

    def sparce_to_strings(self, tokens_categories, tokens_attributes):
        """Host-side lookup (reference :113-124): nested lists of strings."""
        tc = tokens_categories.squeeze(-1).cpu().tolist()
        ta = tokens_attributes.cpu().tolist()
        cats = [[self._category_vocab[t] if 0 <= t < self._vocab_size_category else self.out_of_vocab_token for t in row] for row in tc]
        attrs = [[[self._attribute_vocab[t] if 0 <= t < self._vocab_size_attributes else self.out_of_vocab_token for t in q] for q in row]
                 for row in ta]
        return cats, attrs

    def call(self, inputs, training=False, strings=True):
        tok_c, tok_a = self.tokens(inputs)
        return self.sparce_to_strings(tok_c, tok_a) if strings else (tok_c, tok_a)

    __call__ = call
