"""Input rewriting for sequence-to-sequence pre-training (ref
`lingvo/tasks/mt/data_augmenter.py`).

`MASS` (ref :26) wraps the native MASS span-masking op (`ops/csrc_host/text_ops.cpp`):
a contiguous span of each sentence is masked on the encoder side and becomes the
decoder's prediction target.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.ops import host_ops


class MASS(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('mask_id', 3, 'Id of the mask token.')
    p.Define('mask_ratio', 0.5, 'Fraction of each sentence that is masked.')
    p.Define('mask_minlen', 0, 'Sentences shorter than this are not masked.')
    p.Define('span_len', 100000, 'Masked-span length (≥ sentence length → one span).')
    p.Define('random_start_prob', 0.6, 'Probability of a random (vs. edge) span start.')
    p.Define('keep_prob', 0.1, 'Probability a selected token is kept as is.')
    p.Define('rand_prob', 0.1, 'Probability a selected token becomes a random token.')
    p.Define('mask_prob', 0.8, 'Probability a selected token becomes the mask token.')
    p.Define('mask_target', True, 'Mask the decoder inputs outside the span.')
    p.Define('vocab_size', 0, 'Vocabulary size (for random replacement).')
    p.Define('first_unreserved_id', 4, 'First id eligible as a random replacement.')
    p.name = 'mass'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if abs(p.keep_prob + p.rand_prob + p.mask_prob - 1.0) > 1e-6:
      raise ValueError('keep_prob + rand_prob + mask_prob must sum to 1')
    self._calls = 0

  def Mask(self, seq_ids, weights, actual_seq_len):
    """seq_ids/weights [B,T], actual_seq_len [B] → NestedMap(src(ids), tgt(ids, labels,
    weights)) with the masked encoder input and span-only decoder targets (ref :81)."""
    p = self.params
    dev = seq_ids.device if isinstance(seq_ids, torch.Tensor) else None
    to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    seed = (p.random_seed or 0) * 7919 + self._calls
    self._calls += 1
    src, tgt, lab, w = host_ops.Mass(
        to_np(seq_ids), to_np(weights), to_np(actual_seq_len), mask_id=p.mask_id,
        mask_ratio=p.mask_ratio, mask_minlen=p.mask_minlen, span_len=p.span_len,
        random_start_prob=p.random_start_prob, keep_prob=p.keep_prob, rand_prob=p.rand_prob,
        mask_prob=p.mask_prob, mask_target=p.mask_target, vocab_size=p.vocab_size,
        first_unreserved_id=p.first_unreserved_id, seed=seed)
    t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt, device=dev)
    out = NestedMap(src=NestedMap(ids=t(src, torch.int64)),
                    tgt=NestedMap(ids=t(tgt, torch.int64), labels=t(lab, torch.int64),
                                  weights=t(w, torch.float32)))
    return out
