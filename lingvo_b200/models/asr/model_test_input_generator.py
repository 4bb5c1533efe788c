"""Deterministic random input for ASR model tests (ref
`lingvo/tasks/asr/model_test_input_generator.py:21`).

Every call to `GetPreprocessedInputBatch` yields a fresh random batch with the requested
shapes; sequences get random valid lengths (≥ half of max) with trailing padding.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core.nested_map import NestedMap


class TestInputGenerator(base_input_generator.BaseSequenceInputGenerator):
  __test__ = False   # not a pytest class

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('feature_dims', 240, 'Feature dims (unused when source_shape is given).')
    p.Define('num_channels', 3, 'Feature channels.')
    p.Define('source_shape', [2, 10, 8, 3], '[batch, time, freq, channels].')
    p.Define('target_shape', [2, 5], '[batch, target_len].')
    p.Define('fixed_target_labels', None, 'Use these labels instead of random ones.')
    p.Define('fixed_target_ids', None, 'Use these ids instead of shifted labels.')
    p.Define('cur_iter_in_seed', True, 'Mix the batch counter into the seed.')
    p.Define('integer_source_max', None, 'If set, sources are integers in [0, max).')
    p.Define('float_source_max', None, 'If set, sources are uniform in [0, max).')
    p.Define('for_mt', False, 'MT-style batch (src.ids / src.paddings).')
    p.Define('target_key', '', 'If set, nest targets under batch.tgt[target_key].')
    p.Define('target_key_target_shape', [2, 5], 'Shape of the keyed targets.')
    p.Define('set_tgt_and_additional_tgts', False, 'Provide both tgt and additional_tgts.')
    p.Define('target_language', 'ENGLISH', 'Kept for parity.')
    p.Define('align_label_with_frame', False, 'Frame-aligned labels (RNN-T style).')
    p.Define('bprop_filters', [], 'Variable-name filters (multi-task tests).')
    p.Define('number_sources', None, 'Multi-source input: number of sources.')
    p.Define('source_selected', None, 'Multi-source input: active source.')
    p.Define('target_transcript', 'dummy_transcript', 'Transcript string.')
    p.random_seed = 20349582
    return p

  def __init__(self, params):
    super().__init__(params)
    self._iter = 0

  def _Gen(self):
    p = self.params
    g = torch.Generator()
    g.manual_seed(int(p.random_seed) + (self._iter if p.cur_iter_in_seed else 0))
    return g

  def _Paddings(self, g, batch, length):
    lens = torch.randint(max(length // 2, 1), length + 1, (batch,), generator=g)
    return (torch.arange(length).unsqueeze(0) >= lens.unsqueeze(1)).float()

  def GlobalBatchSize(self):
    return self.params.source_shape[0]

  def InfeedBatchSize(self):
    return self.params.source_shape[0]

  def SampleIds(self):
    return torch.arange(self.params.source_shape[0])

  def _Sources(self, g):
    p = self.params
    shape = list(p.source_shape)
    pad = self._Paddings(g, shape[0], shape[1])
    if p.for_mt:
      ids = torch.randint(0, p.integer_source_max or p.tokenizer.vocab_size,
                          (shape[0], shape[1]), generator=g)
      return NestedMap(ids=ids, paddings=pad)
    if p.integer_source_max:
      x = torch.randint(0, p.integer_source_max, shape, generator=g).float()
    elif p.float_source_max:
      x = torch.rand(shape, generator=g) * p.float_source_max
    else:
      x = torch.randn(shape, generator=g)
    return NestedMap(src_inputs=x, paddings=pad)

  def _Targets(self, g, shape):
    p = self.params
    b, t = shape
    vocab = p.tokenizer.vocab_size
    if p.fixed_target_labels is not None:
      labels = torch.as_tensor(p.fixed_target_labels).long().reshape(b, t)
    else:
      labels = torch.randint(3, max(vocab, 4), (b, t), generator=g)
    if p.fixed_target_ids is not None:
      ids = torch.as_tensor(p.fixed_target_ids).long().reshape(b, t)
    else:
      sos = torch.full((b, 1), p.tokenizer.target_sos_id, dtype=torch.long)
      ids = torch.cat([sos, labels[:, :-1]], 1)
    pad = self._Paddings(g, b, t)
    return NestedMap(ids=ids, labels=labels, paddings=pad, weights=1.0 - pad,
                     transcripts=[p.target_transcript] * b)

  def _InputBatch(self):
    p = self.params
    g = self._Gen()
    self._iter += 1
    batch = NestedMap(src=self._Sources(g), sample_ids=self.SampleIds())
    if p.target_key:
      keyed = self._Targets(g, p.target_key_target_shape)
      if p.set_tgt_and_additional_tgts:
        batch.tgt = self._Targets(g, p.target_shape)
        batch.additional_tgts = NestedMap({p.target_key: keyed})
      else:
        batch.tgt = NestedMap({p.target_key: keyed})
    else:
      batch.tgt = self._Targets(g, p.target_shape)
    if p.number_sources:
      batch.src.source_selected = torch.full(
          (p.source_shape[0],), int(p.source_selected or 0), dtype=torch.long)
    return batch

  def GetPreprocessedInputBatch(self):
    return self._InputBatch()

  def GetBpropParams(self):
    return self.params.bprop_filters
