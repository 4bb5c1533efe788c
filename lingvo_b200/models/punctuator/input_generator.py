"""Punctuator input (ref `lingvo/tasks/punctuator/input_generator.py`).

Each text line is the *target*; the *source* is the same line lower-cased with
punctuation stripped (ref :100-140). Lines stream through the native shuffling
yielder (`TextLines` = the reference's `tf.data.TextLineDataset` source) and are
bucketed by length (`BatchBySequenceLength`).
"""

from __future__ import annotations

import string

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import datasource
from lingvo_b200.core import generic_input
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap

_PUNCT = str.maketrans('', '', string.punctuation)


class TextLines(datasource.DataSource):
  """Lines of text files, shuffled unless sequential order is required (ref :25)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('file_pattern', None, 'Text file glob.')
    p.Define('shuffle_buffer_size', 10000, 'Shuffle buffer (lines).')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._y = None

  def GetNext(self):
    p = self.params
    if self._y is None:
      seq = self.cluster.require_sequential_input_order
      pat = p.file_pattern if ':' in p.file_pattern.split('/')[0] else 'text:' + p.file_pattern
      self._y = generic_input.MakeYielder(
          pat, p.random_seed or 0, p.shuffle_buffer_size, 1 if seq else 4,
          repeat_count=1 if seq else -1, require_sequential_order=seq)
    rec = self._y.next()
    if rec is None:
      raise StopIteration()
    return rec[0].decode('utf-8', errors='replace')

  def Reset(self, sess=None):
    self._y = None


class PunctuatorInput(base_input_generator.BaseInputGenerator):
  """Batches of `src{ids,paddings}` / `tgt{ids,labels,paddings,weights}` (ref :56)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.file_datasource = TextLines.Params()
    p.Define('tokenizer', tokenizers.WpmTokenizer.Params(), 'Tokenizer.')
    p.Define('source_max_length', None, 'Max source length (None: batch max).')
    p.Define('target_max_length', None, 'Max target length.')
    p.Define('bucket_upper_bound', [], 'Length buckets.')
    p.Define('bucket_batch_limit', [], 'Batch size per bucket.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('tokenizer', p.tokenizer)
    self.CreateChild('lines', p.file_datasource)
    self._buckets = [[] for _ in p.bucket_upper_bound]

  def IdsToStrings(self, ids, lens, key=None):
    return self.tokenizer.IdsToStrings(ids, lens)

  def _Example(self, line):
    p = self.params
    tgt = line.strip()
    if not tgt:
      return None
    src = tgt.lower().translate(_PUNCT)
    _, s_lab, s_pad = self.tokenizer.StringsToIds([src], p.source_max_length or 512)
    t_ids, t_lab, t_pad = self.tokenizer.StringsToIds([tgt], p.target_max_length or 512)
    ns, nt = int((1 - s_pad[0]).sum()), int((1 - t_pad[0]).sum())
    if ns == 0 or nt == 0:
      return None
    return (s_lab[0, :ns], t_ids[0, :nt], t_lab[0, :nt]), max(ns, nt)

  def _Merge(self, items):
    def _Pad(seqs, value=0):
      n = max(len(s) for s in seqs)
      out = torch.full((len(seqs), n), value, dtype=torch.int64)
      mask = torch.zeros(len(seqs), n)
      for i, s in enumerate(seqs):
        out[i, :len(s)] = s
        mask[i, :len(s)] = 1.0
      return out, mask
    src, sm = _Pad([it[0][0] for it in items])
    tid, tm = _Pad([it[0][1] for it in items])
    tlab, _ = _Pad([it[0][2] for it in items])
    return NestedMap(
        src=NestedMap(ids=src, paddings=1.0 - sm, weights=sm),
        tgt=NestedMap(ids=tid, labels=tlab, paddings=1.0 - tm, weights=tm),
        bucket_keys=torch.tensor([it[1] for it in items]))

  def _InputBatch(self):
    p = self.params
    bounds = np.asarray(p.bucket_upper_bound)
    while True:
      ex = self._Example(self.lines.GetNext())
      if ex is None:
        continue
      k = int(np.searchsorted(bounds, ex[1], side='left'))
      if k >= len(self._buckets):
        continue
      self._buckets[k].append(ex)
      if len(self._buckets[k]) >= p.bucket_batch_limit[k]:
        items, self._buckets[k] = self._buckets[k], []
        return self._Merge(items)
