"""MT input generators (ref `lingvo/tasks/mt/input_generator.py`).

`NmtInput` (ref :30) reads `tf.Example` records holding pre-tokenised
`source_id, source_padding, target_id, target_padding, target_label,
target_weight`; the bucket key is max(src_len, tgt_len). Output batch:
`NestedMap(src=NestedMap(ids, paddings), tgt=NestedMap(ids, labels, weights,
paddings), bucket_keys)`, all `[B, T]`.
`TextMtInput` reads `source<TAB>target` text lines through the tokenizer.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.utils import tf_example


def _Pack(src_ids, src_pad, tgt_ids, tgt_pad, tgt_labels, tgt_w):
  # `mask` = 1 on real frames; the batcher zero-pads, so paddings = 1 - mask afterwards.
  return NestedMap(
      src=NestedMap(ids=src_ids.astype(np.int32), mask=(1.0 - src_pad).astype(np.float32)),
      tgt=NestedMap(ids=tgt_ids.astype(np.int32), labels=tgt_labels.astype(np.int32),
                    weights=tgt_w.astype(np.float32), mask=(1.0 - tgt_pad).astype(np.float32)))


class NmtInput(base_input_generator.BaseSequenceInputGenerator):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('natural_order_model', True, 'Targets in natural (not reversed) order.')
    p.tokenizer = tokenizers.VocabFileTokenizer.Params()
    p.source_max_length = 300
    return p

  def ProcessRecord(self, record, source_id=0):
    f = tf_example.ParseExample(record)
    src_pad = f['source_padding'].astype(np.float32)
    tgt_pad = f['target_padding'].astype(np.float32)
    key = int(max((1.0 - src_pad).sum(), (1.0 - tgt_pad).sum()))
    n_s, n_t = int((1 - src_pad).sum()), int((1 - tgt_pad).sum())
    out = _Pack(f['source_id'][:n_s], src_pad[:n_s], f['target_id'][:n_t], tgt_pad[:n_t],
                f['target_label'][:n_t], f['target_weight'][:n_t].astype(np.float32))
    return out, key

  def _PreprocessInputBatch(self, batch):
    batch.src.paddings = 1.0 - batch.src.mask
    batch.tgt.paddings = 1.0 - batch.tgt.mask
    batch.src.weights = batch.src.mask
    if not self.params.natural_order_model:
      batch.tgt = batch.tgt.Transform(lambda x: torch.flip(x, [1]))
    return batch


class TextMtInput(NmtInput):
  """`source<TAB>target` text lines, tokenised on the fly."""

  def ProcessRecord(self, record, source_id=0):
    p = self.params
    line = record.decode('utf-8', errors='replace').rstrip('\n')
    if '\t' not in line:
      return None
    src, tgt = line.split('\t', 1)
    _, s_lab, s_pad = self.StringsToIds([src], is_source=True)
    t_ids, t_lab, t_pad = self.StringsToIds([tgt])
    n_s, n_t = int((1 - s_pad[0]).sum()), int((1 - t_pad[0]).sum())
    if n_s == 0 or n_t == 0:
      return None
    key = max(n_s, n_t)
    if key > p.bucket_upper_bound[-1]:
      return None
    out = _Pack(s_lab[0, :n_s].numpy(), s_pad[0, :n_s].numpy(), t_ids[0, :n_t].numpy(),
                t_pad[0, :n_t].numpy(), t_lab[0, :n_t].numpy(),
                (1.0 - t_pad[0, :n_t]).numpy())
    return out, key
