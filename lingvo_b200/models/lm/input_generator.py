"""LM input generators (ref `lingvo/tasks/lm/input_generator.py`).

`LmInput` (ref :30): text files, one sentence per line → tokenizer → bucketed
batches `{ids, labels, paddings, weights, word_count}` (`[B, T]`).
`PackedTextInputGenerator` (ref :150): packs several sentences per row with
`segment_ids/segment_pos` using the native packer.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.ops import host_ops


class LmInput(base_input_generator.BaseSequenceInputGenerator):
  """Reads tokenised LM inputs from text files."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fixed_input_shape', False, 'Pad every batch to target_max_length.')
    p.tokenizer = tokenizers.AsciiTokenizer.Params()
    p.file_pattern = ''
    p.bucket_upper_bound = [10, 20, 30, 40, 50, 100]
    p.bucket_batch_limit = [64] * 6
    p.target_max_length = 100
    return p

  def ProcessRecord(self, record, source_id=0):
    p = self.params
    text = record.decode('utf-8', errors='replace').strip()
    if not text:
      return None
    ids, labels, paddings = self.StringsToIds([text])
    n = int((1.0 - paddings[0]).sum())
    if n <= 0 or n > p.bucket_upper_bound[-1]:
      return None
    t = p.target_max_length if p.fixed_input_shape else n
    out = NestedMap(
        ids=ids[0, :t].numpy().astype(np.int32),
        labels=labels[0, :t].numpy().astype(np.int32),
        paddings=paddings[0, :t].numpy().astype(np.float32),
        weights=(1.0 - paddings[0, :t]).numpy().astype(np.float32),
        word_count=np.int32(len(text.split()) + 1))
    return out, n

  def _PreprocessInputBatch(self, batch):
    # padded frames produced by the batcher are zero: mark them as padding
    if 'weights' in batch:
      batch.paddings = 1.0 - batch.weights
    return batch


class PackedTextInputGenerator(LmInput):
  """Packs sentences into fixed `[B, T]` rows (ref :150)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('packed_len', 128, 'Row length T.')
    p.Define('packed_batch_size', 8, 'Rows per batch B.')
    p.Define('sentences_per_pack', 64, 'Sentences gathered before packing.')
    return p

  def _InputBatch(self):
    p = self.params
    saved_limits = p.bucket_batch_limit
    raw = super()._InputBatch()
    del saved_limits
    ids, labels = raw.ids.numpy(), raw.labels.numpy()
    lens = raw.weights.numpy().sum(1).astype(np.int32)
    seg, pos, idx, _, _, _ = host_ops.PackSequences(
        lens, lens, p.packed_batch_size, p.packed_len, p.packed_len,
        seed=p.file_random_seed)
    out = NestedMap(
        ids=torch.from_numpy(host_ops.ApplyPacking(ids, 0, seg, idx).astype(np.int64)),
        labels=torch.from_numpy(host_ops.ApplyPacking(labels, 0, seg, idx).astype(np.int64)),
        segment_ids=torch.from_numpy(seg.astype(np.int64)),
        segment_pos=torch.from_numpy(pos.astype(np.int64)))
    out.weights = (out.segment_ids > 0).float()
    out.paddings = 1.0 - out.weights
    return out
