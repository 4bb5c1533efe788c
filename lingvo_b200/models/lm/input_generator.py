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


class TFRecordBertInput(base_input_generator.BaseInputGenerator):
  """MLPerf-BERT TFRecords (ref :267): fixed-length `input_ids / input_mask /
  masked_lm_positions / masked_lm_ids / masked_lm_weights`. Restores the un-masked ids,
  drops the first `[SEP]` (documents are `A [SEP] B [SEP]`), and optionally packs several
  documents per row with the native packer. Output `[batch_size, max_sequence_length]`:
  `ids, masked_ids, masked_pos, segment_ids, segment_pos, paddings`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_file', None, 'File pattern(s) of TFRecord files.')
    p.Define('max_sequence_length', 512, 'Tokens per example.')
    p.Define('max_predictions_per_seq', 76, 'Masked positions per example.')
    p.Define('eos_token_id', 102, '[SEP] id.')
    p.Define('shuffle', False, 'Shuffle records.')
    p.Define('file_buffer_size', 10000000, 'Shuffle buffer (records).')
    p.Define('enable_packing', False, 'Pack several documents per row.')
    p.Define('prepacking_batch_size', 1 << 14, 'Documents gathered before packing.')
    p.Define('remove_mask', False, 'Drop the stored masking (mask on the fly instead).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if self.do_eval:
      p.shuffle = False
      p.enable_packing = False
    self._files = [p.input_file] if isinstance(p.input_file, str) else list(p.input_file)
    self._yielder = None
    self._pending = []

  def _Yielder(self):
    if self._yielder is None:
      import glob
      from lingvo_b200 import ops
      p = self.params
      files = sorted(f for pat in self._files for f in glob.glob(pat))
      if not files:
        raise FileNotFoundError('TFRecordBertInput: no files match %s' % self._files)
      pattern = 'tfrecord:' + ','.join(files)
      h = ops.host()
      if p.shuffle:
        self._yielder = h.basic_record_yielder(
            pattern, seed=int(p.random_seed or 0) + 1,
            bufsize=int(min(p.file_buffer_size, 1 << 16)), parallelism=4, num_epochs=0)
      else:
        self._yielder = h.sequential_record_yielder(pattern, 1 if self.do_eval else -1)
    return self._yielder

  def _ParseRecord(self, record):
    from lingvo_b200.utils import tf_example
    p = self.params
    f = tf_example.ParseExample(record)
    masked_ids = np.asarray(f['input_ids'], np.int64)[:p.max_sequence_length]
    valid = np.asarray(f['input_mask'], np.float32)[:p.max_sequence_length]
    w = np.asarray(f['masked_lm_weights'], np.float32)
    n = int(w.sum())
    pos = np.asarray(f['masked_lm_positions'], np.int64)[:n]
    ids = masked_ids.copy()
    ids[pos] = np.asarray(f['masked_lm_ids'], np.int64)[:n]
    masked_pos = np.zeros_like(valid)
    masked_pos[pos] = 1.0
    sep = np.nonzero(ids == p.eos_token_id)[0]
    def _Drop(x):
      if sep.size == 0:
        return x
      k = int(sep[0])
      return np.concatenate([x[:k], x[k + 1:], np.zeros(1, x.dtype)])
    out = NestedMap(ids=_Drop(ids), masked_ids=_Drop(masked_ids), masked_pos=_Drop(masked_pos),
                    segment_ids=_Drop(valid))
    out.paddings = 1.0 - out.segment_ids
    out.segment_pos = (out.segment_ids * np.arange(len(valid))).astype(np.int64)
    if p.remove_mask:
      del out['masked_pos']
      del out['masked_ids']
    return out

  def _Next(self):
    rec = self._Yielder().next()
    if rec is None:
      return None
    return self._ParseRecord(rec[0] if isinstance(rec, tuple) else rec)

  def _Pack(self, docs):
    """List of parsed documents → list of packed rows."""
    p = self.params
    lens = np.asarray([int(d.segment_ids.sum()) for d in docs], np.int32)
    seg, pos, idx, _, _, _ = host_ops.PackSequences(
        lens, lens, 0, p.max_sequence_length, p.max_sequence_length,
        seed=int(p.random_seed or 0))
    rows = NestedMap()
    for k in docs[0].keys():
      if k in ('segment_ids', 'segment_pos', 'paddings'):
        continue
      rows[k] = host_ops.ApplyPacking(np.stack([d[k] for d in docs]), 0, seg, idx)
    rows.segment_ids = seg.astype(np.float32)
    rows.segment_pos = pos.astype(np.int64)
    rows.paddings = (seg == 0).astype(np.float32)
    n = seg.shape[0]
    return [rows.Transform(lambda x, i=i: x[i]) for i in range(n) if seg[i].any()]

  def _InputBatch(self):
    p = self.params
    rows = []
    while len(rows) < p.batch_size:
      if self._pending:
        rows.append(self._pending.pop())
        continue
      if p.enable_packing:
        docs = []
        while len(docs) < p.prepacking_batch_size:
          d = self._Next()
          if d is None:
            break
          docs.append(d)
        if not docs:
          break
        self._pending = self._Pack(docs)
        continue
      d = self._Next()
      if d is None:
        break
      rows.append(d)
    if not rows:
      raise StopIteration('TFRecordBertInput: end of data')
    batch = NestedMap()
    for k in rows[0].keys():
      x = np.stack([r[k] for r in rows])
      need = p.batch_size - x.shape[0]
      if need > 0:                     # pad the final eval batch
        fill = 1.0 if k == 'paddings' else 0
        x = np.concatenate([x, np.full((need,) + x.shape[1:], fill, x.dtype)])
      batch[k] = torch.from_numpy(np.ascontiguousarray(x))
    return batch

  def Reset(self, sess=None):
    self._yielder = None
    self._pending = []
