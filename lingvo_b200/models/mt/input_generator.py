"""MT input generators (ref `lingvo/tasks/mt/input_generator.py`).

`NmtInput` (ref :30) reads `tf.Example` records holding pre-tokenised
`source_id, source_padding, target_id, target_padding, target_label,
target_weight`; the bucket key is max(src_len, tgt_len). Output batch:
`NestedMap(src=NestedMap(ids, paddings), tgt=NestedMap(ids, labels, weights,
paddings), bucket_keys)`, all `[B, T]`.
`TextMtInput` reads `source<TAB>target` text lines through the tokenizer;
`MlPerfInput` (ref :159), `TextPackedInput` (ref :409) and `NmtDoubleInput` (ref :1112)
cover the MLPerf, packed multi-task and XEnDec pipelines.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.ops import host_ops
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


class MlPerfInput(base_input_generator.BaseSequenceInputGenerator):
  """MLPerf-Transformer style records (ref :159): `tf.Example`s with int64 lists
  `inputs` / `targets` (already word-pieces, ending in EOS). With `packed_input` the
  records are pre-packed and also carry `inputs_position/targets_position` and
  `inputs_segmentation/targets_segmentation`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('natural_order_model', True, 'Only natural target order is supported.')
    p.Define('sos_id', 0, 'Start-of-sentence id used for the shifted decoder input.')
    p.Define('packed_input', False, 'Records are pre-packed rows.')
    p.tokenizer = tokenizers.VocabFileTokenizer.Params()
    p.source_max_length = 300
    return p

  def ProcessRecord(self, record, source_id=0):
    p = self.params
    f = tf_example.ParseExample(record)
    src = np.asarray(f['inputs'], np.int32)
    lab = np.asarray(f['targets'], np.int32)
    if p.packed_input:
      n_s, n_t = p.source_max_length, p.target_max_length
      fix = lambda a, n: np.pad(np.asarray(a, np.int32)[:n], (0, max(0, n - len(a))))
      src, lab = fix(src, n_s), fix(lab, n_t)
      src_seg, tgt_seg = fix(f['inputs_segmentation'], n_s), fix(f['targets_segmentation'], n_t)
      src_pos, tgt_pos = fix(f['inputs_position'], n_s), fix(f['targets_position'], n_t)
      # decoder input: labels shifted right inside every segment, SOS at segment starts
      ids = np.where(tgt_pos == 0, p.sos_id, np.concatenate([[p.sos_id], lab[:-1]]))
      out = NestedMap(
          src=NestedMap(ids=src, mask=(src_seg > 0).astype(np.float32), segment_ids=src_seg,
                        segment_pos=src_pos),
          tgt=NestedMap(ids=ids.astype(np.int32), labels=lab,
                        weights=(tgt_seg > 0).astype(np.float32),
                        mask=(tgt_seg > 0).astype(np.float32), segment_ids=tgt_seg,
                        segment_pos=tgt_pos))
      return out, 1
    n_s, n_t = len(src), len(lab)
    key = max(n_s, n_t)
    if key == 0 or key > p.bucket_upper_bound[-1]:
      return None
    ids = np.concatenate([[p.sos_id], lab[:-1]]).astype(np.int32)
    return _Pack(src, np.zeros(n_s, np.float32), ids, np.zeros(n_t, np.float32), lab,
                 np.ones(n_t, np.float32)), key

  def _PreprocessInputBatch(self, batch):
    batch.src.paddings = 1.0 - batch.src.mask
    batch.tgt.paddings = 1.0 - batch.tgt.mask
    batch.src.weights = batch.src.mask
    return batch


class TextPackedInput(base_input_generator.BaseSequenceInputGenerator):
  """Text pairs → tokenise → (optional MASS) → pack into fixed `[B, T]` rows (ref :409).

  Records are `source<TAB>target` lines (or single-column lines for monolingual MASS
  tasks). Each file pattern carries a task id (`file_pattern_task_ids`) which maps to
  source / target language ids. `packing_factor > 0` gathers `packing_factor × batch`
  sentence pairs and packs them with the native `PackSequences` op."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('file_pattern_task_ids', [], 'Task id per file pattern.')
    p.Define('task_to_src_lang_map', [], 'task id → source language id.')
    p.Define('task_to_tgt_lang_map', [], 'task id → target language id.')
    p.Define('packing_factor', 0.0, 'Average sentences per packed row (0: no packing).')
    p.Define('quality_score_filter_threshold', -1e9, 'Drop pairs scoring below this.')
    p.Define('natural_order_model', True, 'Only natural order is supported.')
    p.Define('target_language', '', 'Target language tag.')
    p.Define('mass_layer', None, 'MASS layer params for monolingual tasks.')
    p.Define('mass_task_ids', [], 'Task ids that go through MASS.')
    p.Define('enable_mass_for_eval', False, 'Apply MASS in eval too.')
    p.Define('single_column_input', False, 'Lines hold one sentence (MASS / LM data).')
    p.Define('suppress_id_histograms', True, 'Kept for parity.')
    p.Define('bt_task_ids', [], 'Back-translation task ids (tagged, otherwise normal).')
    p.tokenizer = tokenizers.VocabFileTokenizer.Params()
    p.source_max_length = 100
    p.target_max_length = 100
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if p.mass_layer is not None:
      self.CreateChild('mass_layer', p.mass_layer)

  def _GetTaskIds(self, source_id):
    ids = self.params.file_pattern_task_ids
    return ids[source_id] if ids else 0

  def _GetLangIds(self, source_id):
    p = self.params
    task = self._GetTaskIds(source_id)
    s = p.task_to_src_lang_map[task] if p.task_to_src_lang_map else 0
    t = p.task_to_tgt_lang_map[task] if p.task_to_tgt_lang_map else 0
    return s, t

  def _ReadRecordTsv(self, record):
    cols = record.decode('utf-8', errors='replace').rstrip('\n').split('\t')
    if self.params.single_column_input:
      return cols[0], cols[0], 0.0
    if len(cols) < 2:
      return None
    score = float(cols[2]) if len(cols) > 2 and cols[2] else 0.0
    return cols[0], cols[1], score

  def ProcessRecord(self, record, source_id=0):
    p = self.params
    rec = self._ReadRecordTsv(record)
    if rec is None or rec[2] < p.quality_score_filter_threshold:
      return None
    src, tgt, _ = rec
    task = self._GetTaskIds(source_id)
    src_lang, tgt_lang = self._GetLangIds(source_id)
    _, s_lab, s_pad = self.StringsToIds([src], is_source=True)
    t_ids, t_lab, t_pad = self.StringsToIds([tgt])
    n_s, n_t = int((1 - s_pad[0]).sum()), int((1 - t_pad[0]).sum())
    if n_s == 0 or n_t == 0:
      return None
    src_ids = s_lab[0, :n_s].numpy()
    tgt_ids, tgt_lab = t_ids[0, :n_t].numpy(), t_lab[0, :n_t].numpy()
    tgt_w = np.ones(n_t, np.float32)
    use_mass = (p.mass_layer is not None and task in p.mass_task_ids and
                (not self.do_eval or p.enable_mass_for_eval))
    if use_mass:
      m = self.mass_layer.Mask(torch.as_tensor(src_ids[None]), torch.ones(1, n_s),
                               torch.tensor([n_s]))
      src_ids = m.src.ids[0].numpy()
      tgt_ids, tgt_lab = m.tgt.ids[0].numpy(), m.tgt.labels[0].numpy()
      tgt_w = m.tgt.weights[0].numpy()
      n_t = n_s
    out = _Pack(src_ids, np.zeros(n_s, np.float32), tgt_ids, np.zeros(n_t, np.float32),
                tgt_lab, tgt_w)
    out.src.task_ids = np.full(n_s, task, np.int32)
    out.tgt.task_ids = np.full(n_t, task, np.int32)
    out.src.source_ids = np.full(n_s, src_lang, np.int32)
    out.tgt.target_ids = np.full(n_t, tgt_lang, np.int32)
    key = max(n_s, n_t)
    if key > p.bucket_upper_bound[-1]:
      return None
    return out, key

  def _ApplyPacking(self, batch):
    """Packs a `[N, T]` batch into `[B, L]` rows with segment ids / positions (ref :873)."""
    p = self.params
    n = batch.src.ids.shape[0]
    rows = max(1, int(round(n / max(p.packing_factor, 1.0))))
    s_len = batch.src.mask.sum(1).numpy().astype(np.int32)
    t_len = batch.tgt.mask.sum(1).numpy().astype(np.int32)
    s_seg, s_pos, s_idx, t_seg, t_pos, t_idx = host_ops.PackSequences(
        s_len, t_len, rows, p.source_max_length, p.target_max_length,
        seed=p.file_random_seed)
    def _Apply(x, seg, idx):
      return torch.from_numpy(host_ops.ApplyPacking(x.numpy(), 0, seg, idx))
    out = NestedMap(src=NestedMap(), tgt=NestedMap())
    for k, v in batch.src.items():
      out.src[k] = _Apply(v, s_seg, s_idx)
    for k, v in batch.tgt.items():
      out.tgt[k] = _Apply(v, t_seg, t_idx)
    out.src.segment_ids = torch.from_numpy(s_seg.astype(np.int64))
    out.src.segment_pos = torch.from_numpy(s_pos.astype(np.int64))
    out.tgt.segment_ids = torch.from_numpy(t_seg.astype(np.int64))
    out.tgt.segment_pos = torch.from_numpy(t_pos.astype(np.int64))
    out.src.mask = (out.src.segment_ids > 0).float()
    out.tgt.mask = (out.tgt.segment_ids > 0).float()
    return out

  def _PreprocessInputBatch(self, batch):
    if self.params.packing_factor > 0:
      batch = self._ApplyPacking(batch)
    batch.src.paddings = 1.0 - batch.src.mask
    batch.tgt.paddings = 1.0 - batch.tgt.mask
    batch.src.weights = batch.src.mask
    batch.tgt.weights = batch.tgt.weights * batch.tgt.mask
    return batch


class NmtDoubleInput(NmtInput):
  """Parallel data prepared for XEnDec (ref :1112): besides the normal batch it emits

    * `src.source_mask [B,S]` — 1 where the crossover takes the *partner* sentence's token
      (a `source_mask_ratio` fraction of the shorter of the two, chosen at random);
    * optionally `other_src/other_tgt`: a noised copy of the batch (word shuffling within
      `permutation_distance`, `<mask>` replacement) used as the partner instead of the
      rolled batch.
  """

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('packed_input', False, 'Kept for parity.')
    p.Define('source_mask_ratio', 0.0, 'Fraction of source tokens taken from the partner.')
    p.Define('source_mask_ratio_beta', '', "'a,b': draw the ratio from Beta(a, b) per batch.")
    p.Define('permutation_distance', 0, 'Max displacement when shuffling words.')
    p.Define('mask_word_id', 3, '<mask> id.')
    p.Define('mask_words_ratio', 0.0, 'Fraction of words replaced by <mask> in the noised copy.')
    p.Define('pad_id', 4, '<pad> id.')
    p.Define('vocab_file', None, 'Kept for parity.')
    return p

  def _Rng(self):
    if not hasattr(self, '_rng'):
      self._rng = torch.Generator()
      self._rng.manual_seed(int(self.params.random_seed or 0) + 1234567)
    return self._rng

  def _SelectMaskPositions(self, paddings, ratio):
    """Random subset (≈ratio of each row's valid tokens) as a {0,1} mask."""
    valid = 1.0 - paddings
    noise = torch.rand(paddings.shape, generator=self._Rng()) + paddings * 2.0
    rank = noise.argsort(1).argsort(1).float()
    k = torch.floor(valid.sum(1, keepdim=True) * ratio)
    return (rank < k).float() * valid

  def _ShuffleWords(self, seq, paddings, distance):
    """Local shuffle: sort positions by position + U(0, distance+1)."""
    pos = torch.arange(seq.shape[1]).float().unsqueeze(0) + \
        torch.rand(seq.shape, generator=self._Rng()) * (distance + 1)
    pos = pos + paddings * 1e6
    return seq.gather(1, pos.argsort(1))

  def _MaskWords(self, seq, paddings, ratio):
    m = self._SelectMaskPositions(paddings, ratio)
    return torch.where(m > 0, torch.full_like(seq, self.params.mask_word_id), seq)

  def _GenerateNoiseSents(self, seq, paddings):
    p = self.params
    if p.permutation_distance > 0:
      seq = self._ShuffleWords(seq, paddings, p.permutation_distance)
    if p.mask_words_ratio > 0:
      seq = self._MaskWords(seq, paddings, p.mask_words_ratio)
    return seq

  def _CreateSourceLambdas(self, source_paddings):
    p = self.params
    ratio = p.source_mask_ratio
    if p.source_mask_ratio_beta:
      a, b = (float(v) for v in p.source_mask_ratio_beta.split(','))
      ratio = float(torch.distributions.Beta(a, b).sample())
    partner_pad = torch.roll(source_paddings, 1, 0)
    both = torch.maximum(source_paddings, partner_pad)       # pad where either is padded
    return self._SelectMaskPositions(both, ratio)

  def _PreprocessInputBatch(self, batch):
    p = self.params
    batch = super()._PreprocessInputBatch(batch)
    batch.src.source_mask = self._CreateSourceLambdas(batch.src.paddings)
    if p.permutation_distance > 0 or p.mask_words_ratio > 0:
      batch.other_src = batch.src.DeepCopy()
      batch.other_tgt = batch.tgt.DeepCopy()
      batch.other_src.ids = self._GenerateNoiseSents(batch.src.ids, batch.src.paddings)
    return batch
