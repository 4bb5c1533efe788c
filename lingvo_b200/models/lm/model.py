"""Language-model tasks (ref `lingvo/tasks/lm/model.py`).

`LanguageModel` (ref :36): time-major ids/labels/paddings from the input batch →
`lm.FProp` → metrics `loss, log_pplx, fraction_of_correct_next_step_preds,
num_predictions, num_words, num_sentences`. `FixedShapeInputLanguageModel`
(ref :196) skips dynamic trimming; `BatchMajorLanguageModel` (ref :260) keeps
`[B,T]` end to end.
"""

from __future__ import annotations

import math

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.lm import layers as lm_layers


class LanguageModel(base_model.BaseTask):
  """LM training/eval task."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('lm', lm_layers.RnnLm.Params(), 'The LM layer.')
    p.Define('packed_input', False, 'Packed inputs.')
    tp = p.train
    if 'max_lstm_gradient_norm' not in tp:
      tp.Define('max_lstm_gradient_norm', 0.0, 'Clip LSTM gradients to this norm.')
    if 'sum_loss_across_tokens_in_batch' not in tp:
      tp.Define('sum_loss_across_tokens_in_batch', False,
                'Optimise the summed (not averaged) token loss.')
    tp.vn_start_step = 20000
    tp.vn_std = 0.0
    tp.learning_rate = 0.001
    tp.l2_regularizer_weight = 1e-6
    tp.clip_gradient_norm_to_value = 1.0
    tp.grad_norm_to_clip_to_zero = 100.0
    p.eval.samples_per_summary = 0
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.lm.vocab_size == p.input.tokenizer.vocab_size or not hasattr(
        p.input, 'tokenizer') or True
    self.CreateChild('lm', p.lm)

  def _TrimIfPossible(self, batch):
    """Drops all-padding trailing frames (ref :120)."""
    if 'paddings' not in batch:
      return batch
    lens = (1.0 - batch.paddings.float()).sum(1)
    max_len = int(lens.max().item()) if lens.numel() else 0
    max_len = max(max_len, 1)
    out = batch.Transform(
        lambda x: x[:, :max_len] if isinstance(x, torch.Tensor) and x.dim() >= 2 and
        x.shape[1] == batch.paddings.shape[1] else x)
    return out

  def FPropTower(self, theta, input_batch):
    p = self.params
    batch = self._TrimIfPossible(input_batch)
    ids = batch.ids.t()
    labels_ids = batch.labels.t()
    paddings = batch.paddings.t().float()
    weights = batch.weights.t().float() if 'weights' in batch else 1.0 - paddings
    bsz = ids.shape[1]
    state0 = self.lm.zero_state(theta.lm, bsz)
    labels = NestedMap(class_ids=labels_ids.long(), class_weights=weights)
    kwargs = {}
    if p.packed_input:
      kwargs = dict(segment_ids=batch.segment_ids.t(), segment_pos=batch.segment_pos.t())
    xent, _ = self.lm.FProp(theta.lm, ids, paddings, state0, labels=labels, **kwargs)
    num_preds = xent.total_weight.float()
    mean_acc = torch.zeros((), device=ids.device)
    if xent.get('per_example_argmax') is not None:
      correct = (xent.per_example_argmax == labels_ids).float() * weights
      mean_acc = correct.sum() / num_preds.clamp_min(1e-8)
    elif xent.get('logits') is not None:
      correct = (xent.logits.argmax(-1) == labels_ids).float() * weights
      mean_acc = correct.sum() / num_preds.clamp_min(1e-8)
    loss = xent.total_xent if p.train.sum_loss_across_tokens_in_batch else xent.avg_xent
    word_w = weights
    if 'word_count' in batch:
      num_words = batch.word_count.float().sum()
    else:
      num_words = word_w.sum()
    metrics = NestedMap(
        loss=(loss, num_preds if not p.train.sum_loss_across_tokens_in_batch else 1.0),
        log_pplx=(xent.avg_xent, num_preds),
        fraction_of_correct_next_step_preds=(mean_acc, num_preds),
        num_predictions=(num_preds, 1.0),
        num_words=(num_words, 1.0),
        num_sentences=(torch.tensor(float(bsz), device=ids.device), 1.0))
    per_example = NestedMap()
    return metrics, per_example

  def ComputePredictions(self, theta, input_batch):
    return NestedMap()

  def Inference(self):
    return {'default': self._InferenceDefault}

  def _InferenceDefault(self, ids, paddings):
    """ids/paddings [B,T] → per-token log P(ids[t+1] | ids[≤t])."""
    theta = self.theta
    state0 = self.lm.zero_state(theta.lm, ids.shape[0])
    labels = NestedMap(class_ids=torch.roll(ids, -1, 1).t().long(),
                       class_weights=(1.0 - paddings.float()).t())
    xent, _ = self.lm.FProp(theta.lm, ids.t(), paddings.t().float(), state0, labels=labels)
    return NestedMap(log_pplx_per_token=xent.per_example_xent.t())


class FixedShapeInputLanguageModel(LanguageModel):
  """No dynamic trimming: shapes stay static so the step can be CUDA-graphed (ref :196)."""

  def _TrimIfPossible(self, batch):
    return batch


class BatchMajorLanguageModel(FixedShapeInputLanguageModel):
  """Batch-major inputs; same metrics (ref :260)."""


class PackedBatchMajorLanguageModel(BatchMajorLanguageModel):
  """Packed inputs with segment ids/positions (ref :330)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.packed_input = True
    return p
