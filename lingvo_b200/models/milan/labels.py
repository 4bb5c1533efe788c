"""Pair labels and the multi-label contrastive loss (ref `lingvo/tasks/milan/labels.py`).

An `ExamplePairs` describes the (query example, result example) grid to be labelled:
result examples are a superset of the query examples (the local batch against the
cross-replica global batch), `correspondences[i, j]` says result j *is* query i. Label
functions map an `ExamplePairs` to an int tensor of 1 (positive) / 0 (negative) /
−1 (ignore).
"""

from __future__ import annotations

import torch

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.milan import tpu_utils
from lingvo_b200.models.milan import utils

IGNORE_PAIR_LABEL = -1


class ExamplePairs:
  """ref :94."""

  def __init__(self, query_examples, result_examples, correspondences=None,
               query_modality='', result_modality=''):
    self.query_examples = query_examples
    self.result_examples = result_examples
    self.query_modality = query_modality
    self.result_modality = result_modality
    self.query_batch_size = utils.InferBatchSize(query_examples)
    self.result_batch_size = utils.InferBatchSize(result_examples)
    if correspondences is None:
      dev = next(iter(NestedMap(query_examples).Flatten())).device
      correspondences = torch.eye(self.query_batch_size, self.result_batch_size,
                                  dtype=torch.bool, device=dev)
    assert correspondences.dtype == torch.bool
    assert tuple(correspondences.shape) == (self.query_batch_size, self.result_batch_size)
    self.correspondences = correspondences

  @classmethod
  def WithinBatch(cls, batch, **kwargs):
    """All pairs inside one batch (ref :126)."""
    return cls(batch, batch, None, **kwargs)

  @classmethod
  def BetweenLocalAndGlobalBatches(cls, local_batch, **kwargs):
    """Local queries against the cross-replica concatenation of all batches (ref :145)."""
    local_batch = NestedMap(local_batch)
    global_batch = tpu_utils.ConcatenateAcrossReplicas(local_batch)
    n, m = utils.InferBatchSize(local_batch), utils.InferBatchSize(global_batch)
    dev = next(iter(local_batch.Flatten())).device
    corr = torch.zeros(n, m, dtype=torch.bool, device=dev)
    off = tpu_utils.ReplicaOffset(n)
    corr[torch.arange(n), torch.arange(n) + off] = True
    return cls(local_batch, global_batch, corr, **kwargs)


def _IgnorePairsWhere(condition, labels):
  return torch.where(condition, torch.full_like(labels, IGNORE_PAIR_LABEL), labels)


class ExamplePairLabeler:
  """Positives = corresponding pairs; other pairs that agree on any feature named in
  `drop_pairs_that_match` (e.g. an image id shared by several captions) are ignored
  instead of being used as negatives (ref :175)."""

  def __init__(self, drop_pairs_that_match=()):
    if isinstance(drop_pairs_that_match, str):
      drop_pairs_that_match = [drop_pairs_that_match]
    self._drop_on_match = list(drop_pairs_that_match)

  def __call__(self, inputs: ExamplePairs):
    labels = inputs.correspondences.to(torch.int32)
    drop = None
    for name in self._drop_on_match:
      q, r = inputs.query_examples.get(name), inputs.result_examples.get(name)
      if q is None:
        raise ValueError('No feature {} in query batch'.format(name))
      if r is None:
        raise ValueError('No feature {} in result batch'.format(name))
      assert q.shape == (inputs.query_batch_size,) and r.shape == (inputs.result_batch_size,)
      m = q[:, None] == r[None, :]
      drop = m if drop is None else (drop | m)
    if drop is not None:
      labels = _IgnorePairsWhere(~inputs.correspondences & drop, labels)
    return labels


def _BroadcastExamplePairLabelsToAllItemPairs(example_pair_labels, queries_shape, results_shape):
  """[Q, R] → `queries_shape + results_shape` (every item pair inherits its examples'
  label) (ref :222)."""
  assert example_pair_labels.dim() == 2
  q_extra, r_extra = len(queries_shape) - 1, len(results_shape) - 1
  view = example_pair_labels.reshape(
      (queries_shape[0],) + (1,) * q_extra + (results_shape[0],) + (1,) * r_extra)
  return view.expand(tuple(queries_shape) + tuple(results_shape))


class MultiItemExampleWrapper:
  """Lifts a single-item labeler to examples holding several items per modality
  (e.g. 5 captions per image) (ref :254). For intra-modal retrieval an item paired with
  itself is ignored."""

  def __init__(self, example_pair_labeler, modality_batch_shapes):
    self._labeler = example_pair_labeler
    self._shapes = dict(modality_batch_shapes)

  def __call__(self, inputs: ExamplePairs):
    ex = self._labeler(inputs)
    assert tuple(ex.shape) == (inputs.query_batch_size, inputs.result_batch_size)
    qs = utils.ResolveBatchDim(self._shapes[inputs.query_modality], inputs.query_batch_size)
    rs = utils.ResolveBatchDim(self._shapes[inputs.result_modality], inputs.result_batch_size)
    labels = _BroadcastExamplePairLabelsToAllItemPairs(ex, qs, rs).clone()
    if inputs.query_modality == inputs.result_modality:
      assert len(qs) == 2 and qs[1] > 1
      n = qs[1]
      self_pair = (inputs.correspondences[:, None, :, None] &
                   torch.eye(n, dtype=torch.bool, device=ex.device)[None, :, None, :])
      labels = _IgnorePairsWhere(self_pair, labels)
    return labels


def MultiLabelContrastiveLoss(labels, logits, axis: int = -1):
  """Softmax cross-entropy with (possibly several) positives per row (ref :328).

  loss_i = −log( Σ_{j∈pos(i)} e^{z_ij} / Σ_{j∉ignored(i)} e^{z_ij} ); rows without a
  positive contribute 0. `labels`: 1 positive, 0 negative, −1 ignore.
  """
  labels = labels.float()
  z = logits.float().masked_fill(labels == IGNORE_PAIR_LABEL, -1e30)
  log_den = torch.logsumexp(z, axis)
  log_num = torch.logsumexp(z.masked_fill(labels <= 0, -1e30), axis)
  has_pos = (labels > 0).any(axis)
  return torch.where(has_pos, log_den - log_num, torch.zeros_like(log_den))
