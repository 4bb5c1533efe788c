"""Pair labels and the multi-label contrastive loss (ref `lingvo/tasks/milan/labels.py`)."""

from __future__ import annotations

import torch

IGNORE_PAIR_LABEL = -1


class ExamplePairs:
  """Pairs of examples: in-batch all-pairs with a bool "should ignore" mask (ref :94)."""

  def __init__(self, query_examples, result_examples, labels=None):
    self.query_examples = query_examples
    self.result_examples = result_examples
    self.labels = labels

  @classmethod
  def WithinBatch(cls, batch, query_modality, result_modality):
    return cls(batch[query_modality], batch[result_modality])


class ExamplePairLabeler:
  """Labels pair (i, j) positive iff i == j, or via `positive_fn(batch)` (ref :175)."""

  def __init__(self, positive_fn=None):
    self._fn = positive_fn

  def __call__(self, batch_size, device, batch=None):
    if self._fn is not None and batch is not None:
      return self._fn(batch)
    return torch.eye(batch_size, device=device)


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
