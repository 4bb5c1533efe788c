"""Host-side evaluation metrics (reference `lingvo/core/metrics.py`).

`BaseMetric` / `AverageMetric` / `F1Metric` / `MCCMetric` / `CorpusBleuMetric`
/ `AUCMetric` / `CorrelationMetric` / `SamplingMetric` / `MultiClassAUCMetric`
/ `AverageKeyedCustomMetric` and a device-side accumulator
(`DeviceEvalMetrics`, the B200 analogue of `TpuEvalMetrics` :258-384: metric
(value, weight) pairs are accumulated on the GPU across loop steps and only
synced to host once per eval program).
"""

from __future__ import annotations

import collections
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def CreateScalarSummary(name: str, simple_value: float):
  from lingvo_b200.utils import tfevents
  return tfevents.ScalarValue(name, simple_value)


class BaseMetric:
  """Base class for aggregating statistics to compute a metric."""

  def Update(self, *args, **kwargs):
    raise NotImplementedError()

  @property
  def value(self):
    raise NotImplementedError()

  def Summary(self, name: str):
    return CreateScalarSummary(name, self.value)


class ConstantMetric(BaseMetric):

  def __init__(self):
    self._value = 0.0

  def Update(self, value):
    self._value = value

  @property
  def value(self):
    return self._value


class AverageMetric(BaseMetric):
  """Weighted average."""

  def __init__(self):
    self._total_value = 0.0
    self._total_weight = 0.0

  def Update(self, value, weight=1.0):
    if weight < 0.0:
      raise ValueError('weight must be non-negative, got %s' % weight)
    self._total_value += float(value) * float(weight)
    self._total_weight += float(weight)

  total_value = property(lambda self: self._total_value)
  total_weight = property(lambda self: self._total_weight)

  @property
  def value(self):
    return (self._total_value / self._total_weight
            if self._total_weight > 0 else 0.0)


class UniqueAverageMetric(AverageMetric):
  """Average over unique keys; asserts values are consistent per key."""

  def __init__(self, mismatch_is_error=True):
    super().__init__()
    self._map = {}
    self._mismatch_is_error = mismatch_is_error

  def Update(self, key, value, weight=1.0):
    if key in self._map:
      if self._mismatch_is_error and self._map[key] != (value, weight):
        raise ValueError('Conflicting value for key %s' % key)
      return
    self._map[key] = (value, weight)
    super().Update(value, weight)


class F1Metric(BaseMetric):

  def __init__(self):
    self._true_pos = 0.0
    self._false_pos = 0.0
    self._false_neg = 0.0

  def UpdateTruePositive(self, count=1.0):
    self._true_pos += count

  def UpdateFalsePositive(self, count=1.0):
    self._false_pos += count

  def UpdateFalseNegative(self, count=1.0):
    self._false_neg += count

  @property
  def value(self):
    if self._true_pos + self._false_pos > 0:
      precision = self._true_pos / (self._true_pos + self._false_pos)
    else:
      precision = 0.0
    if self._true_pos + self._false_neg > 0:
      recall = self._true_pos / (self._true_pos + self._false_neg)
    else:
      recall = 0.0
    if precision + recall > 0:
      return 2.0 * precision * recall / (precision + recall)
    return 0.0


class MCCMetric(F1Metric):
  """Matthews correlation coefficient."""

  def __init__(self):
    super().__init__()
    self._true_neg = 0.0

  def UpdateTrueNegative(self, count=1.0):
    self._true_neg += count

  @property
  def value(self):
    tp, fp, fn, tn = self._true_pos, self._false_pos, self._false_neg, self._true_neg
    num = tp * tn - fp * fn
    den = math.sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn))
    return num / den if den > 0 else 0.0


def _Ngrams(tokens: Sequence[str], n: int):
  return collections.Counter(
      tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


class CorpusBleuMetric(BaseMetric):
  """Corpus-level BLEU-4 with brevity penalty."""

  def __init__(self, separator_type=None, max_order: int = 4):
    self._max_order = max_order
    self._matches = [0] * max_order
    self._possible = [0] * max_order
    self._ref_len = 0
    self._hyp_len = 0
    self._separator_type = separator_type

  def _Split(self, s: str) -> List[str]:
    if self._separator_type == 'wpm':
      s = s.replace(' ', '').replace('▁', ' ')
    return s.split()

  def Update(self, ref_str: str, hyp_str: str):
    ref, hyp = self._Split(ref_str), self._Split(hyp_str)
    self._ref_len += len(ref)
    self._hyp_len += len(hyp)
    for n in range(1, self._max_order + 1):
      r, h = _Ngrams(ref, n), _Ngrams(hyp, n)
      self._matches[n - 1] += sum((r & h).values())
      self._possible[n - 1] += max(len(hyp) - n + 1, 0)

  @property
  def unsegmenter(self):
    return None

  @property
  def value(self):
    if self._hyp_len == 0 or min(self._possible) == 0:
      return 0.0
    precisions = []
    smooth = 1.0
    for m, p in zip(self._matches, self._possible):
      if m > 0:
        precisions.append(m / p)
      else:
        smooth *= 2
        precisions.append(1.0 / (smooth * p))
    geo = math.exp(sum(math.log(x) for x in precisions) / self._max_order)
    ratio = self._hyp_len / max(self._ref_len, 1)
    bp = 1.0 if ratio >= 1.0 else math.exp(1 - 1.0 / max(ratio, 1e-9))
    return geo * bp


class AUCMetric(BaseMetric):
  """ROC-AUC or PR-AUC over accumulated (label, prob[, weight])."""

  def __init__(self, mode='roc', samples=-1):
    assert mode in ('roc', 'pr')
    self._mode = mode
    self._label, self._prob, self._weight = [], [], []
    self._samples = samples

  def Update(self, label, prob, weight=None):
    self._label += list(label)
    self._prob += list(prob)
    self._weight += list(weight) if weight is not None else [1.0] * len(label)
    if self._samples > 0:
      self._label = self._label[-self._samples:]
      self._prob = self._prob[-self._samples:]
      self._weight = self._weight[-self._samples:]

  @property
  def value(self):
    y = np.asarray(self._label, dtype=np.float64)
    s = np.asarray(self._prob, dtype=np.float64)
    w = np.asarray(self._weight, dtype=np.float64)
    if y.size == 0 or (y * w).sum() == 0 or ((1 - y) * w).sum() == 0:
      return 0.0
    order = np.argsort(-s, kind='mergesort')
    y, s, w = y[order], s[order], w[order]
    tp = np.cumsum(y * w)
    fp = np.cumsum((1 - y) * w)
    last = np.r_[np.where(np.diff(s))[0], y.size - 1]
    tp, fp = tp[last], fp[last]
    if self._mode == 'roc':
      tpr = np.r_[0.0, tp / tp[-1]]
      fpr = np.r_[0.0, fp / fp[-1]]
      return float(np.trapz(tpr, fpr))
    precision = tp / np.maximum(tp + fp, 1e-12)
    recall = tp / tp[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


class MultiClassAUCMetric(BaseMetric):

  def __init__(self, num_classes, mode='roc', samples=-1):
    self._metrics = [AUCMetric(mode, samples) for _ in range(num_classes)]

  def Update(self, labels, probs, weights=None):
    labels, probs = np.asarray(labels), np.asarray(probs)
    for c, m in enumerate(self._metrics):
      m.Update(labels[:, c], probs[:, c],
               None if weights is None else np.asarray(weights)[:, c])

  @property
  def value(self):
    return float(np.mean([m.value for m in self._metrics]))


class CorrelationMetric(BaseMetric):

  def __init__(self, mode='pearson'):
    assert mode in ('pearson', 'spearman')
    self._mode = mode
    self._t, self._p = [], []

  def Update(self, target, pred):
    self._t += list(target)
    self._p += list(pred)

  @property
  def value(self):
    t, p = np.asarray(self._t, np.float64), np.asarray(self._p, np.float64)
    if t.size < 2:
      return 0.0
    if self._mode == 'spearman':
      t = np.argsort(np.argsort(t)).astype(np.float64)
      p = np.argsort(np.argsort(p)).astype(np.float64)
    c = np.corrcoef(t, p)[0, 1]
    return float(0.0 if np.isnan(c) else c)


class SamplingMetric(BaseMetric):
  """Keeps a uniform sample of `num_samples` updates (reservoir)."""

  def __init__(self, num_samples):
    self._num_samples = num_samples
    self._samples = []
    self._num_seen = 0
    self._rng = np.random.RandomState(0)

  def Update(self, *args, **kwargs):
    sample = (args, kwargs)
    self._num_seen += 1
    if len(self._samples) < self._num_samples:
      self._samples.append(sample)
    else:
      i = self._rng.randint(0, self._num_seen)
      if i < self._num_samples:
        self._samples[i] = sample

  samples = property(lambda self: self._samples)

  @property
  def value(self):
    return 0

  def Summary(self, name):
    return self._CreateSummary(name)

  def _CreateSummary(self, name):
    return CreateScalarSummary(name, float(len(self._samples)))


class AverageKeyedCustomMetric(BaseMetric):
  """Groups values by key; value = mean over keys of fn(values)."""

  def __init__(self, fn=np.mean):
    self._fn = fn
    self._vals = collections.defaultdict(list)

  def Update(self, key, value):
    self._vals[key].append(value)

  @property
  def value(self):
    if not self._vals:
      return 0.0
    return float(np.mean([self._fn(v) for v in self._vals.values()]))


class GroupPairAUCMetric(BaseMetric):
  """Pairwise AUC within groups."""

  def __init__(self):
    self._groups = collections.defaultdict(list)

  def UpdateRaw(self, group_ids, target, logits, weight=None):
    for g, t, l in zip(group_ids, target, logits):
      self._groups[g].append((t, l))

  @property
  def value(self):
    good = total = 0.0
    for items in self._groups.values():
      for i in range(len(items)):
        for j in range(len(items)):
          if items[i][0] > items[j][0]:
            total += 1
            good += 1.0 if items[i][1] > items[j][1] else (
                0.5 if items[i][1] == items[j][1] else 0.0)
    return good / total if total else 0.0


class DeviceEvalMetrics:
  """Accumulates {name: (value, weight)} on device across steps.

  Equivalent of reference `TpuEvalMetrics` (:258-384): one fused
  `value·weight` / `weight` accumulator tensor pair, finalised (and optionally
  all-reduced across ranks) once per program run.
  """

  def __init__(self):
    self._names: List[str] = []
    self._acc: Optional[torch.Tensor] = None  # [n, 2] (sum v·w, sum w)

  def Update(self, metrics: Dict[str, Tuple[torch.Tensor, torch.Tensor]]):
    if not self._names:
      self._names = list(metrics.keys())
    vals, wts = [], []
    dev = None
    for k in self._names:
      v, w = metrics[k]
      v = torch.as_tensor(v, dtype=torch.float32)
      w = torch.as_tensor(w, dtype=torch.float32)
      if v.is_cuda:
        dev = v.device
      vals.append(v.detach().reshape(()))
      wts.append(w.detach().reshape(()))
    if dev is not None:
      vals = [v.to(dev) for v in vals]
      wts = [w.to(dev) for w in wts]
    v, w = torch.stack(vals).float(), torch.stack(wts).float()
    cur = torch.stack([v * w, w], dim=1)
    self._acc = cur if self._acc is None else self._acc + cur

  def AllReduce(self, group=None):
    import torch.distributed as dist
    if self._acc is not None and dist.is_available() and dist.is_initialized():
      dist.all_reduce(self._acc, group=group)

  def Finalize(self) -> Dict[str, Tuple[float, float]]:
    if self._acc is None:
      return {}
    acc = self._acc.cpu()
    out = {}
    for i, k in enumerate(self._names):
      s, w = float(acc[i, 0]), float(acc[i, 1])
      out[k] = (s / w if w > 0 else 0.0, w)
    return out

  def Reset(self):
    self._acc = None


TpuEvalMetrics = DeviceEvalMetrics
