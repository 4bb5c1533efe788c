"""Summaries, step-rate tracking and model analysis.

Reference `lingvo/core/summary_utils.py`: `scalar/histogram/image/text` gated
by `cluster.add_summary` (:38-95), `StepRateTracker` (:393-429),
`ModelAnalysis` (:432-510). Here summaries are collected into a thread-local
`SummaryCollector` (host scalars) that runners flush to an `EventFileWriter`.
"""

from __future__ import annotations

import threading
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from lingvo_b200.core import cluster_factory


class _Collectors(threading.local):

  def __init__(self):
    super().__init__()
    self.stack: List['SummaryCollector'] = []


_COLLECTORS = _Collectors()


class SummaryCollector:
  """Context that captures summaries emitted by layers during FProp/BProp."""

  def __init__(self):
    self.scalars: Dict[str, Any] = {}
    self.histograms: Dict[str, Any] = {}
    self.texts: Dict[str, str] = {}
    self.images: Dict[str, Any] = {}

  def __enter__(self):
    _COLLECTORS.stack.append(self)
    return self

  def __exit__(self, *args):
    _COLLECTORS.stack.pop()

  def Resolve(self) -> Dict[str, float]:
    """Host-side float values (one sync for all device scalars)."""
    out = {}
    dev = {k: v for k, v in self.scalars.items() if isinstance(v, torch.Tensor)}
    if dev:
      stacked = torch.stack([v.detach().float().reshape(()) for v in dev.values()])
      for k, v in zip(dev.keys(), stacked.cpu().tolist()):
        out[k] = v
    for k, v in self.scalars.items():
      if k not in out:
        out[k] = float(v)
    return out

  def WriteTo(self, writer, step: int):
    vals = self.Resolve()
    if vals:
      writer.add_scalars(vals, step)
    for k, v in self.histograms.items():
      arr = v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else v
      writer.add_histogram(k, arr, step)
    for k, v in self.texts.items():
      writer.add_text(k, v, step)


def _Current() -> Optional[SummaryCollector]:
  if not _COLLECTORS.stack:
    return None
  if not cluster_factory.Current().add_summary:
    return None
  return _COLLECTORS.stack[-1]


def _ShouldAddSummary() -> bool:
  return _Current() is not None


def scalar(name: str, value, **kwargs):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.scalars[name] = value


def scalar_input_stats(*args, **kwargs):  # pylint: disable=invalid-name
  scalar(*args, **kwargs)


def histogram(name: str, tensor):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.histograms[name] = tensor


def text(name: str, value: str):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.texts[name] = value


def image(name: str, tensor, **kwargs):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.images[name] = tensor


def AddNormSummary(name: str, vs_gs) -> Tuple[torch.Tensor, torch.Tensor]:
  """Returns (var_norm, grad_norm) and emits both as scalars."""
  from lingvo_b200.core import py_utils
  leaves = [vg for vg in vs_gs.Flatten() if isinstance(vg, py_utils.VarGrad)]
  vn = torch.sqrt(py_utils.SumSquared([vg.var.detach() for vg in leaves]))
  gn = torch.sqrt(py_utils.SumSquared([vg.grad for vg in leaves]))
  scalar('var_norm/' + name, vn)
  scalar('grad_norm/' + name, gn)
  return vn, gn


def CollectVarHistogram(vs_gs):
  from lingvo_b200.core import py_utils
  for vg in vs_gs.Flatten():
    if isinstance(vg, py_utils.VarGrad):
      n = getattr(vg.var, 'var_name', 'var')
      histogram('var_hist/' + n, vg.var)
      histogram('grad_hist/' + n, vg.grad)


class StepRateTracker:
  """steps/sec & examples/sec over a sliding window (reference :393-429)."""

  def __init__(self):
    self._first_step = -1
    self._time_steps: List[Tuple[float, int, float]] = []

  def ComputeStepRate(self, current_steps: int, total_examples: float):
    if self._time_steps:
      total_examples += self._time_steps[-1][-1]
    else:
      self._first_step = current_steps
    self._time_steps.append((time.time(), current_steps, total_examples))
    # Keep a window of ~1000 steps.
    t_n, s_n, e_n = self._time_steps[-1]
    i = 0
    while i + 1 < len(self._time_steps) and s_n - self._time_steps[i + 1][1] > 1000:
      i += 1
    self._time_steps = self._time_steps[i:]
    (t0, s0, e0), (t1, s1, e1) = self._time_steps[0], self._time_steps[-1]
    rate = example_rate = 0.0
    if t1 > t0 + 1e-9:
      elapsed = t1 - t0
      rate = (s1 - s0) / elapsed
      example_rate = (e1 - e0) / elapsed
    return rate, example_rate, total_examples


def ModelAnalysis(model) -> Tuple[str, int]:
  """Table of variables (shape, size, dtype) + total (reference :432-510)."""
  rows = []
  total = 0
  for key, v in model.vars.FlattenItems():
    n = int(np.prod(v.shape)) if v.dim() else 1
    total += n
    rows.append((getattr(v, 'var_name', key), tuple(v.shape), n,
                 str(v.dtype).split('.')[-1]))
  w_name = max([len(r[0]) for r in rows] + [4])
  w_shape = max([len(str(r[1])) for r in rows] + [5])
  lines = ['%-*s  %-*s  %-12s %s' % (w_name, 'name', w_shape, 'shape', 'size',
                                    'dtype')]
  for r in rows:
    lines.append('%-*s  %-*s  %-12d %s' % (w_name, r[0], w_shape, str(r[1]),
                                          r[2], r[3]))
  lines.append('')
  lines.append('=' * 30)
  lines.append('total #params: %10d' % total)
  lines.append('')
  return '\n'.join(lines), total
