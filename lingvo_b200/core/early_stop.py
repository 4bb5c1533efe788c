"""Metric history + early stopping (reference `core/early_stop.py:25-199`).

`MetricHistory` appends `step score` rows to a tsv the evaler writes and the
trainer reads; `EarlyStop.Stop(step)` checks window/tolerance via `BestStep`
(native op `best_step_op_kernels.cc` in the reference; C++ in
`ops/csrc/native_text.cpp` here, python fallback below).
"""

import os
from typing import Tuple

from lingvo_b200.core import hyperparams


def BestStep(hist_file: str, tol: float = 0.0, minimize: bool = True,
             metric: str = '') -> Tuple[int, int]:
  """Returns (best_step, last_step) from a `step score` text history."""
  best_step, last_step = 0, 0
  best_val = None
  if not hist_file or not os.path.exists(hist_file):
    return 0, 0
  if hist_file.endswith('.tsv') or not os.path.isdir(hist_file):
    with open(hist_file) as f:
      for line in f:
        parts = line.split()
        if len(parts) < 2:
          continue
        step, val = int(float(parts[0])), float(parts[1])
        if not minimize:
          val = -val
        if best_val is None or val < best_val - tol:
          best_val, best_step = val, step
        last_step = step
  return best_step, last_step


class MetricHistory:
  """Records a metric over time to a file shared between jobs."""

  @classmethod
  def Params(cls):
    p = hyperparams.InstantiableParams(cls)
    p.Define('jobname', 'eval_test', 'Job that evaluates the metric.')
    p.Define('metric', 'log_pplx', 'Metric to monitor.')
    p.Define('minimize', True, 'True if lower is better.')
    p.Define('logdir', '', 'Root dir for the experiment.')
    p.Define('tfevent_file', False, 'Kept for parity.')
    p.Define('local_filesystem', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    self.params = params.Copy()
    self._hist_file = None
    self._minimize = params.minimize

  def _Key(self, jobname, metric):
    return jobname + '.' + metric

  @property
  def hist_file(self):
    p = self.params
    if self._hist_file is None:
      from lingvo_b200.core import cluster_factory
      logdir = p.logdir or cluster_factory.Current().logdir
      self._hist_file = os.path.join(
          logdir, 'history_' + self._Key(p.jobname, p.metric) + '.txt')
    return self._hist_file

  @property
  def minimize(self):
    return self._minimize

  @property
  def metric(self):
    return self.params.metric

  @property
  def tfevent_file(self):
    return self.params.tfevent_file if 'tfevent_file' in self.params else False

  @staticmethod
  def SetLogdirInMetricHistories(params, logdir):
    """Points every MetricHistory params found anywhere under `params` at `logdir` (ref
    :100)."""
    def _Visit(p):
      for _, v in p.IterParams():
        if isinstance(v, hyperparams.Params):
          if isinstance(getattr(v, 'cls', None), type) and issubclass(v.cls, MetricHistory):
            v.logdir = logdir
          _Visit(v)
        elif isinstance(v, (list, tuple)):
          for x in v:
            if isinstance(x, hyperparams.Params):
              _Visit(x)
    _Visit(params)

  def Append(self, step, value):
    """Unconditionally records (step, value)."""
    os.makedirs(os.path.dirname(self.hist_file) or '.', exist_ok=True)
    with open(self.hist_file, 'a') as f:
      f.write('%d %f\n' % (step, value))

  def ConditionalAppend(self, jobname, metric, step, value) -> bool:
    p = self.params
    if jobname == p.jobname and metric == p.metric:
      os.makedirs(os.path.dirname(self.hist_file) or '.', exist_ok=True)
      with open(self.hist_file, 'a') as f:
        f.write('%d %f\n' % (step, value))
      return True
    return False


class EarlyStop:
  """Stops when the metric has not improved within `window` steps."""

  @classmethod
  def Params(cls):
    p = hyperparams.InstantiableParams(cls)
    p.Define('name', 'early_stop', 'Early stop name.')
    p.Define('metric_history', MetricHistory.Params(), 'Which metric to use.')
    p.Define('tolerance', 0.0, 'Minimum significant difference.')
    p.Define('window', 0, 'Max steps without improvement; 0 disables.')
    p.Define('verbose', True, 'Log early-stop checks.')
    p.Define('min_steps', 0, 'Minimum training steps.')
    return p

  def __init__(self, params):
    self.params = params.Copy()
    self._metric_history = None
    if self.params.window:
      self._metric_history = MetricHistory(self.params.metric_history)
    self.best_step = 0
    self.last_step = 0

  @property
  def metric_history(self):
    return self._metric_history

  def FProp(self, theta=None):
    return None

  def Stop(self, session=None) -> bool:
    p = self.params
    if not self._metric_history:
      return False
    self.best_step, self.last_step = BestStep(
        self._metric_history.hist_file, p.tolerance,
        self._metric_history.minimize)
    s = self.last_step - self.best_step > p.window and (
        self.last_step >= p.min_steps)
    if p.verbose:
      import logging
      logging.info('early stop: best step=%d last step=%d stop=%s',
                   self.best_step, self.last_step, s)
    return s
