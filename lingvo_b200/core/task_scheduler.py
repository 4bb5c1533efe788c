"""Multi-task samplers (reference `core/task_scheduler.py:23-368`).

All samplers expose `Sample(step) -> task_name` and `cur_probs`. In a
multi-rank job every rank must draw the same task, so sampling uses a
per-scheduler `numpy` generator seeded from `random_seed` ⊕ step instead of
the process-global RNG.
"""

import os

import numpy as np

from lingvo_b200.core import base_layer
from lingvo_b200.core import early_stop


class TaskScheduler(base_layer.BaseLayer):
  """Generic multi-task scheduler."""

  def __init__(self, params):
    super().__init__(params)
    self.cur_probs = None
    self.SetVariableFree = lambda *_: None

  def _Choice(self, tasks, probs, step):
    seed = self.params.random_seed
    if seed is None:
      return str(np.random.choice(tasks, p=probs))
    rng = np.random.RandomState((int(seed) * 1000003 + int(step)) % (2**31 - 1))
    return str(rng.choice(tasks, p=probs))

  def Sample(self, current_step):
    raise NotImplementedError('Abstract method')

  def FProp(self, theta, current_step):
    return self.Sample(current_step)


class AdaptiveScheduler(TaskScheduler):
  """Two-task scheduler driven by dev metric histories."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('tasks', [], 'List of tasks')
    p.Define('expected', [], 'List of final expected scores')
    p.Define('mh_a', early_stop.MetricHistory.Params(), '')
    p.Define('mh_b', early_stop.MetricHistory.Params(), '')
    p.Define('epsilon', 0.05, 'Regularization toward uniform.')
    p.Define('alpha', 1.0, 'Normalized task scores are raised to this power.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if len(p.tasks) != 2 or len(p.expected) != 2:
      raise ValueError('Only two tasks are supported by this scheduler.')
    if p.epsilon < 0:
      raise ValueError('Epsilon should be positive.')
    self.tasks = p.tasks
    self.last_scores = [0.0] * 2
    self._metric_histories = [early_stop.MetricHistory(p.mh_a),
                              early_stop.MetricHistory(p.mh_b)]

  def getMetricHistories(self):  # pylint: disable=invalid-name
    for i, mh in enumerate(self._metric_histories):
      score = 0.0
      if os.path.exists(mh.hist_file):
        with open(mh.hist_file) as f:
          lines = f.readlines()
        if lines:
          try:
            score = float(lines[-1].split()[-1])
          except (IndexError, ValueError):
            score = 0.0
      self.last_scores[i] = score


class SimpleAdaptiveScheduler(AdaptiveScheduler):
  """p ∝ 1 + ε − min(1, score/expected)^α."""

  def Sample(self, current_step):
    self.getMetricHistories()
    p = self.params
    probs = np.array([1 + p.epsilon - min(1, s / p.expected[i])**p.alpha
                      for i, s in enumerate(self.last_scores)])
    probs = tuple(probs / probs.sum())
    self.cur_probs = probs
    return self._Choice(p.tasks, probs, current_step)


class InverseRatioAdaptiveScheduler(AdaptiveScheduler):
  """p ∝ 1 / (min(1, score/expected)^α + ε)."""

  def Sample(self, current_step):
    self.getMetricHistories()
    p = self.params
    probs = np.array([1.0 / (min(1, s / p.expected[i])**p.alpha + p.epsilon)
                      for i, s in enumerate(self.last_scores)])
    probs = tuple(probs / probs.sum())
    self.cur_probs = probs
    return self._Choice(p.tasks, probs, current_step)


class ShiftedExponentialScheduler(TaskScheduler):
  """Unnormalised score a + b·exp(−α·t) per task."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('alpha', 0, 'Rate at which the schedule changes.')
    p.Define('task_probs', [], 'List of (task, prob | (init, final)).')
    return p

  def __init__(self, params):
    super().__init__(params)
    assert isinstance(self.params.task_probs, list)
    self.tasks = []
    self._descriptors = []

  def Sample(self, current_step):
    probs = np.array([a + b * np.exp(-self.params.alpha * current_step)
                      for a, b in self._descriptors], dtype=np.float64)
    probs = tuple(probs / probs.sum())
    self.cur_probs = probs
    return self._Choice(self.tasks, probs, current_step)


class ConstantScheduler(ShiftedExponentialScheduler):

  def __init__(self, params):
    super().__init__(params)
    for key, value in self.params.task_probs:
      self.tasks.append(key)
      self._descriptors.append((value, 0))


class ExponentialScheduler(ShiftedExponentialScheduler):

  def __init__(self, params):
    super().__init__(params)
    for key, value in self.params.task_probs:
      self.tasks.append(key)
      self._descriptors.append((value[1], value[0] - value[1]))


class SigmoidScheduler(ShiftedExponentialScheduler):

  def __init__(self, params):
    super().__init__(params)
    for key, value in self.params.task_probs:
      self.tasks.append(key)
      self._descriptors.append((value[1], 2 * value[0] - value[1]))


class RoundRobinScheduler(TaskScheduler):
  """Deterministic sequential schedule."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('tasks', [], 'List of task names. No repetitions allowed.')
    return p

  def __init__(self, params):
    super().__init__(params)
    assert isinstance(self.params.tasks, list)
    self.tasks = sorted(self.params.tasks)
    self.n_tasks = len(self.tasks)
    self.cur_probs = [1. / self.n_tasks] * self.n_tasks
    self.next_task_idx = 0

  def Sample(self, current_step):
    name = self.tasks[self.next_task_idx]
    self.next_task_idx = (self.next_task_idx + 1) % self.n_tasks
    return name


class SequentialScheduler(TaskScheduler):
  """Stays a fixed number of steps on each task, in order."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('task_steps', [], 'List of (task_name, steps_for_task).')
    return p

  def __init__(self, params):
    super().__init__(params)
    assert isinstance(self.params.task_steps, list) and self.params.task_steps
    self.task_steps = []
    for name, steps in self.params.task_steps:
      assert steps > 0
      prev = self.task_steps[-1][1] if self.task_steps else 0
      self.task_steps.append((name, steps + prev))
    self.n_tasks = len(self.task_steps)
    self.task_idx = 0
    self.cur_probs = [1] + [0] * (self.n_tasks - 1)

  def Sample(self, current_step):
    name, to_step = self.task_steps[self.task_idx]
    if current_step >= to_step and self.task_idx < self.n_tasks - 1:
      self.task_idx += 1
      name = self.task_steps[self.task_idx][0]
      self.cur_probs[self.task_idx - 1] = 0
      self.cur_probs[self.task_idx] = 1
    return name


class PieceWiseScheduler(TaskScheduler):
  """Chains schedulers, each for a number of steps."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('schedule_steps', [], 'List of (scheduler params, num steps).')
    return p

  def __init__(self, params):
    super().__init__(params)
    assert isinstance(self.params.schedule_steps, list)
    self.schedule_steps = []
    sub = []
    for cls_params, steps in self.params.schedule_steps:
      prev = self.schedule_steps[-1] if self.schedule_steps else 0
      self.schedule_steps.append(steps + prev)
      sub.append(cls_params)
    self.CreateChildren('schedules', sub)
    self.n_schedules = len(self.schedule_steps)
    self.schedule_idx = 0
    self.task_step_offset = 0
    self.cur_probs = self.schedules[0].cur_probs

  def Sample(self, current_step):
    to_step = self.schedule_steps[self.schedule_idx]
    if current_step >= to_step and self.schedule_idx < self.n_schedules - 1:
      self.task_step_offset = to_step
      self.schedule_idx += 1
    cur = self.schedules[self.schedule_idx]
    name = cur.Sample(current_step - self.task_step_offset)
    self.cur_probs = cur.cur_probs
    return name
