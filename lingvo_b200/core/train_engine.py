"""The training fast path, shared by every entry point.

`runners.Trainer`, `program.TrainProgram`, `executor.ExecutorTpu` and
`bench.py` all step a task through one `TrainEngine`, so the benchmarked path
*is* the product path:

  * **data parallelism** — `parallel.dp.Attach(task)` when the process group has
    more than one rank (reference: towers + gradient aggregation inside
    `BaseTask.FProp/BProp`, `core/base_model.py:610-649,718-835`);
  * **device prefetch** — `DevicePrefetcher` (pinned host memory → device on a
    side stream; the infeed analogue, `base_input_generator.py:446-686`);
  * **whole-step CUDA graph** — `GraphedTrainStep` when `train.use_cuda_graph`
    allows it and the optimizer is capturable (the `tpu_steps_per_loop`
    on-device loop analogue, `runners.py:744-857`);
  * **checkpoint hooks** — `PreSave()` gathers sharded optimizer state,
    `PostRestore()` refreshes compute copies / fused-optimizer carries so a
    restored model continues exactly (stale Σw² carry).
"""

from __future__ import annotations

import logging
from typing import Optional

import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def _GraphMode(task) -> str:
  tp = task.params.train
  mode = tp.Get('use_cuda_graph') if 'use_cuda_graph' in tp else 'auto'
  if mode is True:
    return 'on'
  if mode is False or mode is None:
    return 'off'
  return str(mode)


def _AsyncMode(task) -> bool:
  """Cluster `mode == 'async'` requested explicitly (`--mode=async`); the default cluster
  params of a bare task count as sync."""
  try:
    return bool(task.cluster.params.mode == 'async' and
                getattr(task.cluster.params, 'job', '') in ('trainer', 'trainer_client',
                                                             'worker') and
                task.params.train.Get('async_data_parallel'))
  except Exception:  # pylint: disable=broad-except
    return False


class TrainEngine:
  """Steps `task` with DP sync, prefetch and (optionally) CUDA-graph replay."""

  def __init__(self, task, use_cuda_graph: Optional[str] = None,
               prefetch_depth: Optional[int] = None, attach_dp: bool = True):
    self.task = task
    self.device = task.Device()
    self._graph_mode = use_cuda_graph or _GraphMode(task)
    tp = task.params.train
    depth = prefetch_depth
    if depth is None:
      depth = tp.Get('device_prefetch_depth') if 'device_prefetch_depth' in tp else 2
    self.dp = None
    self.async_dp = None
    if attach_dp:
      from lingvo_b200.parallel import dp as dp_lib  # pylint: disable=g-import-not-at-top
      from lingvo_b200.parallel import mesh as mesh_lib  # pylint: disable=g-import-not-at-top
      if mesh_lib.Get().world > 1 and not getattr(task, '_dp_attached', False):
        if _AsyncMode(task):
          # `--mode=async`: replicas step independently and exchange parameters with
          # bounded staleness (parallel/async_dp.py) instead of synchronising gradients.
          from lingvo_b200.parallel import async_dp  # pylint: disable=g-import-not-at-top
          every = tp.Get('async_sync_every_n_steps') if 'async_sync_every_n_steps' in tp else 1
          self.async_dp = async_dp.Attach(task, sync_every=every)
          self._graph_mode = 'off' if self._graph_mode == 'auto' else self._graph_mode
        else:
          self.dp = dp_lib.Attach(task)
        task._dp_attached = True   # pylint: disable=protected-access
    self._prefetch = None
    self._prefetch_depth = depth
    self._graphed = None
    self._graph_failed = False
    self.h2d_bytes_last = 0

  # ------------------------------------------------------------------ input --
  @property
  def prefetcher(self):
    if self._prefetch is None:
      self._prefetch = base_input_generator.DevicePrefetcher(
          self.task.input, self.device, depth=max(1, self._prefetch_depth))
    return self._prefetch

  def NextBatch(self) -> NestedMap:
    """Next input batch, already on the device (copy overlapped on a side stream)."""
    b = self.prefetcher.Next()
    self.h2d_bytes_last = self.prefetcher.h2d_bytes_last
    return b

  # ------------------------------------------------------------------- step --
  @property
  def cuda_graph(self) -> bool:
    return self._graphed is not None

  @property
  def launches_per_step(self) -> int:
    return self._graphed.launches_per_step if self._graphed is not None else 0

  def _MaybeCapture(self, batch):
    if (self._graphed is not None or self._graph_failed or
        self._graph_mode == 'off' or self.device.type != 'cuda'):
      return
    task = self.task
    if len(getattr(task.cluster, 'available_devices', [[0]])) and (
        task.cluster.num_splits_per_client > 1):
      self._graph_failed = True     # multi-tower steps are not captured
      return
    if not all(getattr(l.optimizer, 'graph_capturable', False) for l in task.learners):
      if self._graph_mode == 'on':
        raise ValueError('train.use_cuda_graph=on but the optimizer is not capturable')
      self._graph_failed = True
      return
    from lingvo_b200.core import graph_step  # pylint: disable=g-import-not-at-top
    try:
      self._graphed = graph_step.GraphedTrainStep(task, batch, warmup=3)
    except Exception as e:  # pylint: disable=broad-except
      if self._graph_mode == 'on':
        raise
      logging.warning('CUDA-graph capture failed (%r); eager launches.', e)
      self._graph_failed = True

  def Step(self, batch: Optional[NestedMap] = None):
    """One train step → (eval_metrics, per_example); metrics stay on the device."""
    if batch is None:
      batch = self.NextBatch()
    self._MaybeCapture(batch)
    if self._graphed is not None:
      out = self._graphed(batch)
    else:
      out = self.task.TrainStep([batch] if not isinstance(batch, list) else batch)
    if self.async_dp is not None:
      self.async_dp.PostStep()
    return out

  # ------------------------------------------------------------ checkpoints --
  def PreSave(self):
    """Makes `var.data` / optimizer slots authoritative before a checkpoint is cut."""
    if self.async_dp is not None:
      self.async_dp.Finalize()       # replicas agree on what is written
    for lrn in self.task.learners:
      eng = getattr(lrn, 'fused_update', None)
      if eng is not None and hasattr(eng, 'PreSave'):
        eng.PreSave()

  def PostSave(self):
    for lrn in self.task.learners:
      eng = getattr(lrn, 'fused_update', None)
      if eng is not None and hasattr(eng, 'PostSave'):
        eng.PostSave()

  def PostRestore(self):
    """Re-derives every cached quantity from the restored master weights / slots."""
    PostRestore(self.task)


def PostRestore(task):
  """Hook run after any checkpoint load into `task` (also without an engine)."""
  py_utils.RefreshComputeCopies(task.vars.Flatten())
  try:
    from lingvo_b200.ops import optim as optim_ops  # pylint: disable=g-import-not-at-top
    optim_ops.Invalidate()
  except Exception:  # pylint: disable=broad-except
    pass
  for lrn in task.learners:
    eng = getattr(lrn, 'fused_update', None)
    if eng is not None and hasattr(eng, 'PostRestore'):
      eng.PostRestore()
