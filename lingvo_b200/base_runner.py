"""Base class of all jobs (controller / trainer / evaler / decoder / executor).

Reference `lingvo/base_runner.py`: owns cluster, global step, early stop,
summary writer; **retry/fault policy** `_RunLoop` (:398-527, `@Retry(
max_retries=20)`, fatal vs retryable classification); checkpoint polling
(`_FindNewCheckpoint :222-235`, `_RunOnLatestCheckpoints :259`,
`_RunOnAllCheckpoints :297`); `_WriteSummaries :653`; score files (:668-693).

Exceptions map: torch/NCCL/IO errors that indicate a transient peer or
filesystem condition are retryable; shape/type/assertion errors are fatal.
"""

from __future__ import annotations

import logging
import os
import threading
import time
import traceback
from typing import Callable, Dict, List, Optional

import torch

from lingvo_b200 import base_trial
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import early_stop
from lingvo_b200.core import py_utils
from lingvo_b200.core import saver as saver_lib
from lingvo_b200.utils import tfevents

RETRYABLE = (ConnectionError, TimeoutError, BrokenPipeError, EOFError,
             BlockingIOError, InterruptedError)
FATAL = (AssertionError, ValueError, TypeError, KeyError, AttributeError,
         NotImplementedError, FloatingPointError)


class BaseRunner:
  """Base class for all jobs."""

  def __init__(self, params, model_task_name: str, logdir: str, tf_master: str = '',
               trial: base_trial.Trial = None):
    p = params.Copy()
    self._logdir = logdir
    self._model_task_name = model_task_name
    self._tf_master = tf_master
    self._trial = trial or base_trial.NoOpTrial()
    self._params = self._trial.OverrideModelParams(p)
    self._cluster = cluster_factory.Cluster(self._params.cluster)
    self._train_dir = os.path.join(self._logdir, 'train')
    os.makedirs(self._train_dir, exist_ok=True)
    self._max_steps = None
    self._status_msg_fn = None
    self._model = None
    self._summary_writer = None
    self._should_stop = threading.Event()
    self._early_stop = None
    self._daemon = False
    self._job_name = 'runner'

  # ------------------------------------------------------------ properties --
  @property
  def params(self):
    return self._params

  @property
  def cluster(self):
    return self._cluster

  @property
  def model(self):
    return self._model

  def _GetTask(self):
    return self._model.GetTask(self._model_task_name) if self._model_task_name \
        else self._model.tasks[0]

  def _SetStatusMessage(self, message, retrying=False):
    logging.info('%s', message)
    if self._status_msg_fn:
      self._status_msg_fn(message)

  def _InitEarlyStop(self):
    tp = self._params.train
    if tp.early_stop is not None and tp.early_stop.window:
      esp = tp.early_stop.Copy()
      if not esp.metric_history.logdir:
        esp.metric_history.logdir = self._logdir
      self._early_stop = early_stop.EarlyStop(esp)

  def _ShouldStop(self, sess=None, step=None, check_early_stop=True) -> bool:
    """max_steps reached, trial says stop, or early stop (:181-220)."""
    if self._should_stop.is_set():
      return True
    if step is None:
      step = self._GetTask().global_step if self._model is not None else 0
    if self._trial.ShouldStop():
      logging.info('Training skipped (trial requested to stop).')
      return True
    if check_early_stop and self._early_stop is not None and (
        self._early_stop.Stop()):
      logging.info('Training stopped early (best step %d, last step %d).',
                   self._early_stop.best_step, self._early_stop.last_step)
      return True
    if self._max_steps is not None and step >= self._max_steps:
      logging.info('Training finished: step %d >= max_steps %d', step,
                   self._max_steps)
      return True
    return False

  def RequestStop(self):
    self._should_stop.set()

  # -------------------------------------------------------------- summaries --
  def _SummaryWriter(self, subdir: str):
    if self._summary_writer is None:
      self._summary_writer = tfevents.EventFileWriter(
          os.path.join(self._logdir, subdir))
    return self._summary_writer

  def _SummarizeValue(self, steps: int, tag: str, value: float, writer=None):
    (writer or self._summary_writer).add_scalar(tag, float(value), steps)

  def _WriteSummaries(self, writer, job_name: str, global_step: int,
                      summaries: Dict[str, float], text_filename: str = None):
    """Event file + `tag: value` text lines (`score-%08d.txt`) (:653-693)."""
    status = ['%s: step:%6d' % (job_name, global_step)]
    lines = []
    for tag in sorted(summaries):
      val = float(summaries[tag])
      writer.add_scalar(tag, val, global_step)
      status.append('%s:%.8g' % (tag, val))
      lines.append('%s: %s\n' % (tag, val))
      if self._early_stop is not None and self._early_stop.metric_history:
        self._early_stop.metric_history.ConditionalAppend(
            job_name, tag, global_step, val)
    writer.flush()
    self._SetStatusMessage(', '.join(status))
    if text_filename is not None:
      with open(text_filename, 'w') as f:
        f.write(''.join(lines))

  # ------------------------------------------------------ checkpoint polling --
  def _FindNewCheckpoint(self, prev_path: Optional[str], timeout_s: float = None,
                         poll_s: float = 1.0) -> Optional[str]:
    """Blocks until a checkpoint newer than `prev_path` appears (:222-235)."""
    t0 = time.time()
    while not self._should_stop.is_set():
      path = saver_lib.LatestCheckpoint(self._train_dir)
      if path and path != prev_path:
        return path
      if timeout_s is not None and time.time() - t0 > timeout_s:
        return None
      time.sleep(poll_s)
    return None

  def _RunOnLatestCheckpoints(self, runner_fn: Callable[[str], bool],
                              runner_dir: str, start_after: int = 0):
    """Runs `runner_fn(ckpt)` on each newly appearing latest ckpt (:259)."""
    path = None
    processed = set(py_utils.GetProcessedCheckpoints(runner_dir))
    while not self._should_stop.is_set():
      path = self._FindNewCheckpoint(path, timeout_s=self._poll_timeout_s)
      if path is None:
        return
      step = _StepOf(path)
      if step < start_after or path in processed:
        if self._max_steps is not None and step >= self._max_steps:
          return
        continue
      done = runner_fn(path)
      py_utils.UpdateProcessedCheckpoints(runner_dir, path)
      processed.add(path)
      if done or (self._max_steps is not None and step >= self._max_steps):
        return

  def _RunOnAllCheckpoints(self, runner_fn: Callable[[str], bool],
                           runner_dir: str):
    """Every checkpoint exactly once, in order (:297)."""
    processed = set(py_utils.GetProcessedCheckpoints(runner_dir))
    while not self._should_stop.is_set():
      pending = [p for p in saver_lib.AllCheckpoints(self._train_dir)
                 if p not in processed]
      if not pending:
        latest = saver_lib.LatestCheckpoint(self._train_dir)
        if latest and self._max_steps is not None and (
            _StepOf(latest) >= self._max_steps):
          return
        time.sleep(1.0)
        continue
      for path in pending:
        done = runner_fn(path)
        py_utils.UpdateProcessedCheckpoints(runner_dir, path)
        processed.add(path)
        if done:
          return

  _poll_timeout_s = None

  # ---------------------------------------------------------------- run loop --
  def _RunLoop(self, job_name: str, loop_func: Callable, loop_args=(),
               cleanup_func: Callable = None, max_retries: int = 20):
    """Fault policy: retry transient errors, fail fast on fatal ones (:398)."""
    retries = 0
    delay = 1.0
    while True:
      try:
        logging.info('%s started.', job_name)
        loop_func(*loop_args)
        logging.info('%s done.', job_name)
        if self._daemon:
          # In daemon mode an external scheduler restarts us.
          logging.info('%s: daemon mode exit.', job_name)
        return
      except base_trial.TunerManagedError:
        raise
      except Exception as e:  # pylint: disable=broad-except
        retryable = isinstance(e, RETRYABLE) or (
            isinstance(e, RuntimeError) and any(
                s in str(e) for s in ('NCCL', 'Connection', 'timed out',
                                      'unhandled system error')))
        if isinstance(e, FATAL) or not retryable:
          msg = '%s failed: %s\n%s' % (job_name, e, traceback.format_exc())
          self._SetStatusMessage(msg)
          self._trial.ReportDone(infeasible=True, infeasible_reason=str(e))
          raise
        retries += 1
        if retries > max_retries:
          raise
        self._SetStatusMessage('%s exception (retry %d): %s' %
                               (job_name, retries, e), retrying=True)
        if cleanup_func is not None:
          cleanup_func()
        time.sleep(delay)
        delay = min(delay * 1.5, 60.0)

  def Start(self):
    raise NotImplementedError('Abstract method')

  def StartEnqueueOp(self, op):
    return None


def _StepOf(ckpt_path: str) -> int:
  try:
    return int(ckpt_path.rsplit('-', 1)[-1])
  except ValueError:
    return 0
