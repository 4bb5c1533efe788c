"""Jobs: Controller, Trainer, Evaler, Decoder.

Reference `lingvo/runners.py`: `Controller` (:70-185: init/restore/save,
`params.txt`, `model_analysis.txt`), `Trainer` (:192-360 hot loop),
`TrainerTpu` (:363-857), `Evaler` (:860-1100: poll ckpts → eval N samples →
`score-%08d.txt`), `Decoder` (:1105-1343: poll ckpts → decode →
`decoder_out_%09d`). `eager_runners.py` is the same control flow in eager
mode — which is the *only* mode here, so both surfaces map to these classes.

One process per GPU: the Trainer owns its variables (no parameter server), so
it also owns checkpointing; the Controller keeps the bookkeeping artefacts
(`control/params.txt`, `control/model_analysis.txt`).
"""

from __future__ import annotations

import logging
import os
import pickle
import time
from typing import Dict, Optional

import numpy as np
import torch

from lingvo_b200 import base_runner
from lingvo_b200.core import base_model
from lingvo_b200.core import checkpointer
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core import py_utils
from lingvo_b200.core import saver as saver_lib
from lingvo_b200.core import fault_injection
from lingvo_b200.core import summary_utils
from lingvo_b200.core import train_engine
from lingvo_b200.utils import tfevents


def _Rank() -> int:
  import torch.distributed as dist  # pylint: disable=g-import-not-at-top
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _WriteParamsFiles(params, out_dir: str, prefix: str = 'params'):
  os.makedirs(out_dir, exist_ok=True)
  with open(os.path.join(out_dir, prefix + '.txt'), 'w') as f:
    f.write(params.ToText())
  # params.pbtxt: the Hyperparam proto of the reference in protobuf text format;
  # params.pb: the same message serialized (round-trips through Params.FromProto).
  with open(os.path.join(out_dir, prefix + '.pbtxt'), 'w') as f:
    f.write(params.ToProtoText())
  with open(os.path.join(out_dir, prefix + '.pb'), 'wb') as f:
    f.write(params.ToProto())


def _MetricsToFloats(metrics: Dict) -> Dict[str, float]:
  """One device→host sync for all (value, weight) pairs."""
  names = list(metrics.keys())
  vals = []
  for k in names:
    v = metrics[k][0]
    vals.append(v.detach().float().reshape(()) if isinstance(v, torch.Tensor)
                else torch.tensor(float(v)))
  if not vals:
    return {}
  dev = None
  for v in vals:
    if v.is_cuda:
      dev = v.device
  stacked = torch.stack([v.to(dev) if dev is not None else v for v in vals])
  return dict(zip(names, stacked.cpu().tolist()))


class _CheckpointFollower(base_runner.BaseRunner):
  """A job that follows the trainer through its checkpoints and writes *training* summaries
  (the trainer-side counterpart of the Evaler, which follows them in eval mode).

  Every time a new checkpoint appears it is restored into this job's own copy of the model,
  one training batch goes through FProp + backward **without** an optimizer step, and the
  job writes: the task's training metrics, per-layer summaries collected during that eager
  step, global gradient / variable norms, `total_num_params`, and the trainer's progress
  (`global_step`, checkpoints seen). Runs on any device (a spare GPU or the host), never
  touches the trainer's step time.
  """

  _OUT_DIR = 'train_summaries'

  def __init__(self, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._out_dir = os.path.join(self._logdir, self._OUT_DIR)
    if self._model_task_name and self._OUT_DIR != 'control':
      self._out_dir += '_' + str(self._model_task_name)
    os.makedirs(self._out_dir, exist_ok=True)
    with self._cluster:
      self._model = self._params.Instantiate()
      self._model.to(py_utils.CurrentDevice())
    self._max_steps = min([self._params.train.max_steps] + [
        t.params.train.max_steps for t in self._model.tasks])
    self._checkpointer = checkpointer.Checkpointer(
        self._train_dir, self._model, train_params=self._params.train)
    self._summary_writer = tfevents.EventFileWriter(self._out_dir)
    self._model_analysis, self._total_num_params = summary_utils.ModelAnalysis(self._model)
    self._num_summaries = 0
    self._last_step = 0

  def Start(self):
    self._RunLoop(self._job_name, self._Loop)

  @property
  def num_summaries_written(self):
    return self._num_summaries

  def _TasksToSummarize(self):
    if self._model_task_name:
      return [self._model.GetTask(self._model_task_name)]
    return list(self._model.tasks)

  def SummarizeCheckpoint(self, path: str) -> int:
    """Restores `path` and writes one round of training summaries. Returns its step."""
    with self._cluster:
      self._checkpointer.RestoreFromPath(checkpoint_path=path)
      scalars = {}
      step = 0
      collector = summary_utils.SummaryCollector()
      for task in self._TasksToSummarize():
        prefix = '' if len(self._model.tasks) == 1 else task.params.name + '/'
        step = max(step, int(task.global_step))
        with collector, py_utils.GlobalStepContext(int(task.global_step)):
          metrics, _ = task.FPropDefaultTheta()
          loss = task.loss
          trainable = [v for v in task.vars.Flatten() if v.requires_grad]
          for v in trainable:
            v.grad = None
          if isinstance(loss, torch.Tensor) and loss.requires_grad:
            loss.backward()
        for k, v in _MetricsToFloats(metrics).items():
          scalars[prefix + k] = v
        gsq = sum(float(v.grad.detach().float().pow(2).sum()) for v in trainable
                  if v.grad is not None)
        vsq = sum(float(v.detach().float().pow(2).sum()) for v in trainable)
        scalars[prefix + 'grad_norm/all'] = gsq ** 0.5
        scalars[prefix + 'var_norm/all'] = vsq ** 0.5
        for v in trainable:
          v.grad = None
        task._metrics = None   # pylint: disable=protected-access  (no BProp follows)
      scalars['total_num_params'] = float(self._total_num_params)
      scalars['global_step'] = float(step)
      self._summary_writer.add_scalars(scalars, step)
      collector.WriteTo(self._summary_writer, step)
      self._summary_writer.flush()
      self._num_summaries += 1
      self._last_step = step
      self._SetStatusMessage('Write summary @%d' % step)
      return step

  def _Loop(self):
    tp = self._params.train
    next_summary_step = 1
    last_path = None
    while not self._should_stop.is_set():
      path = saver_lib.LatestCheckpoint(self._train_dir)
      if path and path != last_path:
        step = base_runner._StepOf(path)   # pylint: disable=protected-access
        if step >= next_summary_step or (self._max_steps is not None and
                                         step >= self._max_steps):
          self.SummarizeCheckpoint(path)
          next_summary_step = step + (tp.summary_interval_steps or 1)
        last_path = path
      step = base_runner._StepOf(last_path) if last_path else 0   # pylint: disable=protected-access
      if self._max_steps is not None and step >= self._max_steps:
        return
      if getattr(self, '_peer_done', None) is not None and self._peer_done():
        # the trainer finished: pick up its final checkpoint before leaving
        path = saver_lib.LatestCheckpoint(self._train_dir)
        if path and path != last_path:
          self.SummarizeCheckpoint(path)
        return
      time.sleep(self._poll_seconds)

  _poll_seconds = 0.2


class Controller(_CheckpointFollower):
  """The bookkeeping job of a training cluster (ref runners.py:70-189).

  The reference controller shares the trainer's variables through the parameter server and
  so can initialise / checkpoint them and evaluate the summary op against live weights. With
  one process per GPU the trainer owns its variables — and therefore checkpointing — and
  the controller keeps the other responsibilities: it writes the experiment artefacts
  (`control/params.txt`, `params.pbtxt`, `params.pb`, `model_analysis.txt`) and then follows
  the trainer through its checkpoints, writing training summaries and `total_num_params`
  to `control/` every `summary_interval_steps` (see `_CheckpointFollower`).
  """

  _OUT_DIR = 'control'

  def __init__(self, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._job_name = 'controller'
    self._control_dir = self._out_dir
    _WriteParamsFiles(self._params, self._control_dir)
    with open(os.path.join(self._control_dir, 'model_analysis.txt'), 'w') as f:
      f.write(self._model_analysis)
    self._summary_writer.add_text('model_analysis', self._model_analysis, 0)
    self._summary_writer.flush()


class Trainer(base_runner.BaseRunner):
  """Trains a model: the hot loop (reference runners.py:266-360).

  One process per GPU. With `WORLD_SIZE > 1` (torchrun) every rank runs this runner: the
  `TrainEngine` attaches data-parallel gradient sync (and the expert-parallel exchange
  the model asks for), replays the step from a CUDA graph when `train.use_cuda_graph`
  allows it, and feeds it from the device prefetcher. Rank 0 writes summaries and the
  replicated part of the checkpoint; every rank writes its expert / optimizer shards.
  """

  def __init__(self, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._job_name = 'trainer'
    self._rank = _Rank()
    with self._cluster:
      self._model = self._params.Instantiate()
      device = py_utils.CurrentDevice()
      self._model.to(device)
    self._task = self._GetTask()
    tp = self._params.train
    self._max_steps = min(tp.max_steps, self._task.params.train.max_steps)
    self._InitEarlyStop()
    self._step_rate_tracker = summary_utils.StepRateTracker()
    self._checkpointer = checkpointer.Checkpointer(
        self._train_dir, self._model, train_params=tp)
    self._engine = None
    self._summary_writer = None
    if self._rank == 0:
      _WriteParamsFiles(self._params, self._train_dir, 'trainer_params')
      self._summary_writer = tfevents.EventFileWriter(self._train_dir)
    self._done = False

  @property
  def engine(self) -> train_engine.TrainEngine:
    """The fast-path stepper (created after the first Restore)."""
    if self._engine is None:
      with self._cluster:
        self._engine = train_engine.TrainEngine(self._task)
      self._checkpointer.AttachEngine(self._engine)
    return self._engine

  @property
  def task(self):
    return self._task

  def Start(self):
    self._RunLoop('trainer', self._Loop)

  def done(self):  # pylint: disable=invalid-name
    return self._done

  def _Loop(self):
    tp = self._params.train
    with self._cluster:
      # Variables are made identical across ranks when the engine attaches DP (rank 0
      # wins), so restore first and attach afterwards.
      self._checkpointer.Restore()
      task = self._task
      engine = self.engine
      engine.PostRestore()
      global_step = task.global_step
      self._checkpointer.MaybeSave(gsteps=global_step)
      while True:
        if self._ShouldStop(step=global_step):
          break
        injector = fault_injection.Get()
        if injector is not None:
          injector.BeforeStep(global_step, self._train_dir)
        want_summary = (self._rank == 0 and tp.summary_interval_steps and
                        global_step % tp.summary_interval_steps == 0)
        # Layer-level summaries need the eager Python step; all other steps replay the graph.
        collector = summary_utils.SummaryCollector() if (
            want_summary and not engine.cuda_graph) else None
        if collector:
          with collector:
            eval_metrics, per_example = engine.Step()
        else:
          eval_metrics, per_example = engine.Step()
        global_step = task.global_step
        vals = _MetricsToFloats(eval_metrics)
        task.ProcessFPropResults(None, global_step, eval_metrics, per_example)
        n_ex = vals.get('num_samples_in_batch', 0.0)
        step_rate, example_rate, total_examples = (
            self._step_rate_tracker.ComputeStepRate(global_step, n_ex))
        msg = 'step:%6d, steps/sec: %0.2f, examples/sec: %0.2f' % (
            global_step, step_rate, example_rate)
        for key in sorted(vals):
          msg += ' %s:%.8g' % (key, vals[key])
        self._SetStatusMessage(msg)
        if want_summary:
          scalars = dict(vals)
          scalars['global_step/sec'] = step_rate
          scalars['examples/sec'] = example_rate
          scalars['total_samples'] = total_examples
          self._summary_writer.add_scalars(scalars, global_step)
          if collector:
            collector.WriteTo(self._summary_writer, global_step)
          self._summary_writer.flush()
        if self._trial.ShouldStopAndMaybeReport(global_step, vals):
          break
        self._checkpointer.MaybeSave(gsteps=global_step)
      # Always save the final state.
      self._checkpointer.Save(gsteps=global_step, sync=True)
      self._checkpointer.Sync()
      if self._summary_writer is not None:
        self._summary_writer.flush()
      self._done = True


TrainerTpu = Trainer


class TrainSummaries(_CheckpointFollower):
  """Writes training summaries from checkpoints into `train_summaries/` (reference
  eager_runners.py:118): restore the newest checkpoint, run the training graph once without
  applying gradients, write the summaries, wait for the next checkpoint."""

  def __init__(self, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._job_name = 'train_summaries'


class Evaler(base_runner.BaseRunner):
  """Evaluates checkpoints as they appear."""

  def __init__(self, eval_type: str, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._job_name = 'evaler_' + eval_type
    self._output_name = 'eval_' + eval_type
    self._eval_type = eval_type
    self._eval_dir = os.path.join(self._logdir, self._output_name)
    if self._model_task_name:
      self._eval_dir += '_' + str(self._model_task_name)
    os.makedirs(self._eval_dir, exist_ok=True)
    cp = self._params.cluster.Copy()
    cp.do_eval = True
    self._cluster = cluster_factory.Cluster(cp)
    with self._cluster:
      self._model = self._params.Instantiate()
      self._model.to(py_utils.CurrentDevice())
    self._task = self._GetTask()
    self._max_steps = min(self._params.train.max_steps,
                          self._task.params.train.max_steps)
    self._checkpointer = checkpointer.Checkpointer(
        self._train_dir, self._model, train_params=self._params.train)
    _WriteParamsFiles(self._params, self._eval_dir)
    self._summary_writer = tfevents.EventFileWriter(self._eval_dir)
    self._InitEarlyStop()

  def Start(self):
    self._RunLoop(self._job_name, self._Loop)

  def _Loop(self):
    ep = self._task.params.eval
    with self._cluster:
      if ep.load_checkpoint_from:
        self._EvalOnce(checkpointer.GetSpecificCheckpoint(
            ep.load_checkpoint_from))
      elif ep.eval_all_checkpoints:
        self._RunOnAllCheckpoints(self._EvalOnce, self._eval_dir)
      else:
        self._RunOnLatestCheckpoints(self._EvalOnce, self._eval_dir,
                                     ep.start_eval_after)

  def EvalLatestCheckpoint(self, last_path=None):
    path = saver_lib.LatestCheckpoint(self._train_dir)
    if not path:
      logging.info('No checkpoint available.')
      return None
    if path == last_path:
      return None
    with self._cluster:
      self._EvalOnce(path)
    return path

  def EvalCheckpoint(self, ckpt_id: int):
    with self._cluster:
      return self._EvalOnce(os.path.join(self._train_dir, 'ckpt-%08d' % ckpt_id))

  def _EvalOnce(self, path: str) -> bool:
    """Restores `path`, evaluates samples_per_summary examples (:980)."""
    task = self._task
    self._checkpointer.RestoreFromPath(checkpoint_path=path)
    global_step = task.global_step
    p = task.params
    samples = p.eval.samples_per_summary
    if p.input.eval_samples_per_summary is not None:
      samples = p.input.eval_samples_per_summary
    if samples == 0 and p.input.num_samples:
      samples = p.input.num_samples
    task.input.Reset()
    acc = metrics_lib.DeviceEvalMetrics()
    num_samples = 0
    while samples == 0 or num_samples < samples:
      try:
        with torch.no_grad():
          m, _ = task.EvalStep()
      except StopIteration:
        break
      acc.Update(m)
      ns = m.get('num_samples_in_batch')
      num_samples += int(float(ns[0])) if ns is not None else 1
    results = acc.Finalize()
    summaries = {k: v for k, (v, _) in results.items()}
    summaries['total_samples'] = num_samples
    self._WriteSummaries(
        self._summary_writer, os.path.basename(self._eval_dir), global_step,
        summaries, text_filename=os.path.join(
            self._eval_dir, 'score-{:08d}.txt'.format(global_step)))
    # Reference also dumps the eval graph pbtxt once; we dump the metric names.
    marker = os.path.join(self._eval_dir, self._output_name + '.pbtxt')
    if not os.path.exists(marker):
      with open(marker, 'w') as f:
        f.write('\n'.join('metric: "%s"' % k for k in sorted(summaries)) + '\n')
    should_stop = global_step >= (self._max_steps or 1 << 62)
    self._trial.ReportEvalMeasure(global_step, summaries, path)
    return should_stop or self._trial.ShouldStop()


def GetDecoderDir(logdir, decoder_type, model_task_name):
  if model_task_name:
    decoder_dir = '%s_%s' % (decoder_type, model_task_name)
  else:
    decoder_dir = decoder_type
  return os.path.join(logdir, decoder_dir)


class Decoder(base_runner.BaseRunner):
  """Decodes checkpoints as they appear."""

  def __init__(self, decoder_type: str, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._job_name = 'decoder_' + decoder_type
    self._decoder_dir = GetDecoderDir(self._logdir, self._job_name,
                                      self._model_task_name)
    os.makedirs(self._decoder_dir, exist_ok=True)
    cp = self._params.cluster.Copy()
    cp.do_eval = True
    self._cluster = cluster_factory.Cluster(cp)
    with self._cluster:
      self._model = self._params.Instantiate()
      self._model.to(py_utils.CurrentDevice())
    self._task = self._GetTask()
    self._max_steps = min(self._params.train.max_steps,
                          self._task.params.train.max_steps)
    self._checkpointer = checkpointer.Checkpointer(
        self._train_dir, self._model, train_params=self._params.train)
    _WriteParamsFiles(self._params, self._decoder_dir)
    self._summary_writer = tfevents.EventFileWriter(self._decoder_dir)

  def Start(self):
    self._RunLoop(self._job_name, self._Loop)

  def _Loop(self):
    ep = self._task.params.eval
    with self._cluster:
      if ep.load_checkpoint_from:
        self.DecodeCheckpoint(None, checkpointer.GetSpecificCheckpoint(
            ep.load_checkpoint_from))
      elif ep.decode_all_checkpoints:
        self._RunOnAllCheckpoints(lambda p: self.DecodeCheckpoint(None, p),
                                  self._decoder_dir)
      else:
        self._RunOnLatestCheckpoints(lambda p: self.DecodeCheckpoint(None, p),
                                     self._decoder_dir, ep.start_decoder_after)

  @classmethod
  def GetDecodeOutPath(cls, decoder_dir, checkpoint_id):
    return os.path.join(decoder_dir, 'decoder_out_%09d' % checkpoint_id)

  def GetDecoderDir(self):
    return self._decoder_dir

  def DecodeCheckpoint(self, sess, checkpoint_path: str) -> bool:
    """Decodes `samples_per_summary` examples with `checkpoint_path` (:1211)."""
    task = self._task
    p = task.params
    self._checkpointer.RestoreFromPath(checkpoint_path=checkpoint_path)
    global_step = task.global_step
    samples = p.eval.decoder_samples_per_summary
    if samples is None:
      samples = p.eval.samples_per_summary
    if p.input.decoder_samples_per_summary is not None:
      samples = p.input.decoder_samples_per_summary
    if samples == 0 and p.input.num_samples:
      samples = p.input.num_samples
    task.input.Reset()
    dec_metrics = task.CreateDecoderMetrics()
    if not dec_metrics:
      logging.info('Empty decoder metrics')
      return False
    buffered = []
    num = 0
    start = time.time()
    while samples == 0 or num < samples:
      try:
        dec_out = self._model.ConstructDecodeGraph(self._model_task_name)
      except StopIteration:
        break
      host = {}
      for k, v in dec_out.items():
        if isinstance(v, (tuple, list)):
          host[k] = tuple(x.detach().cpu().numpy() if isinstance(x, torch.Tensor)
                          else x for x in v)
        elif isinstance(v, torch.Tensor):
          host[k] = v.detach().cpu().numpy()
        else:
          host[k] = v
      post = task.PostProcessDecodeOut(host, dec_metrics)
      if post:
        buffered.extend(post)
      if 'num_samples_in_batch' in host:
        v = host['num_samples_in_batch']
        num += int(np.asarray(v[0] if isinstance(v, tuple) else v))
      else:
        num += 1
    summaries = {k: m.value for k, m in dec_metrics.items()}
    summaries['decode_secs'] = time.time() - start
    self._WriteSummaries(
        self._summary_writer, os.path.basename(self._decoder_dir), global_step,
        summaries, text_filename=os.path.join(
            self._decoder_dir, 'score-{:08d}.txt'.format(global_step)))
    out_path = self.GetDecodeOutPath(self._decoder_dir, global_step)
    with open(out_path, 'wb') as f:
      pickle.dump(buffered, f, protocol=pickle.HIGHEST_PROTOCOL)
    task.DecodeFinalize(base_model.DecodeFinalizeArgs(
        decode_out_path=out_path, decode_out=buffered))
    should_stop = global_step >= (self._max_steps or 1 << 62)
    self._trial.ReportEvalMeasure(global_step, summaries, checkpoint_path)
    return should_stop or self._trial.ShouldStop()

  def DecodeLatestCheckpoint(self, last_path=None):
    path = saver_lib.LatestCheckpoint(self._train_dir)
    if not path or path == last_path:
      return None
    with self._cluster:
      self.DecodeCheckpoint(None, path)
    return path
