"""Tasks and models: the train/eval/decode step contract.

Reference `lingvo/core/base_model.py`: `BaseTask.Params` (:124-335), `FProp`
→ towers → `_FPropResult` (:570-693), `BProp` → `_BPropGenTrainOps`
(:718-835, order: learners → BN/moving-average updates →
`PostTrainingStepUpdate` → EMA → `PostEmaUpdate` → mask update →
`global_step += 1`), EMA (:859-915), decode API (:918-1014), `BaseModel`
(:1138), `SingleTaskModel` (:1379), `MultiTaskModel` (:1480-1640).

PyTorch-first: a train step is an eager `FProp` (autograd tape) followed by
`BProp` (learner → optimizer, in place). There is no graph construction —
`ConstructFPropBPropGraph()` is kept as an alias that runs one step so runner
code written against the reference keeps its shape. The global step is a host
integer (checkpointed as `global_step`).
"""

from __future__ import annotations

import dataclasses

import collections
import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import input_policy
from lingvo_b200.core import base_layer
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import early_stop
from lingvo_b200.core import hyperparams
from lingvo_b200.core import learner as learner_lib
from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core import schedule
from lingvo_b200.core import summary_utils
from lingvo_b200.core import task_scheduler
from lingvo_b200.core.nested_map import NestedMap


class DecodeFinalizeArgs(
    collections.namedtuple('DecodeFinalizeArgs',
                           ['decode_out_path', 'decode_out'])):
  """Arguments to BaseTask.DecodeFinalize()."""


def _VarByName(layer) -> Dict[str, Tuple[base_layer.BaseLayer, str, torch.nn.Parameter]]:
  out = {}
  for _, sub in layer.Walk():
    for k, v in sub._private_vars.items():  # pylint: disable=protected-access
      out[v.var_name] = (sub, k, v)
  return out


def DecodeOutAsTensors(dec_out):
  """Decode outputs reach `PostProcessDecodeOut` either as device/CPU tensors (direct
  calls) or as the host numpy dict the runners / programs hand over: normalise to a
  NestedMap of CPU tensors so task code can use tensor ops in both cases."""
  def conv(x):
    if isinstance(x, np.ndarray) and x.dtype.kind not in 'OUS':
      return torch.from_numpy(np.ascontiguousarray(x))
    if isinstance(x, torch.Tensor):
      return x.detach().cpu()
    if isinstance(x, (tuple, list)):
      return type(x)(conv(v) for v in x)
    return x
  out = NestedMap()
  for k, v in dict(dec_out).items():
    out[k] = conv(v)
  return out


@dataclasses.dataclass(frozen=True)
class DecodeEmailOptions:
  """Options for `BaseTask.EmailDecodeSummary` (ref :57)."""
  job_name: str
  train_executions_per_eval: int
  global_step: int


@dataclasses.dataclass(frozen=True)
class ExecutorEma:
  """What an executor prepares for EMA (ref :69): the name → shadow-tensor map and the
  decay (a float or a schedule layer)."""
  ema: Optional[Dict[str, torch.Tensor]] = None
  ema_decay: Any = None


def _VariablesForEMA(params, model_var_list):
  """The variables EMA applies to (ref :77): trainable floating-point ones, plus the
  non-trainable moving statistics when `train.ema_decay_moving_vars`; de-duplicated, in
  `model_var_list` order."""
  out, seen = [], set()
  for v in model_var_list:
    name = getattr(v, 'var_name', '')
    moving = bool(params.train.ema_decay_moving_vars) and 'moving' in name
    if (v.requires_grad or moving) and v.is_floating_point() and id(v) not in seen:
      seen.add(id(v))
      out.append(v)
  return out


class BaseTask(base_layer.BaseLayer):
  """A single task: one input generator, learners, metrics."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input', None, 'Input generator Params.')
    p.Define('encoder', None, 'Encoder Params.')
    p.Define('online_encoder', None, 'Online Encoder Params.')
    p.Define('decoder', None, 'Decoder Params.')
    p.Define('task_global_step', False,
             'Use a task-specific global step (multi-task).')
    p.Define('defer_global_step_update', False, 'Kept for parity.')
    p.Define('train', hyperparams.Params(), 'Params to control training.')
    p.Define('ml_perf', hyperparams.Params(), 'MlPerf configuration.')
    tp = p.train
    tp.Define('start_up_delay_steps', 200, 'i-th replica starts at step '
              'i*(i+1)*start_up_delay_steps/2 (async only).')
    tp.Define('max_steps', 4 * 10**6, 'Maximum number of training steps.')
    tp.Define('tpu_steps_per_loop', 1000,
              'Steps per device loop (CUDA-graph replay length).')
    tp.Define('use_cuda_graph', 'auto',
              "Capture FProp+BProp+sync+optimizer into one CUDA graph and replay it: 'auto' "
              "(when on a GPU and the optimizer is capturable), 'on' (required) or 'off'.")
    tp.Define('device_prefetch_depth', 2,
              'Input batches staged on the device ahead of the step (pinned H2D copies on '
              'a side stream).')
    tp.Define('async_data_parallel', False,
              'With cluster mode `async` and several replicas: every replica applies its own '
              'gradients and parameters are reconciled by delayed non-blocking averaging '
              '(parallel/async_dp.py) — the parameter-server-free form of asynchronous training. '
              '`trainer --mode=async` turns it on.')
    tp.Define('async_sync_every_n_steps', 1,
              'Asynchronous mode: steps between parameter reconciliations (staleness bound).')
    tp.Define('tpu_device_order_mode', None, 'Kept for parity.')
    tp.Define('tpu_computation_shape', None, 'Kept for parity.')
    tp.Define('vn_start_step', 200000000, 'Step at which VN starts.')
    tp.Define('vn_std', 0.0, 'Std of the variational noise.')
    tp.Define('early_stop', early_stop.EarlyStop.Params(),
              'Early stopping based on dev-set performance.')
    tp.Define('ema_decay', 0.0, 'EMA decay; 0 disables.')
    tp.Define('ema_decay_moving_vars', None, 'Apply EMA to moving stats too.')
    tp.Define('ema_schedule', None, 'Schedule layer params for the EMA decay.')
    tp.Define('init_from_checkpoint_rules', {},
              '{ckpt: ([(regex, fmt)], [ignore_regex])} warm-start rules.')
    tp.Define('init_from_checkpoint_override', None, 'Overrides the ckpt path.')
    tp.Define('pruning_hparams_dict', None, 'Pruning hyper-parameters.')
    tp.Define('enqueue_max_steps', -1, 'Kept for parity.')
    tp.Define('save_interval_seconds', 60 * 10, 'Checkpoint interval (s).')
    tp.Define('save_interval_steps', None, 'Checkpoint interval (steps).')
    tp.Define('save_max_to_keep', 100, 'Max recent checkpoints kept.')
    tp.Define('save_keep_checkpoint_every_n_hours', 0.5, 'Keep-forever rate.')
    tp.Define('async_checkpointing', True, 'Snapshot on device, write async.')
    tp.Define('checkpoint_finite_check', False, 'Refuse to save NaN/Inf.')
    tp.Define('keep_per_example_loss', False, 'Kept for parity.')
    tp.Define('summary_interval_steps', 100, 'Summary interval.')
    tp.Define('learner', None, 'Learner params or list; None ⇒ legacy knobs.')
    # Legacy learner knobs (copied into a Learner when `learner` is None).
    tp.Define('l2_regularizer_weight', None, 'L2 weight.')
    tp.Define('l1_regularizer_weight', None, 'L1 weight.')
    tp.Define('learning_rate', 0.0, 'Learning rate.')
    tp.Define('clip_gradient_norm_to_value', 0.0, 'Global-norm clip.')
    tp.Define('clip_gradient_single_norm_to_value', 0.0, 'Per-tensor clip.')
    tp.Define('grad_norm_to_clip_to_zero', 0.0, 'Zero grads above this norm.')
    tp.Define('grad_norm_tracker', None, 'GradNormTracker params.')
    tp.Define('optimizer', optimizer.Adam.Params(), 'Optimizer params.')
    tp.Define('lr_schedule', schedule.ContinuousSchedule.Params(), 'LR schedule.')
    tp.Define('bprop_variable_filter', None, 'Include regex.')
    tp.Define('bprop_variable_exclusion', None, 'Exclude regex.')
    tp.Define('grad_aggregation_method', None, 'Kept for parity.')
    tp.Define('gate_gradients', False, 'Kept for parity.')
    tp.Define('colocate_gradients_with_ops', True, 'Kept for parity.')
    tp.Define('scale_gradients', True, 'Apply gradient scaling.')
    tp.Define('learner_use_variable_scope', True, 'Kept for parity.')
    tp.Define('sum_loss_across_tokens_in_batch', False, 'Kept for parity.')
    p.Define('eval', hyperparams.Params(), 'Params to control evaluation.')
    ep = p.eval
    ep.Define('samples_per_summary', 1000, 'Samples per eval; 0 = one epoch.')
    ep.Define('decoder_samples_per_summary', None, 'Samples per decode run.')
    ep.Define('load_checkpoint_from', '', 'Evaluate this ckpt / dir instead.')
    ep.Define('start_eval_after', 0, 'Start evaluation after this step.')
    ep.Define('start_decoder_after', 0, 'Start decoding after this step.')
    ep.Define('eval_all_checkpoints', False, 'Eval every checkpoint.')
    ep.Define('decode_all_checkpoints', False, 'Decode every checkpoint.')
    mp = p.ml_perf
    mp.Define('benchmark_name', None, 'MLPerf benchmark name.')
    mp.Define('steps_per_epoch', None, 'Steps per epoch.')
    mp.Define('decoder_metric_name', None, 'Metric name.')
    mp.Define('decoder_metric_success_threshold', None, 'Success threshold.')
    mp.Define('max_steps_to_train', 1000000000, 'Max steps.')
    return p

  def __init__(self, params):
    assert issubclass(params.cls, BaseTask)
    tp = params.train
    if tp and tp.learner is not None:
      if isinstance(tp.learner, (list, tuple)):
        names = [l.name for l in tp.learner]
        assert len(set(names)) == len(names), 'learner names must be unique'
    params = params.Copy()
    self._UpdateVnConfigOn(params)
    super().__init__(params)
    p = self.params
    self._post_training_loop_op = None
    self._per_input_gradient_mask = None
    self._encoder = None
    self._online_encoder = None
    self._decoder = None
    self._loss = None
    self._num_predictions = None
    self._train_op = None
    self._post_train_ops = []
    self._eval_metrics: Dict[str, Tuple[Any, Any]] = {}
    self._per_example = {}
    self._global_step = 0
    self._ema_applied_vars = None
    self._ema_map = None
    self._metrics = None
    self._last_var_grads = None

    # Input generator is created under the train/eval cluster of the caller.
    if p.input:
      if not p.input.name:
        p2 = p.input.Copy()
        p2.name = 'input'
      else:
        p2 = p.input
      self.CreateChild('input', input_policy.Apply(p2))

    tp = p.train
    if tp:
      if tp.learner is None:
        learner_params = [learner_lib.ExtractLearnerFromLegacyParams(tp)]
      elif isinstance(tp.learner, (list, tuple)):
        learner_params = list(tp.learner)
      else:
        learner_params = [tp.learner]
      self.CreateChildren('learners', learner_params)
      if tp.ema_schedule is not None:
        self.CreateChild('ema_schedule', tp.ema_schedule)
      if tp.early_stop and tp.early_stop.window:
        self._early_stop = early_stop.EarlyStop(tp.early_stop)
      else:
        self._early_stop = None

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    """Propagates the vocabulary size (and word-piece model) to the decoder (ref :338)."""
    dp = p.decoder
    p.decoder = dp.cls.UpdateTargetVocabSize(dp, vocab_size, wpm_model)
    return p

  @staticmethod
  def _UpdateVnConfigOn(p):
    """`train.vn_std` / `train.vn_start_step` drive the layers' variational noise (ref
    :1084): noise is on only when training, `vn_std > 0` and `p.vn` asks for global or
    per-step noise; the layer-level scale / start_step must be left unset."""
    from lingvo_b200.core import cluster_factory  # pylint: disable=g-import-not-at-top
    tp = p.train
    if not tp:
      return
    enabled = (tp.vn_std > 0) and p.vn and (p.vn.global_vn or p.vn.per_step_vn)
    if cluster_factory.Current().do_eval or not enabled:
      p.vn = py_utils.VariationalNoiseParams(None, False, False)
      return
    if p.vn.scale is not None:
      raise ValueError('A value should not be specified for p.vn.scale. It will be '
                       'overwritten by p.train.vn_std.')
    if p.vn.start_step:
      raise ValueError('A value should not be specified for p.vn.start_step. It will be '
                       'overwritten by p.train.vn_start_step.')
    p.vn = p.vn.Copy().Set(scale=tp.vn_std, start_step=tp.vn_start_step)

  def _UpdateVnConfig(self):
    """Applied to the params copy before construction; see `_UpdateVnConfigOn`."""

  def _SetLearnerFromLegacyParams(self, tp):
    if tp.learner is None:
      tp.learner = learner_lib.ExtractLearnerFromLegacyParams(tp)

  def _ComputeGradientMask(self, bprop_variable_filters):
    """mask[var][i] = 1 iff the variable's name matches filter i (ref :837): with
    cross-batch input mixing, the batch's source one-hot dotted with this row decides
    whether the variable is updated by that batch."""
    import re  # pylint: disable=g-import-not-at-top
    n = len(bprop_variable_filters)
    self._per_input_gradient_mask = NestedMap()
    for v in self.vars.Flatten():
      row = torch.zeros(n, dtype=torch.float32)
      for i, rx in enumerate(bprop_variable_filters):
        if re.search(rx, v.var_name):
          row[i] += 1.0
      self._per_input_gradient_mask[v.var_name] = row
    return self._per_input_gradient_mask

  def _GetMaskUpdateOp(self):
    """The pruning mask update to run after the optimizer step (ref :1105), or None."""
    tp = self.params.train
    if not tp.pruning_hparams_dict:
      return None
    assert isinstance(tp.pruning_hparams_dict, dict)
    from lingvo_b200.core import pruning_utils  # pylint: disable=g-import-not-at-top
    getter = getattr(pruning_utils.PruningOp, 'GetPruningUpdate', None)
    return getter() if getter is not None else None

  def CreateExponentialMovingAverage(self, ema=None):
    """Allocates the EMA shadows (initialised to the current values) ahead of the first
    `ApplyExponentialMovingAverage` (ref :859) — checkpoints restored before step 1 then
    find their `<var>/ExponentialMovingAverage` targets."""
    tp = self.params.train
    if ema is None and not (tp.ema_decay and tp.ema_decay > 0):
      return
    if self._ema_map is None:
      self._ema_map = {}
    for layer, key, var in self._EmaVars():
      if var.var_name not in self._ema_map:
        shadow = var.detach().clone()
        self._ema_map[var.var_name] = shadow
        layer.SetEmaShadow(key, shadow)

  def PostTrainingLoop(self, outfeed=None):
    """Runs every learner's post-loop hook (ref :708) and keeps the results."""
    del outfeed
    with py_utils.GlobalStepContext(self._global_step):
      self._post_training_loop_op = [
          l.ApplyPostTrainingLoop() if hasattr(l, 'ApplyPostTrainingLoop') else None
          for l in self.learners]

  @property
  def post_training_loop_op(self):
    assert self._post_training_loop_op is not None, (
        'No post_training_loop_op op is defined. Call PostTrainingLoop first.')
    return self._post_training_loop_op

  def InferenceEager(self):
    """{signature: callable} for eager serving; the default reuses `Inference()`."""
    return self.Inference()

  def EmailDecodeSummary(self, summaries, emails, options):
    raise NotImplementedError('Abstract method')

  def Export(self, train_dir):
    """Hook an eval job calls before evaluation to write extra files (ref :1126)."""
    del train_dir

  # ------------------------------------------------------------- properties --
  @property
  def input_generator(self):
    return self.input

  @property
  def global_step(self) -> int:
    return self._global_step

  @global_step.setter
  def global_step(self, v):
    self._global_step = int(v)

  @property
  def loss(self):
    assert self._loss is not None, 'No loss is defined. Call FProp first.'
    return self._loss

  @property
  def eval_metrics(self):
    return self._eval_metrics

  @property
  def per_example_tensors(self):
    return self._per_example

  @property
  def train_op(self):
    return self._train_op

  @property
  def learners(self):
    return self.children['learners']

  @property
  def has_early_stop(self):
    return getattr(self, '_early_stop', None) is not None

  # ------------------------------------------------------------------ fprop --
  def ComputePredictions(self, theta, input_batch):
    raise NotImplementedError('Abstract method')

  def ComputeLoss(self, theta, predictions, input_batch):
    raise NotImplementedError('Abstract method')

  def FilterPerExampleTensors(self, per_example):
    return per_example

  def ProcessFPropResults(self, sess, global_step, metrics, per_example):
    """Host callback after each train step (reference runners.py:330)."""
    return None

  def FPropTower(self, theta, input_batch):
    """(metrics {name: (value, weight)}, per_example {name: tensor})."""
    predictions = self.ComputePredictions(theta, input_batch)
    return self.ComputeLoss(theta, predictions, input_batch)

  def GetInputBatch(self):
    """Next list of per-split batches, on this task's device."""
    num = self.cluster.num_splits_per_client
    return self.input.SplitInputBatch(num)

  def _MoveBatch(self, batch: NestedMap, device) -> NestedMap:
    def mv(x):
      if isinstance(x, torch.Tensor) and x.device != device:
        return x.to(device, non_blocking=True)
      if isinstance(x, np.ndarray) and x.dtype.kind not in 'OUS':
        return torch.from_numpy(x).to(device, non_blocking=True)
      return x
    return batch.Transform(mv)

  def Device(self) -> torch.device:
    for v in self.vars.Flatten():
      return v.device
    return py_utils.CurrentDevice()

  def EnableMixedPrecision(self) -> int:
    """bf16 compute copies of fp32 master weights (see py_utils)."""
    self._mixed_precision_attached = True
    return py_utils.AttachComputeCopies(self)

  def FPropDefaultTheta(self, input_batch=None):
    if input_batch is None:
      input_batch = self.GetInputBatch()
    if not getattr(self, '_mixed_precision_attached', False):
      # First step on a GPU: switch bf16-fprop layers to persistent compute
      # copies (theta would otherwise re-cast 1.4 B weights every step).
      self._mixed_precision_attached = True
      if self.Device().type == 'cuda' and not self.do_eval:
        py_utils.AttachComputeCopies(self)
    return self.FProp(self.theta, input_batch)

  def FProp(self, theta, input_batch):
    """Forward over all towers of this process; returns (metrics, per_example)."""
    p = self.params
    with py_utils.GlobalStepContext(self._global_step):
      py_utils.ResetStepSeed(0)
      batches = input_batch if isinstance(input_batch, list) else [input_batch]
      dev = self.Device()
      all_metrics, all_per_example = [], []
      for split_id, batch in enumerate(batches):
        batch = self._MoveBatch(batch, dev)
        with cluster_factory.SetModelSplit(split_id):
          metrics, per_example = self.FPropTower(theta, batch)
        if not isinstance(metrics, dict):
          raise ValueError('FPropTower must return a dict of metrics')
        for name, item in metrics.items():
          if not (isinstance(item, (tuple, list)) and len(item) == 2):
            raise ValueError('metric %s must be a (value, weight) pair' % name)
        all_metrics.append(metrics)
        all_per_example.append(per_example or {})
      return self._FPropResult(all_metrics, all_per_example, batches)

  def _FPropResult(self, all_metrics, all_per_example, batches):
    if len(all_metrics) == 1:
      metrics = dict(all_metrics[0])
      per_example = all_per_example[0]
    else:
      metrics = py_utils.WeightedAvgOfMetrics(all_metrics)
      per_example = py_utils.ConcatPerExampleTensors(all_per_example)
    if 'num_samples_in_batch' not in metrics:
      n = 0
      for b in batches:
        for t in b.Flatten():
          if isinstance(t, torch.Tensor) and t.dim() > 0:
            n += t.shape[0]
            break
      metrics['num_samples_in_batch'] = (torch.tensor(float(n)),
                                         torch.tensor(1.0))
    loss_name = None
    for lrn in self.learners:
      names = lrn.params.loss_name or lrn.params.name
      names = names if isinstance(names, (list, tuple)) else [names]
      for n in names:
        if n not in metrics:
          raise ValueError('Loss %s of learner %s not in metrics %s' %
                           (n, lrn.params.name, sorted(metrics)))
      loss_name = loss_name or names[0]
    self._loss = metrics[loss_name][0] if loss_name else None
    if self._loss is not None:
      self._loss = py_utils.CheckNumerics(self._loss) if False else self._loss
    self._num_predictions = metrics[loss_name][1] if loss_name else None
    self._metrics = metrics
    self._eval_metrics = dict(metrics)
    self._per_example = self.FilterPerExampleTensors(per_example)
    return metrics, per_example

  # ------------------------------------------------------------------ bprop --
  def AdjustGradients(self, vars_gradients):
    return vars_gradients

  def BProp(self):
    """Learners → post-step hooks → EMA → global_step += 1 (:718-835)."""
    assert self._metrics is not None, 'Call FProp before BProp'
    with py_utils.GlobalStepContext(self._global_step):
      vmap = self.vars
      n = len(self.learners)
      # Only hand the learner an adjuster when a subclass really overrides it: the fused
      # optimizers fold the global-norm reduction into their statistics pass, which they
      # can only do when nothing rewrites the gradients in between.
      adjuster = self.AdjustGradients
      if getattr(type(self).AdjustGradients, '__func__', type(self).AdjustGradients) is \
          BaseTask.AdjustGradients:
        adjuster = None
      for i, lrn in enumerate(self.learners):
        losses, lm = lrn.Apply(self._metrics, vmap, gradient_adjuster=adjuster,
                               retain_graph=i < n - 1)
        self._last_var_grads = lrn.GetVarGrads()
        self._eval_metrics.update(lm)
      self.PostTrainingStepUpdate()
      self.ApplyExponentialMovingAverage()
      self.PostEmaUpdate()
    self._global_step += 1
    py_utils.SetGlobalStep(self._global_step)
    self._metrics = None

  def TrainStep(self, input_batch=None):
    """One full training step; returns (eval_metrics, per_example)."""
    self.FPropDefaultTheta(input_batch)
    self.BProp()
    return self._eval_metrics, self._per_example

  def EvalStep(self, input_batch=None):
    with torch.no_grad():
      return self.FPropDefaultTheta(input_batch)

  # -------------------------------------------------------------------- EMA --
  def _EmaVars(self):
    tp = self.params.train
    out = []
    for name, (layer, key, var) in _VarByName(self).items():
      if var.requires_grad or (tp.ema_decay_moving_vars and
                               'moving' in name):
        if var.is_floating_point():
          out.append((layer, key, var))
    return out

  @property
  def ema(self):
    return self._ema_map

  def ApplyExponentialMovingAverage(self):
    tp = self.params.train
    if not tp.ema_decay or tp.ema_decay <= 0:
      return
    decay = float(tp.ema_decay)
    if 'ema_schedule' in self.children:
      decay = float(self.ema_schedule.Value(self._global_step))
    if self._ema_map is None:
      self._ema_map = {}
      for layer, key, var in self._EmaVars():
        shadow = var.detach().clone()
        self._ema_map[var.var_name] = shadow
        layer.SetEmaShadow(key, shadow)
    vars_, shadows = [], []
    for layer, key, var in self._EmaVars():
      vars_.append(var.detach())
      shadows.append(self._ema_map[var.var_name])
    with torch.no_grad():
      # shadow -= (1 - decay) * (shadow - var)
      torch._foreach_lerp_(shadows, vars_, 1.0 - decay)

  def EmaShadowTensors(self) -> Dict[str, torch.Tensor]:
    """ckpt key → shadow (`<var>/ExponentialMovingAverage`)."""
    if not self._ema_map:
      return {}
    return {k[:-len('/var')] + '/var/ExponentialMovingAverage': v
            for k, v in self._ema_map.items()}

  def LoadEmaShadowTensors(self, tensors: Dict[str, torch.Tensor]) -> List[str]:
    used = []
    by_name = _VarByName(self)
    for key, t in tensors.items():
      if not key.endswith('/var/ExponentialMovingAverage'):
        continue
      vname = key[:-len('/ExponentialMovingAverage')]
      if vname in by_name:
        layer, k, var = by_name[vname]
        if self._ema_map is None:
          self._ema_map = {}
        shadow = t.to(var.device, var.dtype).clone()
        self._ema_map[vname] = shadow
        layer.SetEmaShadow(k, shadow)
        used.append(key)
    return used

  # ----------------------------------------------------------------- decode --
  def Decode(self, input_batch):
    """Decodes `input_batch` → dict of tensors."""
    raise NotImplementedError('Abstract method')

  def DecodeWithTheta(self, theta, input_batch):
    return self.Decode(input_batch)

  def Inference(self):
    """{subgraph_name: (fetches fn / feeds spec)} for serving export."""
    raise NotImplementedError('Abstract method')

  def CreateDecoderMetrics(self):
    raise NotImplementedError('Abstract method')

  def PostProcessDecodeOut(self, decode_out_dict, decode_metrics_dict):
    raise NotImplementedError('Abstract method')

  def DecodeFinalize(self, decode_finalize_args):
    return None

  def Eval(self, input_batch):
    return self.FPropDefaultTheta(input_batch)

  def ComputeEvalMetrics(self, *args, **kwargs):
    return {}

  def AddEvalMetric(self, name, value, weight, raise_if_already_added=True):
    if name in self._eval_metrics and raise_if_already_added:
      raise ValueError('Metric %s has already been defined.' % name)
    self._eval_metrics[name] = (value, weight)

  def AddPerExampleTensor(self, name, value):
    if name in self._per_example:
      raise ValueError('Metric %s has already been defined.' % name)
    self._per_example[name] = value


class BaseModel(base_layer.BaseLayer):
  """The abstract model: a collection of tasks."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('model', None, 'Which model this params is created for.')
    p.Define('cluster', cluster_factory.Cluster.Params(), 'Cluster params.')
    p.Define('input', None, 'Input generator Params (single-task).')
    p.Define('train', hyperparams.Params(), 'Params shared between tasks.')
    tp = p.train
    for name, default, doc in [
        ('start_up_delay_steps', 200, 'See BaseTask.'),
        ('max_steps', 4 * 10**6, 'Maximum number of training steps.'),
        ('tpu_steps_per_loop', 1000, 'Steps per device loop.'),
        ('tpu_device_order_mode', None, 'Kept for parity.'),
        ('tpu_computation_shape', None, 'Kept for parity.'),
        ('ema_decay', 0.0, 'EMA decay.'),
        ('ema_decay_moving_vars', None, 'EMA of moving vars.'),
        ('init_from_checkpoint_rules', {}, 'Warm-start rules.'),
        ('init_from_checkpoint_override', None, 'Warm-start override.'),
        ('early_stop', None, 'Early stop params.'),
        ('enqueue_max_steps', -1, 'Kept for parity.'),
        ('save_interval_seconds', 60 * 10, 'Checkpoint interval (s).'),
        ('save_interval_steps', None, 'Checkpoint interval (steps).'),
        ('save_max_to_keep', 100, 'Max checkpoints kept.'),
        ('save_keep_checkpoint_every_n_hours', 0.5, 'Keep-forever rate.'),
        ('async_checkpointing', True, 'Async checkpoint writes.'),
        ('checkpoint_finite_check', False, 'Finite check before saving.'),
        ('summary_interval_steps', 100, 'Summary interval.'),
    ]:
      tp.Define(name, default, doc)
    p.Define('eval', hyperparams.Params(), 'Eval params.')
    p.eval.Define('samples_per_summary', 1000, 'Samples per summary.')
    p.eval.Define('decoder_samples_per_summary', None, 'Samples per decode.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._global_step = 0
    self.modules = []

  @property
  def global_step(self):
    return self._global_step

  @property
  def tasks(self) -> List[BaseTask]:
    raise NotImplementedError('Abstract method')

  @property
  def task_names(self):
    raise NotImplementedError('Abstract method')

  def GetTask(self, task_name=None) -> BaseTask:
    raise NotImplementedError('Abstract method')

  def SampleTask(self, global_step):
    raise NotImplementedError('Abstract method')

  @property
  def ema(self):
    return None

  def ConstructFPropBPropGraph(self):
    """One FProp+BProp of the (sampled) task (reference :1358)."""
    task = self.SampleTask(self._global_step)
    out = task.TrainStep()
    self._global_step += 1
    return out

  def ConstructFPropGraph(self):
    task = self.SampleTask(self._global_step)
    return task.EvalStep()

  def ConstructPostTrainingLoop(self, *args, **kwargs):
    return None

  @property
  def variables_for_ema(self):
    return _VariablesForEMA(self.params, self.vars.Flatten())

  @property
  def ema_decay(self):
    return self.params.train.ema_decay

  def MakeEMAVariablesDictTF2(self):
    """{checkpoint key: shadow tensor} over all tasks (ref :1253)."""
    res = {}
    for task in self.tasks:
      res.update(task.EmaShadowTensors())
    self._ema_variables_dict = res
    return res

  def ProcessFPropResults(self, sess, global_step, metrics, per_example):
    for task in self.tasks:
      task.ProcessFPropResults(sess, global_step, metrics, per_example)

  def Export(self, train_dir):
    for task in self.tasks:
      task.Export(train_dir)

  def ConstructDecodeGraph(self, task_name=None, input_batch=None):
    task = self.GetTask(task_name)
    if input_batch is None:
      input_batch = task.GetInputBatch()[0]
    input_batch = task._MoveBatch(input_batch, task.Device())  # pylint: disable=protected-access
    with torch.no_grad():
      return task.Decode(input_batch)


class SingleTaskBase(BaseModel):
  """Represents a single task model (reference :1363)."""

  @property
  def tasks(self):
    return [self._task]

  @property
  def task_names(self):
    return [None]

  def GetTask(self, task_name=None):
    assert not task_name, 'SingleTaskModel has a single, unnamed task'
    return self._task

  def SampleTask(self, global_step):
    return self._task

  @property
  def ema(self):
    return self._task.ema


class SingleTaskModel(SingleTaskBase):
  """Model that consists of a single task (reference :1379)."""

  # The task's variables live at the root scope (ckpt key `lenet5/conv0/...`).
  _child_variable_scope_override = {'_task': []}

  @classmethod
  def Params(cls, task_params=None):
    p = super().Params()
    p.Define('task', None,
             '`InstantiableParams` object for a `BaseTask` subclass.')
    if task_params is not None:
      cls.CopyTaskParams(task_params, p)
    return p

  @classmethod
  def CopyTaskParams(cls, task_params, p):
    """Mirrors the task's name / train / eval / input settings on the model params (ref
    :1392) so runners can read them without reaching into the task."""
    assert task_params is not None
    p.task = task_params
    p.Set(name=task_params.name)
    tp = p.train
    tt = task_params.train
    for k, _ in tp.IterParams():
      if k in tt:
        tp.Set(**{k: tt.Get(k)})
    p.eval.samples_per_summary = task_params.eval.samples_per_summary
    p.eval.decoder_samples_per_summary = (
        task_params.eval.decoder_samples_per_summary)
    p.input = task_params.input
    return p

  def __init__(self, params):
    p = params
    assert p.name == p.task.name, (p.name, p.task.name)
    super().__init__(params)
    p = self.params
    tp = p.task.Copy()
    if p.input is not None:
      tp.input = p.input
    self.CreateChild('_task', tp)

  def InstantiateVariables(self):
    # Root scope holds the task directly: <task name>/<layer>/<var>/var.
    if self._variables_instantiated:
      return
    with base_layer._ManagedScope():  # pylint: disable=protected-access
      self._variables_instantiated = True
      with py_utils.VariableScope([self._task.params.name]):
        self._task._InstantiateSelfAndChildren()  # pylint: disable=protected-access


class MultiTaskSubModel(SingleTaskBase):
  """'Model' consisting of a task from a multi-task model (:1442)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'multi_task_sub_model'
    p.Define('task_name', '', 'The name of the task to execute.')
    return p

  def __init__(self, params, shared_model=None):
    super().__init__(params)
    p = self.params
    self._model = shared_model
    self._task = self._model.children[p.task_name]

  def GetVariablesDict(self):
    """The whole shared model's variables: a sub-model checkpoints everything (ref
    :1468)."""
    return {v.var_name: v for v in self._model.vars.Flatten()}


class MultiTaskModel(BaseModel):
  """Model that consists of multiple tasks (reference :1480-1640)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('task_params', hyperparams.Params(),
             'Params object mapping task name to task Params.')
    p.Define('task_probs', hyperparams.Params(),
             'Params object mapping task name to sampling probability.')
    p.Define('task_schedule', None, 'Task schedule params.')
    p.Define('task_global_step', False, 'Per-task global steps.')
    p.Define('task_name_var_scope', True, 'Scope task vars by task name.')
    p.Define('share_model_object', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert len(p.task_params) > 0
    self._task_names = sorted(k for k, _ in p.task_params.IterParams())
    if p.task_schedule is None:
      sched = task_scheduler.ConstantScheduler.Params()
      sched.task_probs = sorted(
          (k, v) for k, v in p.task_probs.IterParams())
    else:
      sched = p.task_schedule
    sched.name = 'task_schedule'
    self.CreateChild('task_schedule', sched)
    for name in self._task_names:
      tp = p.task_params.Get(name).Copy()
      assert tp.name == name or not tp.name, (tp.name, name)
      tp.name = name
      if p.task_global_step:
        tp.task_global_step = True
      self.CreateChild(name, tp)

  @staticmethod
  def TaskNames(params):
    return sorted(name for name, _ in params.task_params.IterParams())

  def _ChildScope(self, child_key, child):
    if not self.params.task_name_var_scope and child_key in self._task_names:
      return []
    return super()._ChildScope(child_key, child)

  @property
  def task_names(self):
    return self._task_names

  @property
  def tasks(self):
    return [self.children[n] for n in self._task_names]

  def GetTask(self, task_name=None):
    assert task_name, 'It is required to specify task_name'
    return self.children[task_name]

  def SampleTask(self, global_step):
    name = self.task_schedule.Sample(global_step)
    return self.children[name]

  def ConstructFPropBPropGraph(self):
    task = self.SampleTask(self._global_step)
    self.last_task_name = task.params.name
    out = task.TrainStep()
    self._global_step += 1
    return out
