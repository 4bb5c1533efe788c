"""Checkpoint save/restore orchestration.

Reference `lingvo/core/checkpointer.py`: `Checkpointer.{Restore, Save,
MaybeSave, RestoreFromPath, RestoreIfNeeded, ShouldSave}` (:138-398), key
scheme = variable names (`lenet5/conv0/w/var`), `global_step`, optimizer
slots `<var>/Adam`…, EMA shadows `<var>/ExponentialMovingAverage`
(:425-462); `init_from_checkpoint_rules` warm start (:214-266, 354-388,
`py_utils.py:2745-2907`); save policy (:281-336).
"""

from __future__ import annotations

import logging
import os
import re
import time
from typing import Dict, List, Optional, Tuple

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import py_utils
from lingvo_b200.core import saver as saver_lib
from lingvo_b200.core import train_engine
from lingvo_b200.utils import tensor_bundle


def _RankWorld() -> Tuple[int, int]:
  import torch.distributed as dist  # pylint: disable=g-import-not-at-top
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def _EpSlices(model) -> Dict[str, Tuple[int, int, int, int]]:
  """`<var base name>` → (lo, hi, num_experts, ep_size) for expert-parallel variables:
  this rank holds rows [lo, hi) of dim 0 of the logical `[E, …]` tensor."""
  out = {}
  for v in model.vars.Flatten():
    shard = getattr(v, 'ep_shard', None)
    if shard is None or not getattr(v, 'expert_parallel', False):
      continue
    ep_rank, ep_size, e = shard
    el = e // ep_size
    base = v.var_name[:-len('/var')] if v.var_name.endswith('/var') else v.var_name
    out[base] = (ep_rank * el, (ep_rank + 1) * el, e, ep_size)
  return out


def _EpBaseOf(key: str, ep: Dict[str, Tuple[int, int, int, int]]) -> Optional[str]:
  """The expert-parallel variable a checkpoint key (var, slot or EMA shadow) belongs to."""
  probe = key
  while probe:
    if probe in ep:
      return probe
    probe, _, _ = probe.rpartition('/')
  return None


def _TpShards(model) -> Dict[str, Tuple[int, int, int, int]]:
  """`<var base name>` → (tp_rank, tp_size, dim, logical_size) of tensor-parallel shards."""
  out = {}
  for v in model.vars.Flatten():
    shard = getattr(v, 'tp_shard', None)
    if shard is None:
      continue
    base = v.var_name[:-len('/var')] if v.var_name.endswith('/var') else v.var_name
    out[base] = tuple(shard)
  return out


def _TpKey(key: str, tp_rank: int, tp_size: int) -> str:
  """Name of a per-TP-rank tensor (optimizer slots / EMA of a sharded variable)."""
  return '%s/__tp%d_of_%d' % (key, tp_rank, tp_size)


def _TpTensors(named, model, rank) -> Dict[str, torch.Tensor]:
  """Tensor-parallel jobs. Variables are written once, *gathered to their logical shape*
  (so evalers, decoders and jobs with another TP degree read an ordinary checkpoint);
  optimizer slots and EMA shadows of sharded variables stay per-rank (`__tp<r>_of_<n>`
  keys: factored second-moment statistics of a shard are not a slice of the unsharded
  statistics). Written by the dp_rank-0 replica: rank 0 the replicated tensors and the
  gathered variables, every rank of TP group 0 its own slot shards."""
  from lingvo_b200.parallel import mesh as mesh_lib   # pylint: disable=g-import-not-at-top
  from lingvo_b200.parallel import tp_layers   # pylint: disable=g-import-not-at-top
  ctx = mesh_lib.TensorParallel()
  tp = _TpShards(model)
  out = {}
  var_keys = {v.var_name for v in model.vars.Flatten()}
  for key in sorted(named):                       # same order on all ranks: collectives inside
    t = named[key]
    base = _EpBaseOf(key, tp)
    if base is None:
      if rank == 0:
        out[key] = t
      continue
    _, tp_size, dim, _ = tp[base]
    if key in var_keys:
      full = tp_layers.GatherShards(t.detach(), ctx, dim)
      if rank == 0:
        out[key] = full
    elif ctx.dp_rank == 0:
      out[_TpKey(key, ctx.tp_rank, tp_size)] = t
  return out


def _ModelTensors(model, rank: int = 0, world: int = 1) -> Dict[str, torch.Tensor]:
  """The tensors *this rank* writes. Rank 0 owns everything replicated; every rank of the
  first EP group additionally owns its dim-0 slice of the expert-parallel variables and of
  their optimizer slots / EMA shadows (saved under `tensor_bundle.SliceKey` names)."""
  ep = _EpSlices(model) if world > 1 else {}
  named = {}
  for v in model.vars.Flatten():
    named[v.var_name] = v
  for task in model.tasks:
    for lrn in task.learners:
      named.update(lrn.optimizer.GetOptimizerSlots())
    named.update(task.EmaShadowTensors())
  if world <= 1:
    return named
  if _TpShards(model):
    return _TpTensors(named, model, rank)
  out = {}
  for key, t in named.items():
    base = _EpBaseOf(key, ep)
    if base is not None:
      lo, hi, e, ep_size = ep[base]
      if isinstance(t, torch.Tensor) and t.dim() >= 1 and t.shape[0] == hi - lo:
        if rank < ep_size:
          out[tensor_bundle.SliceKey(key, lo, hi, e)] = t
        continue
    if rank == 0:
      out[key] = t
  return out


class Checkpointer:
  """Checkpointing utility bound to one model and one train dir."""

  def __init__(self, train_dir: str, model, init_op=None, train_params=None,
               save_only=False, check_loading_status=True):
    self._train_dir = train_dir
    self._model = model
    self._save_only = save_only
    self._params = train_params or model.params.train
    tp = self._params
    self._save_path = os.path.join(train_dir, 'ckpt')
    self._next_checkpoint_seconds = 0
    self._save_interval_seconds = tp.save_interval_seconds
    self._save_interval_steps = tp.save_interval_steps
    self._prev_ckpt_step = None
    self._saved_first = False
    checks = []
    if getattr(tp, 'checkpoint_finite_check', False):
      checks.append((lambda name: True, [saver_lib.IsFinite()]))
    self._rank, self._world = _RankWorld()
    # Sharded bundle only when some variable differs per rank (expert parallelism);
    # otherwise rank 0 alone writes a single-shard bundle.
    self._sharded = self._world > 1 and bool(_EpSlices(model) or _TpShards(model))
    self._saver = saver_lib.Saver(
        train_dir, lambda: _ModelTensors(model, self._rank, self._world),
        sanity_checks=checks,
        keep_latest_n=tp.save_max_to_keep,
        keep_every_n_hours=tp.save_keep_checkpoint_every_n_hours,
        async_save=bool(tp.async_checkpointing),
        shard_id=self._rank if self._sharded else 0,
        num_shards=self._world if self._sharded else 1)
    self._engines = []
    self._init_rules_applied = False
    os.makedirs(train_dir, exist_ok=True)

  def AttachEngine(self, engine):
    """`TrainEngine`s whose `PreSave()` must run before tensors are snapshotted."""
    self._engines.append(engine)

  @property
  def checkpoint_dir(self):
    return self._train_dir

  @property
  def async_checkpointing(self):
    return bool(self._params.async_checkpointing)

  # ---------------------------------------------------------------- restore --
  def _GlobalStep(self) -> int:
    return self._model.tasks[0].global_step if len(self._model.tasks) == 1 \
        else self._model.global_step

  def _SetGlobalStep(self, step: int):
    self._model._global_step = int(step)  # pylint: disable=protected-access
    for t in self._model.tasks:
      t.global_step = int(step)
    py_utils.SetGlobalStep(int(step))

  def RestoreFromPath(self, sess=None, checkpoint_path: str = None,
                      strict: bool = True) -> int:
    """Loads every model variable (+ slots, EMA) present in the bundle."""
    assert not self._save_only
    reader = tensor_bundle.BundleReader(checkpoint_path)
    keys = set(reader.LogicalKeys())
    ep = _EpSlices(self._model) if self._world > 1 else {}
    tp = _TpShards(self._model)
    missing = []

    def read(key):
      base = _EpBaseOf(key, ep)
      if base is not None:
        lo, hi, e, _ = ep[base]
        shape = reader.LogicalShape(key)
        if shape and shape[0] == e:
          return saver_lib.FromNumpy(reader.ReadRange(key, lo, hi))
      tbase = _EpBaseOf(key, tp)
      if tbase is not None:
        tp_rank, tp_size, dim, logical = tp[tbase]
        shape = reader.LogicalShape(key)
        if shape and len(shape) > dim and shape[dim] == logical:
          # a variable stored with its logical shape: keep this rank's slice
          n = logical // tp_size
          if dim == 0:
            return saver_lib.FromNumpy(reader.ReadRange(key, tp_rank * n, (tp_rank + 1) * n))
          full = saver_lib.FromNumpy(reader.ReadRange(key))
          return full.narrow(dim, tp_rank * n, n).contiguous()
      return saver_lib.FromNumpy(reader.ReadRange(key))

    with torch.no_grad():
      for v in self._model.vars.Flatten():
        k = v.var_name
        if k not in keys:
          missing.append(k)
          continue
        t = read(k)
        if tuple(t.shape) != tuple(v.shape):
          raise ValueError('Shape mismatch for %s: ckpt %s vs model %s' %
                           (k, tuple(t.shape), tuple(v.shape)))
        v.data.copy_(t.to(v.device, v.dtype))
    if missing and strict:
      raise KeyError('Variables missing from checkpoint %s: %s' %
                     (checkpoint_path, missing[:10]))
    rest = {k: read(k) for k in keys
            if not k.endswith('/var') and k != 'global_step' and '/__tp' not in k}
    if tp:
      # per-TP-rank slot / EMA tensors of this rank (absent when the checkpoint was written
      # with another TP degree: those slots then start from their initial values)
      tp_rank, tp_size = next(iter(tp.values()))[:2]
      suffix = '/__tp%d_of_%d' % (tp_rank, tp_size)
      for k in keys:
        if k.endswith(suffix):
          rest[k[:-len(suffix)]] = saver_lib.FromNumpy(reader.ReadRange(k))
    for task in self._model.tasks:
      for lrn in task.learners:
        # Slots are created lazily; load restores them on the var's device.
        dev = task.Device()
        lrn.optimizer.LoadOptimizerSlots(
            {k: t.to(dev) if t.dim() else t for k, t in rest.items()})
      task.LoadEmaShadowTensors(rest)
      # compute copies, carried optimizer scratch, sharded-optimizer engines
      train_engine.PostRestore(task)
    step = 0
    if 'global_step' in keys:
      step = int(reader.Read('global_step').reshape(-1)[0])
    else:
      m = re.search(r'ckpt-(\d+)$', checkpoint_path)
      if m:
        step = int(m.group(1))
    self._SetGlobalStep(step)
    reader.Close()
    logging.info('Restored %s at step %d', checkpoint_path, step)
    return step

  def _ApplyInitFromCheckpointRules(self):
    """Warm start: {ckpt: ([(regex, fmt)], [ignore_regex])} (:354-388)."""
    tp = self._params
    rules = dict(getattr(tp, 'init_from_checkpoint_rules', {}) or {})
    for task in self._model.tasks:
      rules.update(task.params.train.init_from_checkpoint_rules or {})
    override = getattr(tp, 'init_from_checkpoint_override', None)
    if not rules:
      return
    loaded = set()
    for ckpt_path, (var_rules, ignore_rules) in rules.items():
      path = override or ckpt_path
      if os.path.isdir(path):
        path = saver_lib.LatestCheckpoint(path)
      reader = tensor_bundle.BundleReader(path)
      keys = set(reader.Keys())
      with torch.no_grad():
        for v in self._model.vars.Flatten():
          name = v.var_name[:-len('/var')] if v.var_name.endswith('/var') \
              else v.var_name
          if name in loaded:
            continue
          if any(re.match(r, name) for r in ignore_rules):
            continue
          for regex, fmt in var_rules:
            m = re.match(regex, name)
            if not m:
              continue
            src = fmt % m.groups() if m.groups() else fmt
            cands = [src, src + '/var']
            hit = [c for c in cands if c in keys]
            if not hit:
              raise KeyError('%s → %s not found in %s' % (name, src, path))
            t = saver_lib.FromNumpy(reader.Read(hit[0]))
            v.data.copy_(t.to(v.device, v.dtype))
            loaded.add(name)
            break
      reader.Close()
    for task in self._model.tasks:
      train_engine.PostRestore(task)
    logging.info('init_from_checkpoint_rules loaded %d variables', len(loaded))

  def Restore(self, sess=None, force_reinitialize=False) -> Optional[str]:
    """Latest checkpoint in train_dir, else (re)initialise + warm-start."""
    path = None if force_reinitialize else saver_lib.LatestCheckpoint(
        self._train_dir)
    if path:
      self.RestoreFromPath(checkpoint_path=path)
      return path
    if not self._init_rules_applied:
      self._ApplyInitFromCheckpointRules()
      self._init_rules_applied = True
    return None

  def RestoreIfNeeded(self, sess=None):
    return self.Restore(sess)

  def RestoreGlobalStepIfNeeded(self, sess=None):
    path = saver_lib.LatestCheckpoint(self._train_dir)
    if path:
      m = re.search(r'ckpt-(\d+)$', path)
      if m:
        self._SetGlobalStep(int(m.group(1)))

  # ------------------------------------------------------------------- save --
  def ShouldSave(self, gsteps: int) -> bool:
    if self._save_only and False:
      return False
    if not self._saved_first:
      return True
    if self._save_interval_steps:
      if self._prev_ckpt_step is None:
        return True
      return gsteps - self._prev_ckpt_step >= self._save_interval_steps
    return time.time() >= self._next_checkpoint_seconds

  def Save(self, sess=None, gsteps: Optional[int] = None, sync=True) -> str:
    gsteps = self._GlobalStep() if gsteps is None else gsteps
    if self._saved_first and self._prev_ckpt_step == gsteps:
      if sync:
        self._saver.Wait()
      return '%s-%08d' % (self._save_path, int(gsteps))      # this step is already saved
    for eng in self._engines:
      eng.PreSave()                # collective: every rank calls Save at the same step
    path = '%s-%08d' % (self._save_path, int(gsteps))
    if self._rank == 0 or self._sharded:
      tensors = _ModelTensors(self._model, self._rank, self._world)
      if self._rank == 0:
        tensors['global_step'] = torch.tensor(int(gsteps), dtype=torch.int64)
      path = self._saver.Save(gsteps, tensors)   # device → host snapshot happens here
      if sync:
        self._saver.Wait()
    for eng in self._engines:
      eng.PostSave()
    self._saved_first = True
    self._prev_ckpt_step = gsteps
    self._next_checkpoint_seconds = time.time() + (
        self._save_interval_seconds or 0)
    return path

  def _AgreedShouldSave(self, gsteps: int) -> bool:
    """Same decision on every rank. Step-based policies are deterministic; the wall-clock
    policy is decided by rank 0 and broadcast, at most once every 20 steps."""
    if self._world <= 1 or self._save_interval_steps or not self._saved_first:
      return self.ShouldSave(gsteps)
    if gsteps % 20:
      return False
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    flag = torch.tensor([int(self.ShouldSave(gsteps))], device=dev)
    dist.broadcast(flag, src=0)
    return bool(flag.item())

  def MaybeSave(self, sess=None, gsteps: Optional[int] = None):
    gsteps = self._GlobalStep() if gsteps is None else gsteps
    if self._AgreedShouldSave(gsteps):
      return self.Save(sess, gsteps, sync=not self.async_checkpointing)
    return None

  def Sync(self):
    self._saver.Wait()


def GetSpecificCheckpoint(load_checkpoint_from: str) -> Optional[str]:
  """A ckpt prefix or a directory (→ latest) (reference :86-115)."""
  if not load_checkpoint_from:
    return None
  if os.path.isdir(load_checkpoint_from):
    return saver_lib.LatestCheckpoint(load_checkpoint_from)
  if os.path.exists(load_checkpoint_from + '.index'):
    return load_checkpoint_from
  raise ValueError('Invalid load_checkpoint_from: %s' % load_checkpoint_from)


def SortCheckpointPaths(ckpts):
  """Checkpoint prefixes ordered by the step number at the end of the path (ref :31)."""
  return sorted(ckpts, key=lambda x: int(x.split('-')[-1]))


class SaverWrapper:
  """The save / restore / sync trio over a flat `{name: tensor}` view (ref :36): the piece
  runners use when they checkpoint something that is not a `BaseModel` (EMA-substituted
  eval variables, auxiliary state). `variables_to_restore_dict` restricts / renames what is
  restored: {checkpoint name: tensor}."""

  def __init__(self, logdir, train_params, variables_to_restore_dict=None, async_save=False,
               variables_fn=None):
    assert variables_to_restore_dict is not None or variables_fn is not None
    self._logdir = logdir
    self._save_path = os.path.join(logdir, 'ckpt')
    self._restore_dict = variables_to_restore_dict
    vars_fn = variables_fn or (lambda: dict(variables_to_restore_dict))
    checks = []
    max_steps = getattr(train_params, 'max_steps', 0)
    per_loop = getattr(train_params, 'tpu_steps_per_loop', 0)
    if max_steps and per_loop:
      checks.append((r'^global_step$', [saver_lib.InRange(0, max_steps + per_loop)]))
    if getattr(train_params, 'checkpoint_finite_check', False):
      checks.append((lambda name: True, [saver_lib.IsFinite()]))
    self._saver = saver_lib.Saver(
        logdir, vars_fn, sanity_checks=checks,
        keep_latest_n=getattr(train_params, 'save_max_to_keep', None),
        keep_every_n_hours=getattr(train_params, 'save_keep_checkpoint_every_n_hours', None),
        async_save=async_save)

  def Save(self, sess, gsteps):
    del sess
    tensors = dict(self._saver._vars_fn())   # pylint: disable=protected-access
    tensors.setdefault('global_step', torch.tensor(int(gsteps), dtype=torch.int64))
    return self._saver.Save(int(gsteps), tensors)

  def Restore(self, sess, path):
    del sess
    return self._saver.Restore(path=path, strict=self._restore_dict is not None)[1]

  def Sync(self):
    self._saver.Sync()


# Execution is always eager here: the V1 flavour of the reference's eager checkpointer
# (`ckpt-%08d` bundles in the train dir) IS `Checkpointer`.
EagerCheckpointerV1 = Checkpointer


class EagerCheckpointerV2(Checkpointer):
  """Object-graph flavour (ref :638): checkpoints live in `<train_dir>/ckpt_V2/` so both
  generations can coexist in one log dir; optional asynchronous writes."""

  def __init__(self, train_dir, model, train_params=None, save_only=False,
               check_loading_status=True, experimental_enable_async_checkpoint=False):
    tp = (train_params or model.params.train)
    if experimental_enable_async_checkpoint and not tp.async_checkpointing:
      tp = tp.Copy().Set(async_checkpointing=True)
    super().__init__(os.path.join(train_dir, 'ckpt_V2'), model, train_params=tp,
                     save_only=save_only, check_loading_status=check_loading_status)
