"""Checkpoint writer with the reference file layout and policies.

Reference `lingvo/core/saver.py`: sanity checks `IsFinite/InRange` before
commit (:64-93,313-333), `ckpt-%08d` prefix (:196-201), sharded save+merge
(:168-194), keep policies, async save by on-device snapshot then background
write (:212-256,335-393). Also maintains the TF `checkpoint` state file so
`latest_checkpoint` keeps working for reference tooling.
"""

from __future__ import annotations

import glob
import os
import pickle
import re
import threading
import time
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from lingvo_b200.utils import tensor_bundle


class SanityCheck:

  def Check(self, *args):
    raise NotImplementedError()


class InRange(SanityCheck):
  """Every element is in [low, high]."""

  def __init__(self, low, high):
    self._low, self._high = low, high

  def Check(self, name, t: torch.Tensor) -> bool:
    return bool((t >= self._low).all() and (t <= self._high).all())

  def __str__(self):
    return 'InRange({}, {})'.format(self._low, self._high)


class IsFinite(SanityCheck):

  def Check(self, name, t: torch.Tensor) -> bool:
    if not t.is_floating_point():
      return True
    return bool(torch.isfinite(t).all())

  def __str__(self):
    return 'IsFinite'


class SanityCheckFailed(Exception):
  pass


def _ToNumpy(t: torch.Tensor):
  t = t.detach()
  if t.dtype == torch.bfloat16:
    return tensor_bundle.BFloat16Array(
        t.contiguous().view(torch.uint16).cpu().numpy())
  return t.cpu().numpy()


def FromNumpy(a) -> torch.Tensor:
  if isinstance(a, tensor_bundle.BFloat16Array):
    return torch.from_numpy(a.bits.copy()).view(torch.bfloat16)
  return torch.from_numpy(np.ascontiguousarray(a))


def CheckpointStatePath(train_dir: str) -> str:
  return os.path.join(train_dir, 'checkpoint')


def ReadCheckpointState(train_dir: str) -> Optional[Dict[str, List[str]]]:
  path = CheckpointStatePath(train_dir)
  if not os.path.exists(path):
    return None
  state = {'model_checkpoint_path': None, 'all_model_checkpoint_paths': []}
  with open(path) as f:
    for line in f:
      m = re.match(r'\s*(\w+)\s*:\s*"(.*)"\s*$', line)
      if not m:
        continue
      k, v = m.group(1), m.group(2)
      if k == 'model_checkpoint_path':
        state[k] = v
      elif k == 'all_model_checkpoint_paths':
        state[k].append(v)
  return state


def LatestCheckpoint(train_dir: str) -> Optional[str]:
  """Full prefix of the newest complete checkpoint in `train_dir`."""
  state = ReadCheckpointState(train_dir)
  if state and state['model_checkpoint_path']:
    p = state['model_checkpoint_path']
    if not os.path.isabs(p):
      p = os.path.join(train_dir, p)
    if os.path.exists(p + '.index'):
      return p
  cands = sorted(glob.glob(os.path.join(train_dir, 'ckpt-*.index')))
  return cands[-1][:-len('.index')] if cands else None


def AllCheckpoints(train_dir: str) -> List[str]:
  return [p[:-len('.index')] for p in
          sorted(glob.glob(os.path.join(train_dir, 'ckpt-*.index')))]


class Saver:
  """Saves `{name: tensor}` maps as `ckpt-%08d` bundles under a directory."""

  def __init__(self, logdir: str, variables_fn: Callable[[], Dict[str, torch.Tensor]],
               sanity_checks=None, keep_latest_n=None,
               keep_every_n_hours=None, async_save=False, prefix='ckpt',
               shard_id: int = 0, num_shards: int = 1, shard_timeout_s: float = 600.0):
    """`num_shards > 1`: one Saver per rank. Each rank writes the tensors it owns into
    `…data-0000r-of-0000N` plus a small entries side-car; shard 0 merges them into the
    index (reference `saver.py:168-194` sharded save + merge), keeps the `checkpoint`
    state file and runs the keep policy. No collective is involved, so async saves are
    safe on any thread."""
    self._shard, self._num_shards = int(shard_id), int(num_shards)
    self._shard_timeout_s = shard_timeout_s
    self._logdir = logdir
    self._vars_fn = variables_fn
    self._sanity_checks = sanity_checks or []
    self._keep_latest_n = keep_latest_n
    self._keep_every_n_hours = keep_every_n_hours
    self._async = async_save
    self._prefix = os.path.join(logdir, prefix)
    self._lock = threading.Lock()
    self._thread: Optional[threading.Thread] = None
    self._last_kept_time = time.time()
    self._error: Optional[BaseException] = None
    os.makedirs(logdir, exist_ok=True)

  def _DoSanityChecks(self, tensors):
    for pattern_or_fn, checks in self._sanity_checks:
      for name, t in tensors.items():
        hit = (re.search(pattern_or_fn, name) if isinstance(pattern_or_fn, str)
               else pattern_or_fn(name))
        if not hit:
          continue
        for c in checks:
          if not c.Check(name, t):
            raise SanityCheckFailed('Sanity check %s failed for %s' % (c, name))

  def _SidecarPath(self, prefix: str, shard: int, seq: int) -> str:
    # `seq` counts this Saver's saves (all ranks save in lockstep), so a fast rank that is
    # already writing the *next* save of the same prefix cannot be mistaken for this one.
    return '%s.entries-%05d-%d' % (prefix, shard, seq)

  def _Write(self, prefix: str, snapshot: Dict[str, object], global_step: int,
             seq: int = 0):
    if self._num_shards == 1:
      w = tensor_bundle.BundleWriter(prefix)
      for name in sorted(snapshot):
        w.Add(name, snapshot[name])
      w.Finish()
      self._UpdateState(prefix)
      self._GarbageCollect()
      return
    w = tensor_bundle.BundleWriter(prefix, self._shard, self._num_shards)
    for name in sorted(snapshot):
      w.Add(name, snapshot[name])
    entries = w.FinishShard()
    side = self._SidecarPath(prefix, self._shard, seq)
    with open(side + '.tmp', 'wb') as f:
      pickle.dump(entries, f, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(side + '.tmp', side)
    if self._shard != 0:
      return
    # Shard 0 commits the checkpoint once every shard's side-car has appeared.
    deadline = time.time() + self._shard_timeout_s
    all_entries = []
    for r in range(self._num_shards):
      path = self._SidecarPath(prefix, r, seq)
      while not os.path.exists(path):
        if time.time() > deadline:
          raise TimeoutError('checkpoint shard %d of %s never arrived' % (r, prefix))
        time.sleep(0.01)
      with open(path, 'rb') as f:
        all_entries.append(pickle.load(f))
    tensor_bundle.MergeShardIndex(prefix, all_entries, self._num_shards)
    for r in range(self._num_shards):
      try:
        os.remove(self._SidecarPath(prefix, r, seq))
      except OSError:
        pass
    self._UpdateState(prefix)
    self._GarbageCollect()

  def _UpdateState(self, prefix: str):
    with self._lock:
      allp = [os.path.basename(p) for p in AllCheckpoints(self._logdir)]
      base = os.path.basename(prefix)
      if base not in allp:
        allp.append(base)
      tmp = CheckpointStatePath(self._logdir) + '.tmp'
      with open(tmp, 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % base)
        for p in allp:
          f.write('all_model_checkpoint_paths: "%s"\n' % p)
      os.replace(tmp, CheckpointStatePath(self._logdir))

  def _GarbageCollect(self):
    if not self._keep_latest_n:
      return
    ckpts = AllCheckpoints(self._logdir)
    excess = ckpts[:-self._keep_latest_n] if self._keep_latest_n > 0 else []
    for p in excess:
      try:
        mtime = os.path.getmtime(p + '.index')
      except OSError:
        continue
      if self._keep_every_n_hours and (
          mtime - self._last_kept_time >= self._keep_every_n_hours * 3600):
        self._last_kept_time = mtime
        continue
      for f in glob.glob(p + '.*'):
        try:
          os.remove(f)
        except OSError:
          pass

  def Wait(self):
    t = self._thread
    if t is not None:
      t.join()
      self._thread = None
    if self._error is not None:
      e, self._error = self._error, None
      raise e

  def Save(self, global_step: int, tensors: Optional[Dict[str, torch.Tensor]] = None,
           prefix: Optional[str] = None) -> str:
    """Snapshots (device → host) synchronously; writes sync or async."""
    self.Wait()
    tensors = tensors if tensors is not None else self._vars_fn()
    self._DoSanityChecks(tensors)
    snapshot = {k: _ToNumpy(v) for k, v in tensors.items()}
    path = '%s-%08d' % (prefix or self._prefix, int(global_step))
    self._seq = getattr(self, '_seq', 0) + 1
    seq = self._seq
    if not self._async:
      self._Write(path, snapshot, global_step, seq)
      return path

    def run():
      try:
        self._Write(path, snapshot, global_step, seq)
      except BaseException as e:  # pylint: disable=broad-except
        self._error = e

    self._thread = threading.Thread(target=run, name='async_ckpt', daemon=True)
    self._thread.start()
    return path


def _SaverRestore(self, sess=None, path: Optional[str] = None,
                  checkpoint_basename: Optional[str] = None, strict: bool = True):
  """Loads the bundle at `path` (default: the latest in the log dir) into the tensors
  `variables_fn()` returns, in place. Returns (global_step or None, path) (ref :420)."""
  del sess, checkpoint_basename
  from lingvo_b200.utils import tensor_bundle as tb  # pylint: disable=g-import-not-at-top
  self.Wait()
  path = path or LatestCheckpoint(self._logdir)
  if path is None:
    raise FileNotFoundError('No checkpoint under %s' % self._logdir)
  reader = tb.BundleReader(path)
  keys = set(reader.Keys())
  with torch.no_grad():
    for name, t in self._vars_fn().items():
      if name not in keys:
        if strict:
          raise KeyError('%s not found in %s' % (name, path))
        continue
      t.copy_(FromNumpy(reader.Read(name)).to(device=t.device, dtype=t.dtype))
  step = int(reader.Read('global_step')) if 'global_step' in keys else None
  reader.Close()
  return step, path


Saver.Restore = _SaverRestore
Saver.Sync = Saver.Wait


def WriteNpArrays(file_prefix: str, nmap) -> None:
  """Writes a NestedMap of numpy arrays as a TF tensor bundle keyed by the flattened
  NestedMap paths (ref :574)."""
  from lingvo_b200.utils import tensor_bundle as tb  # pylint: disable=g-import-not-at-top
  w = tb.BundleWriter(file_prefix)
  for k, v in sorted(nmap.FlattenItems()):
    assert isinstance(v, np.ndarray), (k, type(v))
    w.Add(k, v)
  w.Finish()


def ReadNpArrays(file_prefix: str, nmap):
  """Reads the bundle back into the structure of `nmap` (a NestedMap of numpy dtypes or
  arrays, whose dtypes the result is cast to) (ref :605)."""
  from lingvo_b200.utils import tensor_bundle as tb  # pylint: disable=g-import-not-at-top
  reader = tb.BundleReader(file_prefix)
  vals = []
  for k, spec in nmap.FlattenItems():
    arr = reader.Read(k)
    dtype = spec.dtype if isinstance(spec, np.ndarray) else np.dtype(spec)
    vals.append(np.asarray(arr).astype(dtype, copy=False))
  reader.Close()
  return nmap.Pack(vals)
