"""Asynchronous data parallelism (SURVEY §2.5 "DP async / parameter server"; ref: the
`--mode=async` trainers of `runners.py:192-360`, where every trainer replica pulls possibly
stale variables from parameter servers, computes a gradient and applies it without waiting
for the other replicas).

There is no parameter server when every GPU owns its replica, so the asynchronous regime is
expressed directly between the replicas — *local steps with delayed, non-blocking parameter
averaging*:

  * every rank applies **its own** gradients immediately; no collective sits between
    backward and the optimizer, so a slow rank never stalls a fast one inside a step;
  * every `sync_every` steps a rank snapshots its parameters into a flat buffer and launches
    a **non-blocking all-reduce** of that snapshot (NCCL on its own stream, overlapped with
    the following training steps);
  * when that all-reduce has finished (it is polled at the next reconciliation point — at
    most `sync_every` steps later, the *bounded staleness* of the scheme) the rank folds the
    other replicas' progress in:   p ← p + (mean(snapshots) − own snapshot).
    The correction is exactly what a parameter server would have added: the average of the
    *other* workers' updates, `sync_every` steps stale.

With `sync_every=1` this is one-step-delayed averaging; larger values trade gradient
staleness for less NVLink traffic (one flat all-reduce of the parameters per `sync_every`
steps instead of one of the gradients per step). `Finalize()` (called before checkpoints and
at the end of training) blocks and makes all replicas identical.

Expert-parallel and tensor-parallel shards are left alone (they are not replicas).
"""

from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from lingvo_b200.parallel import mesh as mesh_lib


class AsyncDataParallel:
  """Delayed parameter averaging between data-parallel replicas."""

  def __init__(self, task, sync_every: int = 1, group=None):
    self.task = task
    self.sync_every = max(1, int(sync_every))
    self.group = group
    self.world = dist.get_world_size(group)
    self._vars: List[torch.nn.Parameter] = [
        v for v in task.vars.Flatten()
        if v.requires_grad and not getattr(v, 'expert_parallel', False) and
        getattr(v, 'tp_shard', None) is None]
    self._numel = sum(v.numel() for v in self._vars)
    dev = self._vars[0].device if self._vars else torch.device('cpu')
    self._snapshot = torch.zeros(self._numel, dtype=torch.float32, device=dev)
    self._reduced = torch.zeros(self._numel, dtype=torch.float32, device=dev)
    self._work: Optional[dist.Work] = None
    self._steps = 0
    self.num_reconciliations = 0
    # identical starting point (rank 0 wins), as in the synchronous mode
    with torch.no_grad():
      for v in self._vars:
        dist.broadcast(v.data, src=dist.get_global_rank(group, 0) if group is not None else 0,
                       group=group)

  # -- flat views ---------------------------------------------------------------------
  def _Gather(self, out):
    off = 0
    for v in self._vars:
      n = v.numel()
      out[off:off + n].copy_(v.data.reshape(-1))
      off += n

  def _ApplyCorrection(self):
    """p += mean(snapshots) − own snapshot, then refresh the bf16 compute copies."""
    from lingvo_b200.core import py_utils   # pylint: disable=g-import-not-at-top
    delta = self._reduced.div_(self.world).sub_(self._snapshot)
    off = 0
    with torch.no_grad():
      for v in self._vars:
        n = v.numel()
        v.data.add_(delta[off:off + n].view_as(v.data).to(v.dtype))
        off += n
      py_utils.RefreshComputeCopies(self._vars)
    self.num_reconciliations += 1

  # -- protocol -------------------------------------------------------------------------
  def PostStep(self):
    """Call after every optimizer step."""
    self._steps += 1
    if self._steps % self.sync_every:
      return
    with torch.no_grad():
      if self._work is not None:
        self._work.wait()                       # launched sync_every steps ago
        self._work = None
        self._ApplyCorrection()
      self._Gather(self._snapshot)
      self._reduced.copy_(self._snapshot)
      self._work = dist.all_reduce(self._reduced, group=self.group, async_op=True)

  def Finalize(self):
    """Blocks until replicas are identical (before a checkpoint / at the end)."""
    with torch.no_grad():
      if self._work is not None:
        self._work.wait()
        self._work = None
        self._ApplyCorrection()
      self._Gather(self._snapshot)
      self._reduced.copy_(self._snapshot)
      dist.all_reduce(self._reduced, group=self.group)
      self._ApplyCorrection()

  # checkpoint hooks used by TrainEngine / Checkpointer
  PreSave = Finalize

  def PostSave(self):
    pass

  def PostRestore(self):
    self._work = None


def Attach(task, sync_every: int = 1):
  """Switches `task` to asynchronous data parallelism; returns the engine (None on 1 rank)."""
  ctx = mesh_lib.Get()
  if ctx.world <= 1:
    return None
  group = ctx.dp_group if ctx.tp_size > 1 else None
  if ctx.tp_size > 1 and ctx.dp_size <= 1:
    return None
  for lrn in task.learners:
    lrn.grad_sync = None                         # gradients stay local
  return AsyncDataParallel(task, sync_every=sync_every, group=group)
