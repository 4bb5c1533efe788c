"""Data-parallel gradient sync over NVLink peer memory.

Two engines, both built on `comm_kernels.cu` + the flag protocol:

* `FusedAllReduce` — flat bf16 gradient buffer in symmetric memory; one
  two-shot all-reduce kernel (peer loads, fp32 accumulate, scale, peer
  stores); the (fused) optimizer then runs replicated. Used with Adafactor,
  whose factored statistics need whole rows/columns.
* `ZeroAdam` — reduce-scatter (+ global-norm partial) → device-side clip
  scale → **one kernel** doing partitioned Adam on the fp32 master shard and
  storing the new bf16 weights into every peer's parameter replica
  (all-gather fused). Optimizer state and master weights are sharded 1/W.
"""

from __future__ import annotations

import math
from typing import List

import torch
import torch.distributed as dist

from lingvo_b200 import ops
from lingvo_b200.core import py_utils
from lingvo_b200.parallel import symm as symm_lib


def _ReplicatedVars(task) -> List[torch.nn.Parameter]:
  vs = [v for v in task.vars.Flatten()
        if v.requires_grad and not getattr(v, 'expert_parallel', False)]
  return sorted(vs, key=lambda v: v.var_name)


class _Flat:
  """Layout of variables inside a flat buffer (each slot 8-element aligned)."""

  def __init__(self, variables, world):
    self.vars = variables
    self.offsets = []
    off = 0
    for v in variables:
      self.offsets.append(off)
      off += (v.numel() + 7) // 8 * 8
    unit = world * 8
    self.total = (off + unit - 1) // unit * unit
    self.shard = self.total // world

  def Views(self, buf):
    return [buf[o:o + v.numel()].view(v.shape)
            for v, o in zip(self.vars, self.offsets)]


class _Channels:
  """Flag channels (signal all peers / wait for all peers)."""

  def __init__(self, arena, world, rank, n=8):
    self.world, self.rank = world, rank
    self.off = arena.Alloc(n * world * 4)
    self.flags = arena.Local(self.off, (n * world,), torch.int32)
    self.peer_flags = arena.PeerPtrs(self.off)
    # Sequence numbers live on the device (bumped by the sync kernel itself), so a Sync
    # has no step-dependent launch argument and replays correctly from a CUDA graph.
    self.counters = torch.zeros(n, dtype=torch.int32, device=self.flags.device)

  def Sync(self, ch):
    ops.native().moe_sync(self.peer_flags, self.flags, self.counters, self.world,
                          self.rank, ch)


class FusedAllReduce:
  """`learner.grad_sync` implementation: mean all-reduce of replicated grads."""

  def __init__(self, task, ctx):
    self.ctx = ctx
    self.world, self.rank = ctx.world, ctx.rank
    dev = task.Device()
    self.flat = _Flat(_ReplicatedVars(task), self.world)
    self.arena = symm_lib.SymmArena(self.flat.total * 2 + 65536, dev)
    self.goff = self.arena.Alloc(self.flat.total * 2)
    self.gbuf = self.arena.Local(self.goff, (self.flat.total,), torch.bfloat16)
    self.gviews = self.flat.Views(self.gbuf)
    self.peers = torch.tensor([b + self.goff for b in self.arena.peer_base],
                              dtype=torch.int64)
    self.chan = _Channels(self.arena, self.world, self.rank)
    self._index = {id(v): i for i, v in enumerate(self.flat.vars)}
    dist.barrier()

  def __call__(self, var_grads):
    leaves = [vg for vg in var_grads.Flatten()
              if isinstance(vg, py_utils.VarGrad) and vg.grad is not None and
              id(vg.var) in self._index]
    dst = [self.gviews[self._index[id(vg.var)]] for vg in leaves]
    with torch.no_grad():
      torch._foreach_copy_(dst, [vg.grad for vg in leaves])   # cast → bf16 slots
      self.chan.Sync(0)                                       # all copies staged
      ops.native().allreduce_mean_bf16(
          self.peers, self.flat.shard, self.rank, self.world, 1.0 / self.world,
          self.gbuf.device.index, True, None)
      self.chan.Sync(1)                                       # all shards written
    for vg, view in zip(leaves, dst):
      vg.grad = view
    from lingvo_b200.parallel import dp as dp_lib  # pylint: disable=g-import-not-at-top
    dp_lib._ReduceExpertReplicas(   # pylint: disable=protected-access
        [vg for vg in var_grads.Flatten() if isinstance(vg, py_utils.VarGrad) and
         vg.grad is not None], self.ctx)
    return var_grads


class ZeroAdam:
  """Takes over a learner whose optimizer is Adam (sharded states, fused RS+Adam+AG)."""

  def __init__(self, task, learner, ctx):
    self.ctx = ctx
    self.world, self.rank = ctx.world, ctx.rank
    self.learner = learner
    self.opt_p = learner.optimizer.params
    dev = task.Device()
    self.flat = _Flat(_ReplicatedVars(task), self.world)
    n, shard = self.flat.total, self.flat.shard
    self.arena = symm_lib.SymmArena(n * 4 + 65536, dev)
    self.goff = self.arena.Alloc(n * 2)
    self.poff = self.arena.Alloc(n * 2)
    self.gbuf = self.arena.Local(self.goff, (n,), torch.bfloat16)
    self.pbuf = self.arena.Local(self.poff, (n,), torch.bfloat16)
    self.gviews = self.flat.Views(self.gbuf)
    self.pviews = self.flat.Views(self.pbuf)
    self.gpeers = torch.tensor([b + self.goff for b in self.arena.peer_base],
                               dtype=torch.int64)
    self.gself = torch.tensor([self.arena.peer_base[self.rank] + self.goff],
                              dtype=torch.int64)
    self.ppeers = torch.tensor([b + self.poff for b in self.arena.peer_base],
                               dtype=torch.int64)
    self.chan = _Channels(self.arena, self.world, self.rank)
    # fp32 master shard + Adam moments for my 1/W slice of the flat space.
    full = torch.zeros(n, dtype=torch.float32, device=dev)
    for v, o in zip(self.flat.vars, self.flat.offsets):
      full[o:o + v.numel()] = v.data.reshape(-1).float()
    self.master = full[self.rank * shard:(self.rank + 1) * shard].clone()
    self.m = torch.zeros_like(self.master)
    self.v = torch.zeros_like(self.master)
    self.pbuf.copy_(full)
    del full
    # The model computes with the replica views; gradients flow to them.
    for var, view in zip(self.flat.vars, self.pviews):
      var.compute = view.requires_grad_(True)
    for _, layer in task.Walk():
      for name, var in layer._private_vars.items():  # pylint: disable=protected-access
        if getattr(var, 'compute', None) is not None:
          layer.SetThetaOverride(name, var.compute)
    self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
    self.step = 0
    self._index = {id(v): i for i, v in enumerate(self.flat.vars)}
    dist.barrier()

  def Apply(self, lr, var_grads, clip_norm: float = 0.0):
    """Fused step. Returns the global gradient norm (device scalar)."""
    p = self.opt_p
    nat = ops.native()
    leaves = [vg for vg in var_grads.Flatten()
              if isinstance(vg, py_utils.VarGrad) and vg.grad is not None and
              id(vg.var) in self._index]
    dst = [self.gviews[self._index[id(vg.var)]] for vg in leaves]
    with torch.no_grad():
      torch._foreach_copy_(dst, [vg.grad for vg in leaves])
      self.chan.Sync(0)
      self.sumsq.zero_()
      nat.allreduce_mean_bf16(self.gpeers, self.flat.shard, self.rank,
                              self.world, 1.0 / self.world,
                              self.gbuf.device.index, False, self.sumsq)
      dist.all_reduce(self.sumsq)                       # 4-byte scalar
      gnorm = self.sumsq.sqrt()
      if clip_norm:
        scale = torch.clamp(clip_norm / gnorm, max=1.0)
      else:
        scale = torch.ones_like(gnorm)
      scale = torch.where(torch.isfinite(gnorm), scale, torch.zeros_like(scale))
      self.step += 1
      t = self.step
      lr_t = float(lr) * math.sqrt(1 - p.beta2**t) / (1 - p.beta1**t)
      nat.zero_adam(self.gself, self.ppeers, self.master, self.m, self.v,
                    self.rank, self.world, 1, 1.0, scale, lr_t, p.beta1,
                    p.beta2, p.epsilon)
      self.chan.Sync(1)
    return gnorm

  # -- checkpoint hooks (called through core/train_engine.py) -------------------------------
  def _Full(self, shard):
    full = torch.empty(self.flat.total, dtype=torch.float32, device=shard.device)
    dist.all_gather_into_tensor(full, shard)
    return full

  def PreSave(self):
    """Makes `var.data` (fp32 masters) and the optimizer's `<var>/Adam`, `<var>/Adam_1`
    slots authoritative on every rank, so the checkpoint has the reference layout."""
    from lingvo_b200.core import optimizer as optimizer_lib  # pylint: disable=g-import-not-at-top
    self.GatherMasters()
    opt = self.learner.optimizer
    fm, fv = self._Full(self.m), self._Full(self.v)
    with torch.no_grad():
      for var, o in zip(self.flat.vars, self.flat.offsets):
        n = var.numel()
        slots = opt._slots.setdefault(optimizer_lib._VarKey(var), {})  # pylint: disable=protected-access
        slots['m'] = fm[o:o + n].view(var.shape).clone()
        slots['v'] = fv[o:o + n].view(var.shape).clone()
    opt._step_count = self.step   # pylint: disable=protected-access

  def PostSave(self):
    """Drops the gathered full-size moments again (they only exist for the snapshot)."""
    from lingvo_b200.core import optimizer as optimizer_lib  # pylint: disable=g-import-not-at-top
    opt = self.learner.optimizer
    for var in self.flat.vars:
      slots = opt._slots.get(optimizer_lib._VarKey(var), {})  # pylint: disable=protected-access
      slots.pop('m', None)
      slots.pop('v', None)

  def PostRestore(self):
    """Re-shards masters / moments / step from the restored variables and slots."""
    from lingvo_b200.core import optimizer as optimizer_lib  # pylint: disable=g-import-not-at-top
    opt = self.learner.optimizer
    n, shard = self.flat.total, self.flat.shard
    dev = self.master.device
    lo, hi = self.rank * shard, (self.rank + 1) * shard
    full = torch.zeros(n, dtype=torch.float32, device=dev)
    fm, fv = torch.zeros_like(full), torch.zeros_like(full)
    with torch.no_grad():
      for var, o in zip(self.flat.vars, self.flat.offsets):
        k = var.numel()
        full[o:o + k] = var.data.reshape(-1).float()
        slots = opt._slots.get(optimizer_lib._VarKey(var), {})  # pylint: disable=protected-access
        if 'm' in slots:
          fm[o:o + k] = slots['m'].reshape(-1).float().to(dev)
        if 'v' in slots:
          fv[o:o + k] = slots['v'].reshape(-1).float().to(dev)
      self.master.copy_(full[lo:hi])
      self.m.copy_(fm[lo:hi])
      self.v.copy_(fv[lo:hi])
      self.pbuf.copy_(full)
    self.step = int(opt._step_count)   # pylint: disable=protected-access

  def GatherMasters(self):
    """fp32 masters of all shards → `var.data` (for checkpointing)."""
    full = torch.empty(self.flat.total, dtype=torch.float32,
                       device=self.master.device)
    dist.all_gather_into_tensor(full, self.master)
    with torch.no_grad():
      for v, o in zip(self.flat.vars, self.flat.offsets):
        v.data.copy_(full[o:o + v.numel()].view(v.shape))
