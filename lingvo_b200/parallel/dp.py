"""Data-parallel gradient synchronisation.

Reference: TPU `cross_replica_sum` per variable (`py_utils.py:2942-3081`),
PS AddN on GPU. Here `Attach(task)` installs `learner.grad_sync`:

* `mode='nccl'`  — bucketed `all_reduce` (stock baseline);
* `mode='fused'` — `parallel/zero.py`: hand-written bucketed reduce-scatter
  over NVLink peer memory fused with cast/scale + the partitioned optimizer
  update, followed by the parameter all-gather.

Expert-parallel variables (`var.expert_parallel`) are owned by exactly one
rank of the EP group and are therefore *not* reduced across it.
"""

from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from lingvo_b200.core import py_utils
from lingvo_b200.parallel import mesh as mesh_lib

_BUCKET_BYTES = 64 << 20


def _AllReduceBuckets(grads: List[torch.Tensor], world: int, group=None):
  """In-place mean all-reduce of `grads` in ~64 MiB flat buckets."""
  bucket, size = [], 0

  def flush():
    if not bucket:
      return
    flat = torch.cat([g.reshape(-1) for g in bucket])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    off = 0
    for g in bucket:
      n = g.numel()
      g.copy_(flat[off:off + n].view_as(g))
      off += n
    bucket.clear()

  for g in grads:
    bucket.append(g)
    size += g.numel() * g.element_size()
    if size >= _BUCKET_BYTES:
      flush()
      size = 0
  flush()


def _ReduceExpertReplicas(leaves, ctx):
  """`ep_size < world`: consecutive EP groups hold replicas of the same experts; their
  gradients are averaged over the ranks with equal `ep_rank`."""
  eps = [vg for vg in leaves if getattr(vg.var, 'expert_parallel', False)]
  if not eps:
    return
  shard = getattr(eps[0].var, 'ep_shard', None)
  if shard is None or shard[1] >= ctx.world:
    return
  ep_size = shard[1]
  group = _ReplicaGroup(ctx, ep_size)
  n_rep = ctx.world // ep_size
  with torch.no_grad():
    by_dtype = {}
    for vg in eps:
      by_dtype.setdefault(vg.grad.dtype, []).append(vg.grad)
    for grads in by_dtype.values():
      _AllReduceBuckets(grads, n_rep, group)


_REPLICA_GROUPS = {}


def _ReplicaGroup(ctx, ep_size):
  key = (ctx.world, ep_size)
  if key not in _REPLICA_GROUPS:
    mine = None
    for r in range(ep_size):
      g = dist.new_group(list(range(r, ctx.world, ep_size)))
      if ctx.rank % ep_size == r:
        mine = g
    _REPLICA_GROUPS[key] = mine
  return _REPLICA_GROUPS[key]


def Attach(task):
  """Installs gradient synchronisation on every learner of `task`."""
  ctx = mesh_lib.Get()
  if ctx.world <= 1:
    return None

  def sync(var_grads):
    leaves = [vg for vg in var_grads.Flatten()
              if isinstance(vg, py_utils.VarGrad) and vg.grad is not None]
    by_dtype = {}
    for vg in leaves:
      if getattr(vg.var, 'expert_parallel', False):
        continue
      by_dtype.setdefault(vg.grad.dtype, []).append(vg.grad)
    with torch.no_grad():
      if ctx.tp_size > 1:
        # [dp, tp] mesh: parameters (sharded or replicated) are replicated only along the dp
        # axis; within a TP group the replicated parameters already carry identical
        # gradients (their inputs are replicated, f/g make the activations' grads whole).
        if ctx.dp_size > 1:
          for grads in by_dtype.values():
            _AllReduceBuckets(grads, ctx.dp_size, ctx.dp_group)
      else:
        for grads in by_dtype.values():
          _AllReduceBuckets(grads, ctx.world)
    _ReduceExpertReplicas(leaves, ctx)
    return var_grads

  # Make replicated variables identical across ranks (rank 0 wins); a tensor-parallel shard
  # is replicated only across the ranks with the same tp_rank (source: dp_rank 0).
  with torch.no_grad():
    for v in task.vars.Flatten():
      if getattr(v, 'expert_parallel', False):
        continue
      if getattr(v, 'tp_shard', None) is not None:
        if ctx.dp_size > 1:
          dist.broadcast(v.data, src=ctx.tp_rank, group=ctx.dp_group)
      else:
        dist.broadcast(v.data, src=0)
  if ctx.tp_size > 1:
    assert not any(getattr(v, 'expert_parallel', False) for v in task.vars.Flatten()), (
        'tensor parallelism and expert parallelism are not combined in one job yet')
    for lrn in task.learners:
      lrn.grad_sync = sync
    return sync
  fused = None
  if ctx.mode == 'fused' and task.Device().type == 'cuda':
    from lingvo_b200.parallel import zero
    # Mixed-precision copies must exist before gradients are produced in bf16.
    has_ep = any(getattr(v, 'expert_parallel', False) for v in task.vars.Flatten())
    adam_only = all(type(l.optimizer).__name__ in ('Adam', 'AdamV2') for l in task.learners)
    if adam_only and not has_ep and len(task.learners) == 1:
      # ZeRO-Adam: optimizer state + fp32 masters sharded 1/W, RS + Adam + AG in one kernel.
      engine = zero.ZeroAdam(task, task.learners[0], ctx)
      task.learners[0].fused_update = engine
      task._mixed_precision_attached = True   # pylint: disable=protected-access
      return engine
    task.EnableMixedPrecision()
    fused = zero.FusedAllReduce(task, ctx)
  for lrn in task.learners:
    lrn.grad_sync = fused if fused is not None else sync
  return fused if fused is not None else sync
