"""Process-group topology for explicit SPMD on NVLink/NVSwitch.

One process per GPU (`torchrun`). `ParallelContext` names the groups every
explicit collective in the framework runs on:

  * `dp`  — data parallel (all ranks unless `tp`/`pp` carve out axes);
  * `ep`  — expert parallel: the first `min(world, E)` ranks … (by default
            every rank) each own `E / ep_size` experts; tokens cross ranks via
            the fused P2P dispatch/combine kernels (`parallel/ep.py`);
  * `tp`  — tensor parallel (GEMM + reduce-scatter / all-gather + GEMM).

Reference analogue: `device_mesh` + `tensor_split_dims_mapping` (§2.5).
"""

from __future__ import annotations

import os
import threading
from typing import Optional

import torch
import torch.distributed as dist

_LOCK = threading.Lock()
_CTX = None


class ParallelContext:

  def __init__(self, ep_size: Optional[int] = None, tp_size: int = 1,
               mode: Optional[str] = None):
    self.initialized = dist.is_available() and dist.is_initialized()
    self.world = dist.get_world_size() if self.initialized else 1
    self.rank = dist.get_rank() if self.initialized else 0
    self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    env_tp = int(os.environ.get('LINGVO_B200_TP', '0') or 0)
    self.tp_size = max(int(tp_size or 1), env_tp, 1)
    assert self.world % self.tp_size == 0, (
        'tp_size %d must divide the world size %d' % (self.tp_size, self.world))
    self.ep_size = ep_size
    # 'fused' = hand-written peer-memory kernels; 'nccl' = stock baseline.
    self.mode = mode or os.environ.get('LINGVO_B200_COMM', 'fused')
    self._ep_engines = {}
    self._tp_group = None
    self._dp_group = None
    self._groups_built = False

  # -- tensor-parallel × data-parallel mesh --------------------------------------------
  # Rank r has coordinates (dp_rank, tp_rank) = (r // tp, r % tp): a TP group is `tp`
  # consecutive ranks (neighbouring GPUs of one NVSwitch domain), a DP group the ranks with
  # equal tp_rank. This is the reference's `device_mesh` [dp, tp] with mesh axis 1 = model.
  @property
  def tp_rank(self):
    return self.rank % self.tp_size

  @property
  def dp_rank(self):
    return self.rank // self.tp_size

  @property
  def dp_size(self):
    return self.world // self.tp_size

  def _BuildGroups(self):
    """Collective creation: every rank creates every group in the same order."""
    if self._groups_built or not self.initialized:
      return
    self._groups_built = True
    tp, world = self.tp_size, self.world
    if tp == world:
      self._tp_group = dist.group.WORLD
    elif tp > 1:
      for k in range(world // tp):
        g = dist.new_group(list(range(k * tp, (k + 1) * tp)))
        if k == self.dp_rank:
          self._tp_group = g
    if tp == 1:
      self._dp_group = dist.group.WORLD
    elif tp < world:
      for r in range(tp):
        g = dist.new_group(list(range(r, world, tp)))
        if r == self.tp_rank:
          self._dp_group = g

  @property
  def tp_group(self):
    self._BuildGroups()
    return self._tp_group

  @property
  def dp_group(self):
    """Ranks holding replicas of this rank's parameters (None: no replicas, tp == world)."""
    self._BuildGroups()
    return self._dp_group

  @property
  def device(self):
    if torch.cuda.is_available():
      return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def Get() -> ParallelContext:
  global _CTX
  with _LOCK:
    if _CTX is None or (_CTX.initialized != (dist.is_available() and
                                             dist.is_initialized())):
      _CTX = ParallelContext()
    return _CTX


def Reset(**kwargs) -> ParallelContext:
  global _CTX
  with _LOCK:
    _CTX = ParallelContext(**kwargs)
    return _CTX


def ExpertParallelFor(num_experts: int):
  """EP engine for `num_experts`, or None when all experts are local."""
  ctx = Get()
  if ctx.world <= 1 or not num_experts:
    return None
  ep = ctx.ep_size or min(ctx.world, num_experts)
  if ep <= 1:
    return None
  key = (num_experts, ep)
  if key not in ctx._ep_engines:  # pylint: disable=protected-access
    from lingvo_b200.parallel import ep as ep_lib
    ctx._ep_engines[key] = ep_lib.ExpertParallel(ctx, num_experts, ep)  # pylint: disable=protected-access
  return ctx._ep_engines[key]  # pylint: disable=protected-access


def TensorParallel():
  """The parallel context when tensor parallelism is on (tp_size > 1), else None."""
  ctx = Get()
  if ctx.world <= 1 or ctx.tp_size <= 1:
    return None
  return ctx


def ConfigureFromMeshShape(device_mesh_shape):
  """Adopts the reference's `device_mesh_shape = [dp, tp]` (axis 1 = model parallel) when it
  matches the running world; returns the context. A mesh that does not match the world
  (e.g. a TPU config run on one GPU) leaves the topology unchanged."""
  ctx = Get()
  if not device_mesh_shape or len(device_mesh_shape) < 2 or not ctx.initialized:
    return ctx
  n = 1
  for d in device_mesh_shape:
    n *= int(d)
  tp = int(device_mesh_shape[-1])
  if n != ctx.world or tp <= 1 or tp == ctx.tp_size:
    return ctx
  assert not ctx._groups_built, 'tensor-parallel topology changed after groups were built'   # pylint: disable=protected-access
  return Reset(ep_size=ctx.ep_size, tp_size=tp, mode=ctx.mode)
