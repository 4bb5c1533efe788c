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
    self.tp_size = tp_size
    self.ep_size = ep_size
    # 'fused' = hand-written peer-memory kernels; 'nccl' = stock baseline.
    self.mode = mode or os.environ.get('LINGVO_B200_COMM', 'fused')
    self._ep_engines = {}

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
