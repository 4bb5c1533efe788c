"""Cross-replica concatenation for in-batch negatives (ref `lingvo/tasks/milan/tpu_utils.py`).

The reference builds the global batch on TPU with `CollectivePermute` rotations. Here every
replica is one process on one GPU: the concat is one NCCL all-gather over NVLink whose
backward returns this replica's slice of the (summed) gradient.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from lingvo_b200.core.nested_map import NestedMap


class _AllGatherWithGrad(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x):
    world = dist.get_world_size()
    outs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(outs, x.contiguous())
    ctx.rank, ctx.n = dist.get_rank(), x.shape[0]
    return torch.cat(outs, 0)

  @staticmethod
  def backward(ctx, g):
    g = g.contiguous()
    dist.all_reduce(g)
    return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n]


def _Distributed():
  return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def CrossReplicaConcat(local_tensor, replica_index=None, num_replicas=None, axis=0):
  """[n, …] on every replica → [world·n, …], replica r's rows at r·n (ref :75)."""
  del replica_index, num_replicas
  if not _Distributed():
    return local_tensor
  x = local_tensor if axis == 0 else local_tensor.transpose(0, axis)
  if x.is_floating_point() and x.requires_grad:
    out = _AllGatherWithGrad.apply(x)
  else:
    outs = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, x.contiguous())
    out = torch.cat(outs, 0)
  return out if axis == 0 else out.transpose(0, axis)


def ConcatenateAcrossReplicas(tensors, tpu_cores=None, axis=0):
  """Applies `CrossReplicaConcat` to every tensor of a NestedMap (ref :25)."""
  del tpu_cores
  return NestedMap(tensors).Transform(lambda t: CrossReplicaConcat(t, axis=axis))


def ReplicaOffset(local_batch_size):
  """Row of the global batch where this replica's examples start."""
  return dist.get_rank() * local_batch_size if _Distributed() else 0
