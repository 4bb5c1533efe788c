"""Pipeline-parallel engine: GPipe schedule over ranks with P2P transfers.

One rank per stage (one process per GPU). Activations move forward and
gradients move backward as `torch.distributed` point-to-point transfers —
NCCL over NVLink on B200, gloo in the CPU tests (ref
`lingvo/core/gpipe.py:421-600` + `recurrent.py:1142-1400` Send/Recv links).

Usage on every rank:
    eng = PipelineEngine(group)            # stage id = rank in group
    pipe_layer.AttachEngine(eng)
    out = pipe_layer.FProp(theta, x)       # only the last stage gets real outputs
    loss = f(out)                          # last stage
    eng.Backward(loss)                     # all ranks call this

Schedule: all micro-batch forwards, then all backwards in reverse order
(GPipe). Each stage keeps its micro-batch autograd graphs alive between the two
phases; `remat=True` re-runs the stage forward in backward instead
(activation memory O(1) per micro-batch, like the reference's recompute).
"""

from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def _Meta(t):
  return (tuple(t.shape), t.dtype)


class PipelineEngine:

  def __init__(self, group=None, remat: bool = False):
    self.group = group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    self.remat = remat
    self._saved = []        # per micro-batch: (inputs requiring grad, outputs)
    self._shapes_cache = {}

  @property
  def is_first(self):
    return self.rank == 0

  @property
  def is_last(self):
    return self.rank == self.world - 1

  def _Peer(self, r):
    return dist.get_global_rank(self.group, r) if self.group is not None else r

  # -- tensor tuple transport ---------------------------------------------------
  def _SendTuple(self, tensors, dst):
    metas = [_Meta(t) for t in tensors]
    dist.send_object_list([metas], self._Peer(dst), group=self.group) \
        if not self._KnownMeta(dst, 'send', metas) else None
    for t in tensors:
      dist.send(t.contiguous(), self._Peer(dst), group=self.group)

  def _RecvTuple(self, src, device):
    key = (src, 'recv')
    metas = self._shapes_cache.get(key)
    if metas is None:
      box = [None]
      dist.recv_object_list(box, self._Peer(src), group=self.group)
      metas = box[0]
      self._shapes_cache[key] = metas
    out = []
    for shape, dtype in metas:
      t = torch.empty(shape, dtype=dtype, device=device)
      dist.recv(t, self._Peer(src), group=self.group)
      out.append(t)
    return out

  def _KnownMeta(self, peer, kind, metas):
    key = (peer, kind)
    if self._shapes_cache.get(key) == metas:
      return True
    self._shapes_cache[key] = metas
    return False

  # -- schedule -----------------------------------------------------------------
  def Forward(self, layer, theta, micro_inputs: List[tuple]):
    """Runs this rank's stage on every micro-batch. Returns per-micro-batch
    outputs (real tensors on the last stage, detached placeholders elsewhere)."""
    assert layer.num_stages == self.world, (
        'PipelineEngine: %d cells but %d ranks' % (layer.num_stages, self.world))
    self._saved = []
    dev = None
    for t in micro_inputs[0]:
      if isinstance(t, torch.Tensor):
        dev = t.device
        break
    outs = []
    for mi in micro_inputs:
      if self.is_first:
        ins = list(mi)
      else:
        ins = self._RecvTuple(self.rank - 1, dev)
        for t in ins:
          if t.is_floating_point():
            t.requires_grad_(True)
      out = layer._RunCells(theta, tuple(ins), self.rank, self.rank + 1)  # pylint: disable=protected-access
      if not self.is_last:
        self._SendTuple([o.detach() for o in out], self.rank + 1)
      self._saved.append((ins, out))
      outs.append(out)
    return outs

  def Backward(self, loss=None):
    """Backward phase; the last stage passes the (already micro-batch-merged)
    loss, other stages pass nothing."""
    n = len(self._saved)
    if self.is_last:
      assert loss is not None
      # One backward through the merged loss populates grads of every micro-batch
      # input on this stage.
      loss.backward()
      for m in reversed(range(n)):
        ins, _ = self._saved[m]
        if not self.is_first:
          grads = [t.grad if t.grad is not None else torch.zeros_like(t)
                   for t in ins if t.is_floating_point()]
          self._SendTuple(grads, self.rank - 1)
    else:
      dev = self._saved[0][1][0].device
      for m in reversed(range(n)):
        ins, out = self._saved[m]
        gouts = self._RecvTuple(self.rank + 1, dev)
        fl = [o for o in out if o.is_floating_point() and o.requires_grad]
        torch.autograd.backward(fl, gouts[:len(fl)])
        if not self.is_first:
          grads = [t.grad if t.grad is not None else torch.zeros_like(t)
                   for t in ins if t.is_floating_point()]
          self._SendTuple(grads, self.rank - 1)
    self._saved = []
