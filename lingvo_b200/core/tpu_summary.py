"""Deferred in-step summaries (ref `lingvo/core/tpu_summary.py`).

Layers call `scalar(name, value)` / `tensor(name, value)` during FProp without a
host sync; the trainer collects them with `merge_all()` after the step (one D2H
for everything). `context()` scopes collection to a step.
"""
import contextlib
import threading

import torch

_LOCAL = threading.local()


class TpuSummaryScalar:

  def __init__(self, name, value, reduce='mean'):
    self.name, self.value, self.while_loop_reduce = name, value, reduce


class TpuSummaryContext:

  def __init__(self):
    self.summary_tensors = []
    self.pw_tensors = []


def _Ctx():
  stack = getattr(_LOCAL, 'stack', None)
  return stack[-1] if stack else None


@contextlib.contextmanager
def context(rewrite_while_loop=False, max_loop_vars=2048):  # pylint: disable=invalid-name
  del rewrite_while_loop, max_loop_vars
  if not hasattr(_LOCAL, 'stack'):
    _LOCAL.stack = []
  ctx = TpuSummaryContext()
  _LOCAL.stack.append(ctx)
  try:
    yield ctx
  finally:
    _LOCAL.stack.pop()


def scalar(name, value, while_loop_reduce='mean'):  # pylint: disable=invalid-name
  ctx = _Ctx()
  if ctx is not None:
    ctx.summary_tensors.append(TpuSummaryScalar(name, value, while_loop_reduce))


def tensor(name, value):  # pylint: disable=invalid-name
  ctx = _Ctx()
  if ctx is not None:
    ctx.summary_tensors.append(TpuSummaryScalar(name, value, 'stack'))


def pw_tensor(name, value):  # pylint: disable=invalid-name
  ctx = _Ctx()
  if ctx is not None:
    ctx.pw_tensors.append(TpuSummaryScalar(name, value, 'stack'))


def merge_all():  # pylint: disable=invalid-name
  """name → python float / tensor (moved to host in one go)."""
  ctx = _Ctx()
  if ctx is None:
    return {}
  out = {}
  scalars = [s for s in ctx.summary_tensors if isinstance(s.value, torch.Tensor) and
             s.value.numel() == 1]
  if scalars:
    host = torch.stack([s.value.detach().float().reshape(()) for s in scalars]).cpu()
    for s, v in zip(scalars, host):
      out[s.name] = float(v)
  for s in ctx.summary_tensors:
    if s.name not in out:
      out[s.name] = s.value.detach().cpu() if isinstance(s.value, torch.Tensor) else s.value
  return out


def merge_all_pw_tensor():  # pylint: disable=invalid-name
  ctx = _Ctx()
  return {} if ctx is None else {s.name: s.value for s in ctx.pw_tensors}
