"""Scatter updates x[i] = v (ref `lingvo/core/scatter_update.py`)."""
import contextlib

import torch

from lingvo_b200.core import thread_local_utils

_STACK = thread_local_utils.ThreadLocalStack()


@contextlib.contextmanager
def SetInplaceUpdate(inplace_update):
  _STACK.stack.append(inplace_update)
  try:
    yield
  finally:
    _STACK.stack.pop()


def UseInplaceUpdate():
  return _STACK.stack[-1] if _STACK.stack else True


def Update(x, i, v, *, inplace_update=None):
  """i None: x = v; scalar: x[i] = v; vector: x[i[j]] = v[j]."""
  if inplace_update is None:
    inplace_update = UseInplaceUpdate()
  if i is None:
    if inplace_update and not x.requires_grad:
      x.copy_(v)
      return x
    return v.reshape(x.shape)
  out = x if (inplace_update and not x.requires_grad) else x.clone()
  idx = torch.as_tensor(i, device=x.device)
  if idx.dim() == 0:
    out[int(idx)] = v
  else:
    out[idx.long()] = v
  return out
