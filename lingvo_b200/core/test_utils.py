"""Test support (ref `lingvo/core/test_utils.py`): `TestCase` with tensor-aware
assertions, deterministic seeding helpers, `CompareToGoldenSingleFloat`, and
`ComputeNumericGradient` for finite-difference checks."""
import os
import re
import unittest

import numpy as np
import torch

from lingvo_b200.core import cluster_factory
from lingvo_b200.core import py_utils


class TestCase(unittest.TestCase):

  def setUp(self):
    super().setUp()
    torch.manual_seed(301)
    np.random.seed(301)

  def _ToNp(self, x):
    return x.detach().cpu().float().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)

  def assertAllClose(self, a, b, rtol=1e-6, atol=1e-6, msg=None):  # pylint: disable=invalid-name
    np.testing.assert_allclose(self._ToNp(a), self._ToNp(b), rtol=rtol, atol=atol,
                               err_msg=msg or '')

  def assertAllEqual(self, a, b, msg=None):  # pylint: disable=invalid-name
    np.testing.assert_array_equal(self._ToNp(a), self._ToNp(b), err_msg=msg or '')

  def SetEval(self, mode=True):
    return cluster_factory.SetEval(mode)

  def session(self, *args, **kwargs):  # pylint: disable=invalid-name
    """Kept for source compatibility with `with self.session():` blocks."""
    import contextlib  # pylint: disable=g-import-not-at-top
    return contextlib.nullcontext()


def CompareToGoldenSingleFloat(testobj, v1, v2, *args, **kwargs):
  testobj.assertAllClose(v1, v2, *args, **kwargs)


def ComputeNumericGradient(fn, x, delta=1e-4, step=1):
  """Central finite differences of scalar `fn(x)` w.r.t. every `step`-th element of x."""
  x = x.detach().clone().double()
  flat = x.reshape(-1)
  grad = torch.zeros_like(flat)
  for i in range(0, flat.numel(), step):
    old = flat[i].item()
    flat[i] = old + delta
    up = float(fn(x.reshape(x.shape)))
    flat[i] = old - delta
    dn = float(fn(x.reshape(x.shape)))
    flat[i] = old
    grad[i] = (up - dn) / (2 * delta)
  return grad.reshape(x.shape)


def ReplaceGoldenSingleFloat(old, float_value):
  m = re.match(r'(?P<prefix>.*)\bCompareToGoldenSingleFloat\(\s*(?P<testobj>[^,]+),\s*'
               r'(?P<old>[-.\d eE]+),\s*(?P<v2>.*)\)(?P<postfix>.*)\n', old)
  if not m:
    return old
  return '%sCompareToGoldenSingleFloat(%s, %f, %s)%s\n' % (
      m.group('prefix'), m.group('testobj'), float_value, m.group('v2'), m.group('postfix'))


def main(*args, **kwargs):  # pylint: disable=invalid-name
  unittest.main(*args, **kwargs)


def FreePort() -> int:
  """A TCP port that is free right now on 127.0.0.1 (for multi-process rendezvous in tests;
  asking the kernel avoids collisions between parallel pytest-xdist workers)."""
  import socket
  with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
    s.bind(('127.0.0.1', 0))
    return int(s.getsockname()[1])


def ToNumpyTree(x):
  """Tensors → numpy arrays, recursively through tuples/lists/dicts. Multi-process tests
  send results through `mp.Queue` as numpy: a torch tensor travels as a shared-memory
  handle that dies with the producer process, numpy arrays are pickled by value."""
  import torch
  if isinstance(x, torch.Tensor):
    return x.detach().cpu().numpy()
  if isinstance(x, dict):
    return {k: ToNumpyTree(v) for k, v in x.items()}
  if isinstance(x, (list, tuple)):
    return type(x)(ToNumpyTree(v) for v in x)
  return x


def ToTorchTree(x):
  """Inverse of `ToNumpyTree`."""
  import numpy as np
  import torch
  if isinstance(x, np.ndarray):
    return torch.from_numpy(x)
  if isinstance(x, dict):
    return {k: ToTorchTree(v) for k, v in x.items()}
  if isinstance(x, (list, tuple)):
    return type(x)(ToTorchTree(v) for v in x)
  return x
