"""Test support (ref `lingvo/core/test_utils.py`): `TestCase` with tensor-aware
assertions, deterministic seeding helpers, `CompareToGoldenSingleFloat`, and
`ComputeNumericGradient` for finite-difference checks."""
import os
import contextlib
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

  def GetScalarSummaryValues(self, logdir, tags=None, is_tf2_writer=False):
    """{tag: {step: value}} read back from the event files under `logdir` (ref :354)."""
    del is_tf2_writer
    import glob  # pylint: disable=g-import-not-at-top
    from lingvo_b200.utils import tfevents  # pylint: disable=g-import-not-at-top
    out = {}
    for path in sorted(glob.glob(os.path.join(str(logdir), 'events.out.tfevents.*'))):
      for step, tag, value in tfevents.ReadScalars(path):
        if tags and tag not in tags:
          continue
        out.setdefault(tag, {})[step] = value
    return out

  GetScalarSummaryValuesTF2 = GetScalarSummaryValues

  def session(self, *args, **kwargs):  # pylint: disable=invalid-name
    """Kept for source compatibility with `with self.session():` blocks."""
    import contextlib  # pylint: disable=g-import-not-at-top
    return contextlib.nullcontext()


UPDATE_GOLDENS = bool(int(os.environ.get('LINGVO_UPDATE_GOLDENS', '0')))


def _SkipIf(test_func, cond, msg):
  """Wraps a test method so it is skipped with `msg` when `cond()` holds (ref :47)."""
  import functools  # pylint: disable=g-import-not-at-top

  @functools.wraps(test_func)
  def _Wrap(self, *args, **kwargs):
    if cond():
      self.skipTest(msg)
    return test_func(self, *args, **kwargs)

  return _Wrap


def SkipIfEager(test_func):
  """Everything here runs eagerly: graph-only tests are skipped."""
  return _SkipIf(test_func, py_utils.IsEagerMode,
                 'Not compatible with eager execution, skipping.')


def SkipIfNonEager(test_func):
  return _SkipIf(test_func, lambda: not py_utils.IsEagerMode(),
                 'Not compatible with graph mode, skipping.')


class TapeIfEager(contextlib.AbstractContextManager):
  """Gradient bookkeeping with the tape interface (ref :69): `watch(x)` marks a tensor as
  a differentiation target, `gradient(ys, xs)` returns d(sum ys)/dxs via autograd."""

  def __init__(self, **kwargs):
    del kwargs

  def __exit__(self, *exc):
    return None

  def watch(self, tensor):  # pylint: disable=invalid-name
    for t in (tensor if isinstance(tensor, (list, tuple)) else [tensor]):
      if isinstance(t, torch.Tensor) and t.is_floating_point() and t.is_leaf:
        t.requires_grad_(True)

  def gradient(self, target, sources, **kwargs):  # pylint: disable=invalid-name
    del kwargs
    single = not isinstance(sources, (list, tuple))
    srcs = [sources] if single else list(sources)
    ys = target if isinstance(target, (list, tuple)) else [target]
    total = sum(y.sum() for y in ys)
    grads = torch.autograd.grad(total, srcs, allow_unused=True, retain_graph=True)
    return grads[0] if single else list(grads)


def DefineAndTrace(*tensor_specs):
  """Decorator that calls the function once on example inputs built from the specs — a
  `torch.Tensor` is used as is, a `(shape, dtype)` pair becomes zeros — and returns the
  result (ref :144: trace-and-run in one place, whatever the execution mode)."""
  def _Example(spec):
    if isinstance(spec, torch.Tensor):
      return spec
    if isinstance(spec, (tuple, list)) and len(spec) == 2 and isinstance(spec[1], torch.dtype):
      return torch.zeros([d or 1 for d in spec[0]], dtype=spec[1])
    if isinstance(spec, dict):
      return type(spec)({k: _Example(v) for k, v in spec.items()})
    return spec

  def _Decorator(fn):
    return fn(*[_Example(s) for s in tensor_specs])

  return _Decorator


def DisableEagerAdapter():
  """No-op: there is no graph-mode session to adapt."""


def DisableTestLevelVariableStore():
  """No-op: variables live on their layers, not in a per-test global store."""


def _ReplaceOneLineInFile(fpath, linenum, old, new):
  lines = open(fpath).readlines()
  assert lines[linenum] == old, ('Expected "%s" at line %d in file %s, but got "%s"' %
                                 (lines[linenum], linenum + 1, fpath, old))
  lines[linenum] = new
  with open(fpath, 'w') as f:
    f.writelines(lines)


def ReplaceGoldenStackAnalysis(new_float_value):
  """Finds the one-line `CompareToGoldenSingleFloat(...)` call site on the stack and
  returns (file, 0-based line number, old line, line with the new golden) (ref :418)."""
  import inspect  # pylint: disable=g-import-not-at-top
  frame = None
  for fr in inspect.stack():
    if fr.code_context and 'CompareToGoldenSingleFloat' in fr.code_context[0] and \
        'def CompareToGoldenSingleFloat' not in fr.code_context[0]:
      frame = fr
      break
  assert frame is not None
  old_line = frame.code_context[0]
  return (frame.filename, frame.lineno - 1, old_line,
          ReplaceGoldenSingleFloat(old_line, new_float_value))


def CompareToGoldenSingleFloat(testobj, v1, v2, *args, **kwargs):
  """assertAllClose(golden v1, value v2); with LINGVO_UPDATE_GOLDENS=1 a mismatching golden
  on a one-line call site is rewritten in the test file instead (ref :434)."""
  if not UPDATE_GOLDENS:
    testobj.assertAllClose(v1, v2, *args, **kwargs)
  else:
    _ReplaceOneLineInFile(*ReplaceGoldenStackAnalysis(float(v2)))


def PickEveryN(np_arr, step=1):
  """Flattens `np_arr` and keeps one value every `step` values."""
  return np.asarray(np_arr).flatten()[::step]


def ComputeNumericGradientEager(fy, x, delta=1e-4, step=1):
  """Central differences of `sum(fy(x))` w.r.t. every `step`-th element of x; the other
  entries are 0. `fy` takes a numpy array (or tensor) shaped like x (ref :537)."""
  x_orig = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.array(x)
  grad = np.zeros(x_orig.size, dtype=x_orig.dtype)

  def _Eval(v):
    arg = torch.from_numpy(v) if isinstance(x, torch.Tensor) else v
    y = fy(arg)
    return float(y.sum()) if hasattr(y, 'sum') else float(y)

  for i in range(0, x_orig.size, step):
    pos, neg = x_orig.copy(), x_orig.copy()
    pos.reshape(-1)[i] += delta
    neg.reshape(-1)[i] -= delta
    grad[i] = (_Eval(pos) - _Eval(neg)) / (2 * delta)
  out = grad.reshape(x_orig.shape)
  return torch.from_numpy(out) if isinstance(x, torch.Tensor) else out


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
