"""Python facade over the native (C++) input-pipeline primitives.

Reference natives: `RandomPermutationSequence`
(`random_ops_kernels.cc`), `CachedCall` (`functional_ops_kernels.cc`), record
yielder / batcher / generic input (`record_yielder.cc`, `record_batcher.cc`,
`generic_input_op_kernels.cc`). The C++ lives in `csrc/native_input.cpp`; the
pure-python fallbacks here keep CPU-only hosts without the extension working.
"""

import os
import threading
from typing import Dict, List, Sequence

import numpy as np

from lingvo_b200 import ops

_CACHE: Dict[str, Dict[str, np.ndarray]] = {}
_CACHE_LOCK = threading.Lock()


def CachedLoadTensors(path: str, names: Sequence[str]) -> Dict[str, np.ndarray]:
  """Loads named tensors from a data file exactly once per process."""
  key = os.path.abspath(path)
  with _CACHE_LOCK:
    if key not in _CACHE:
      if path.endswith('.npz'):
        with np.load(path) as f:
          _CACHE[key] = {k: f[k] for k in f.files}
      else:
        from lingvo_b200.utils import tensor_bundle
        _CACHE[key] = tensor_bundle.BundleReader(path).ReadAll()
    data = _CACHE[key]
  return {n: data[n] for n in names}


class RandomPermutationSequence:
  """Epoch-wise random permutation batches of [0, num)."""

  def __init__(self, num: int, batch: int, repeat: bool, seed: int = 0):
    try:
      mod = ops.host()           # the C++ class lives in the torch-free host library `_H`
    except Exception:  # pylint: disable=broad-except
      mod = None
    self._impl = None
    if mod is not None and hasattr(mod, 'RandomPermutationSequence'):
      self._impl = mod.RandomPermutationSequence(num, batch, repeat, seed)
    else:
      self._num, self._batch, self._repeat = num, batch, repeat
      self._rng = np.random.RandomState(seed if seed else None)
      self._order = self._rng.permutation(num)
      self._pos = 0

  def Next(self) -> np.ndarray:
    if self._impl is not None:
      out = self._impl.next()
      if out is None or len(out) == 0:
        raise StopIteration()
      return np.asarray(out, dtype=np.int64)
    out = []
    while len(out) < self._batch:
      if self._pos == self._num:
        if not self._repeat:
          break
        self._order = self._rng.permutation(self._num)
        self._pos = 0
      take = min(self._batch - len(out), self._num - self._pos)
      out.extend(self._order[self._pos:self._pos + take])
      self._pos += take
    if not out:
      raise StopIteration()
    return np.asarray(out, dtype=np.int64)
