"""`NestedMap`: attribute-style dict used for theta / batches / states.

Contract per reference `lingvo/core/nested_map.py:81-537` (attr access, dotted
`Get/Set/Has` with `[i]` indexing, Flatten/Pack/Transform/Filter, recursion
into lists *and* dicts). Independent implementation: traversal is a single
generic `_Traverse` generator instead of per-method recursion.
"""

from __future__ import annotations

import re
from typing import Any, Callable, Iterator, List, Optional, Sequence, Tuple

_KEY_RE = re.compile(r'[A-Za-z_][A-Za-z0-9_/]*')
_PATH_TOKEN = re.compile(r'([A-Za-z_][A-Za-z0-9_]*)((?:\[\d+\])*)$')


class _Deleted:
  """Sentinel used by Filter."""


_DELETE = _Deleted()


class NestedMap(dict):
  """A dict with attribute access and structure-aware utilities."""

  _RESERVED = frozenset(dir(dict))
  _HAS_DYNAMIC_ATTRIBUTES = True

  def __init__(self, *args, **kwargs):
    super().__init__(*args, **kwargs)
    for k in self.keys():
      self.CheckKey(k)

  # ---------------------------------------------------------------- basics --
  @staticmethod
  def CheckKey(key):
    if not (isinstance(key, str) and _KEY_RE.fullmatch(key)):
      raise ValueError('Invalid NestedMap key %r' % (key,))

  def __setitem__(self, key, value):
    self.CheckKey(key)
    super().__setitem__(key, value)

  def __setattr__(self, name, value):
    if name in self._RESERVED:
      raise AttributeError('%r is a reserved dict attribute' % name)
    self[name] = value

  def __getattr__(self, name):
    try:
      return self[name]
    except KeyError:
      raise AttributeError('%r; available attributes: %s' %
                           (name, sorted(self.keys()))) from None

  def __delattr__(self, name):
    try:
      del self[name]
    except KeyError:
      raise AttributeError(name) from None

  def __getitem__(self, key):
    try:
      return super().__getitem__(key)
    except KeyError:
      raise KeyError('%r; available attributes: %s' %
                     (key, sorted(self.keys()))) from None

  def copy(self):  # shallow, but keeps the NestedMap type
    return type(self)(self)

  def __deepcopy__(self, memo):
    return self.DeepCopy()

  def DeepCopy(self):
    """Copies containers; leaves are shared."""
    return self.Transform(lambda x: x)

  @staticmethod
  def FromNestedDict(x):
    if isinstance(x, dict):
      out = NestedMap()
      for k, v in x.items():
        out[k] = NestedMap.FromNestedDict(v)
      return out
    if isinstance(x, (list, tuple)):
      return type(x)(NestedMap.FromNestedDict(v) for v in x)
    return x

  def ToNestedDict(self):
    def conv(x):
      if isinstance(x, dict):
        return {k: conv(v) for k, v in x.items()}
      if isinstance(x, list):
        return [conv(v) for v in x]
      return x
    return conv(self)

  # ---------------------------------------------------------- dotted paths --
  @staticmethod
  def _SplitPath(path: str) -> List[Tuple[str, List[int]]]:
    out = []
    for seg in path.split('.'):
      m = _PATH_TOKEN.match(seg)
      if not m:
        raise ValueError('Invalid NestedMap key %r' % path)
      idx = [int(i) for i in re.findall(r'\[(\d+)\]', m.group(2))]
      out.append((m.group(1), idx))
    return out

  def GetItem(self, key: str):
    cur: Any = self
    for name, idxs in self._SplitPath(key):
      cur = cur[name]
      for i in idxs:
        cur = cur[i]
    return cur

  @staticmethod
  def SquareBracketIndex(key: str):
    """`k[3]` → ('k', 3); a plain key → (key, None) (ref :206)."""
    m = re.fullmatch(r'([A-Za-z_][A-Za-z0-9_]*)\[(\d+)\]', key)
    return (m.group(1), int(m.group(2))) if m else (key, None)

  @staticmethod
  def FromNestedDataclass(x):
    """A (nested) dataclass instance → NestedMap (ref :173)."""
    import dataclasses  # pylint: disable=g-import-not-at-top
    if not dataclasses.is_dataclass(x):
      raise ValueError('%s must be a dataclass. Got %s.' % (x, type(x)))
    return NestedMap.FromNestedDict(dataclasses.asdict(x))

  def GetSlice(self, keys):
    """Copy holding only the (possibly nested, `a.b[2].c`-style) `keys` (ref :259)."""
    sliced = NestedMap()
    for k in keys:
      sliced.Set(k, self.GetItem(k))
    return sliced

  def Keys(self) -> List[str]:
    """All leaf keys in nested / array form, in `Flatten` order (ref :276)."""
    return [k for k, _ in self.FlattenItems()]

  def Update(self, other):
    """dict.update for nested keys: adds / replaces every leaf of `other` (ref :360)."""
    for k, v in other.FlattenItems():
      self.Set(k, v)
    return self

  def Union(self, other):
    """A new map with the leaves of self, overridden / extended by `other` (ref :366)."""
    return NestedMap().Update(self).Update(other)

  def Get(self, key: str, default=None):
    try:
      return self.GetItem(key)
    except (KeyError, IndexError, TypeError):
      return default

  def Has(self, key: str) -> bool:
    sentinel = object()
    return self.Get(key, sentinel) is not sentinel

  def Set(self, key: str, value) -> None:
    toks = self._SplitPath(key)
    cur: Any = self
    for ti, (name, idxs) in enumerate(toks):
      last = ti == len(toks) - 1
      if not idxs:
        if last:
          cur[name] = value
          return
        if name not in cur:
          cur[name] = NestedMap()
        if not isinstance(cur[name], dict):
          raise ValueError('Error while setting key %s: sub-key %s is not a '
                           'map' % (key, name))
        cur = cur[name]
        continue
      if name not in cur:
        cur[name] = []
      seq = cur[name]
      if not isinstance(seq, list):
        raise ValueError('Error while setting key %s: %s is not a list' %
                         (key, name))
      for ii, i in enumerate(idxs):
        tail = last and ii == len(idxs) - 1
        if i > len(seq):
          raise ValueError('Error while setting key %s: index %d out of range'
                           % (key, i))
        if tail:
          if i == len(seq):
            seq.append(value)
          else:
            seq[i] = value
          return
        if i == len(seq):
          seq.append([] if ii < len(idxs) - 1 else NestedMap())
        seq = seq[i]
      cur = seq

  # ------------------------------------------------------------- traversal --
  @staticmethod
  def _Traverse(node, prefix: str) -> Iterator[Tuple[str, Any]]:
    if isinstance(node, dict):
      for k in sorted(node.keys()):
        yield from NestedMap._Traverse(node[k], (prefix + '.' + k) if prefix else k)
    elif isinstance(node, list):
      for i, v in enumerate(node):
        yield from NestedMap._Traverse(v, '%s[%d]' % (prefix, i))
    else:
      yield prefix, node

  def FlattenItems(self) -> List[Tuple[str, Any]]:
    return list(self._Traverse(self, ''))

  def Flatten(self) -> List[Any]:
    return [v for _, v in self._Traverse(self, '')]

  @staticmethod
  def _Rebuild(node, fn: Callable[[str, Any], Any], prefix: str):
    if isinstance(node, dict):
      out = type(node)() if isinstance(node, NestedMap) else NestedMap()
      for k in sorted(node.keys()):
        r = NestedMap._Rebuild(node[k], fn, (prefix + '.' + k) if prefix else k)
        if r is not _DELETE:
          dict.__setitem__(out, k, r)
      return out
    if isinstance(node, list):
      res = [NestedMap._Rebuild(v, fn, '%s[%d]' % (prefix, i))
             for i, v in enumerate(node)]
      return [r for r in res if r is not _DELETE]
    return fn(prefix, node)

  def Transform(self, fn: Callable[[Any], Any]):
    return self._Rebuild(self, lambda k, v: fn(v), '')

  def TransformWithKey(self, fn: Callable[[str, Any], Any]):
    return self._Rebuild(self, fn, '')

  def Pack(self, values: Sequence[Any]):
    values = list(values)
    n = len(self.Flatten())
    if len(values) != n:
      raise ValueError('Pack: expected %d values, got %d' % (n, len(values)))
    it = iter(values)
    return self.Transform(lambda _: next(it))

  def IsCompatible(self, other) -> bool:
    if not isinstance(other, dict):
      return False
    return ([k for k, _ in self.FlattenItems()] ==
            [k for k, _ in NestedMap._Traverse(other, '')])

  def Filter(self, fn: Callable[[Any], bool]):
    return self.FilterKeyVal(lambda _, v: fn(v))

  def FilterKeyVal(self, fn: Callable[[str, Any], bool]):
    def prune(node, prefix):
      if isinstance(node, dict):
        out = NestedMap()
        for k in sorted(node.keys()):
          r = prune(node[k], (prefix + '.' + k) if prefix else k)
          if r is not _DELETE:
            dict.__setitem__(out, k, r)
        return out if (out or not prefix) else _DELETE
      if isinstance(node, list):
        res = [prune(v, '%s[%d]' % (prefix, i)) for i, v in enumerate(node)]
        res = [r for r in res if r is not _DELETE]
        return res if res else _DELETE
      return node if fn(prefix, node) else _DELETE
    return prune(self, '')

  def DebugString(self) -> str:
    rows = []
    for k, v in self.FlattenItems():
      shp = getattr(v, 'shape', None)
      if shp is not None and hasattr(v, 'dtype'):
        rows.append('%-50s %-10s %s' % (k, str(v.dtype).split('.')[-1],
                                        tuple(shp)))
      else:
        rows.append('%-50s %r' % (k, v))
    return '\n'.join(rows)

  def VLog(self, level=None, prefix=None):
    import logging
    for line in self.DebugString().split('\n'):
      logging.debug('%s %s', prefix or 'nmap:', line)

  def __repr__(self):
    return 'NestedMap(%s)' % ', '.join(
        '%s=%r' % (k, self[k]) for k in sorted(self.keys()))
