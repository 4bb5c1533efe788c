"""Hierarchical hyper-parameter containers (`Params`, `InstantiableParams`).

Behavioural contract follows the reference `lingvo/core/hyperparams.py`
(Define/Set/Get/Delete :377-519, Freeze, Copy, Visit :686, ToText :784-867,
FromText :869-1015, TextDiff :1036, InstantiableParams :1129-1184), but this
is an independent implementation: values are stored in plain slot records, the
text codec is a small table of (predicate, encoder, decoder) triples, and
dtypes are torch dtypes (serialised with the same short names the reference
uses: ``float32``, ``bfloat16`` …).
"""

from __future__ import annotations

import ast
import copy
import dataclasses
import enum
import importlib
import inspect
import re
import sys
from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple

import numpy as np

try:  # torch is optional for the pure-config layer.
  import torch
  _TORCH_DTYPES = {
      str(v).split('.')[-1]: v
      for v in vars(torch).values()
      if isinstance(v, torch.dtype)
  }
except Exception:  # pragma: no cover
  torch = None
  _TORCH_DTYPES = {}

_NAME_RE = re.compile(r'^[a-z_][a-z0-9_]*$')


def _IsDtype(v) -> bool:
  return torch is not None and isinstance(v, torch.dtype)


def DtypeName(v) -> str:
  return str(v).split('.')[-1]


def DtypeFromName(name: str):
  name = name.strip()
  if name.startswith('torch.'):
    name = name[len('torch.'):]
  alias = {'float': 'float32', 'half': 'float16', 'double': 'float64',
           'int': 'int32', 'long': 'int64'}
  name = alias.get(name, name)
  if name not in _TORCH_DTYPES:
    raise ValueError(f'Unknown dtype name {name!r}')
  return _TORCH_DTYPES[name]


def _Quote(s: str) -> str:
  """Lite quoting: choose the delimiter needing fewer escapes; newlines kept."""
  q = "'" if s.count("'") <= s.count('"') else '"'
  body = s.replace('\\', '\\\\').replace(q, '\\' + q)
  return q + body + q


def _Unquote(s: str) -> str:
  if s and s[0] in '"\'':
    body = s[1:-1] if len(s) >= 2 and s[-1] == s[0] else s[1:]
    return re.sub(r"""\\([\\'"])""", r'\1', body)
  return s


def _TerminalQuote(s: str, q: str) -> bool:
  m = re.search(r'(\\*)%s$' % re.escape(q), s)
  return bool(m) and len(m.group(1)) % 2 == 0


def _IsNamedTuple(x) -> bool:
  return isinstance(x, tuple) and hasattr(x, '_fields')


class _OrderedReprDict(dict):
  """dict whose repr is key-sorted, used for leaf dict values in ToText."""

  def __repr__(self):
    parts = []
    for k in sorted(self):
      v = self[k]
      parts.append('%r: %s' % (k, _Quote(v) if isinstance(v, str) else repr(v)))
    return '{' + ', '.join(parts) + '}'


@dataclasses.dataclass
class _Slot:
  """One named hyper-parameter."""
  name: str
  value: Any
  doc: str
  default: Any = None

  def Clone(self) -> '_Slot':
    return _Slot(self.name, _CloneValue(self.value), self.doc, self.default)

  def GetDefault(self):
    """The value the parameter was defined with (ref :193)."""
    return self.default

  def ToString(self, nested_depth: int) -> str:
    return self.Render(nested_depth)

  def Render(self, depth: int) -> str:
    def _r(v):
      if isinstance(v, Params):
        return v._Render(depth)
      if isinstance(v, dict):
        return '{' + ', '.join(
            '%s: %s' % (k, _r(v[k])) for k in sorted(v, key=str)) + '}'
      if isinstance(v, (list, tuple)) and not _IsNamedTuple(v):
        inner = ', '.join(_r(x) for x in v)
        return ('[%s]' if isinstance(v, list) else '(%s)') % inner
      if isinstance(v, str):
        return '"%s"' % v
      return str(v)

    return '%s%s: %s' % ('  ' * depth, self.name, _r(self.value))


def _CloneValue(v):
  """Deep copy that keeps tensors (and other by-reference handles) shared."""
  if torch is not None and isinstance(v, torch.Tensor):
    return v
  if isinstance(v, Params):
    return v.Copy()
  if isinstance(v, list):
    return [_CloneValue(x) for x in v]
  if isinstance(v, tuple) and not _IsNamedTuple(v):
    return tuple(_CloneValue(x) for x in v)
  if isinstance(v, dict):
    return type(v)((k, _CloneValue(x)) for k, x in v.items())
  if inspect.ismodule(v) or inspect.isclass(v) or inspect.isroutine(v):
    return v
  try:
    return copy.deepcopy(v)
  except Exception:
    return v


class Params:
  """Named, nestable, freezable bag of hyper-parameters."""

  def __init__(self):
    object.__setattr__(self, '_immutable', False)
    object.__setattr__(self, '_slots', {})

  # ---------------------------------------------------------------- access --
  def __getattr__(self, name):
    if name in ('_slots', '_immutable'):
      return object.__getattribute__(self, name)
    try:
      return self._slots[name].value
    except KeyError:
      raise AttributeError(self._MissingKeyMsg(name)) from None

  def __setattr__(self, name, value):
    if self._immutable:
      raise TypeError('This Params instance is immutable.')
    if name in ('_slots', '_immutable'):
      object.__setattr__(self, name, value)
      return
    try:
      self._slots[name].value = value
    except KeyError:
      raise AttributeError(self._MissingKeyMsg(name)) from None

  def __dir__(self):
    return sorted(self._slots)

  def __contains__(self, name):
    return name in self._slots

  def __len__(self):
    return len(self._slots)

  def __eq__(self, other):
    if not isinstance(other, Params):
      return False
    if self._slots.keys() != other._slots.keys():
      return False
    for k, s in self._slots.items():
      a, b = s.value, other._slots[k].value
      if torch is not None and (isinstance(a, torch.Tensor) or
                                isinstance(b, torch.Tensor)):
        if a is not b:
          return False
      elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        if not np.array_equal(a, b):
          return False
      elif a != b:
        return False
    return True

  def __ne__(self, other):
    return not self == other

  def __hash__(self):
    return id(self)

  def __str__(self):
    return self._Render(0)

  def __deepcopy__(self, memo):
    return self.Copy()

  def _Render(self, depth: int) -> str:
    body = '\n'.join(
        self._slots[k].Render(depth + 1) for k in sorted(self._slots))
    return '{\n%s\n%s}' % (body, '  ' * depth)

  def _Neighbours(self, name: str) -> List[str]:
    def overlap(key):
      grams = [name[i:i + 3] for i in range(len(name) - 3)]
      if not grams:
        return 0.0
      return sum(g in key for g in grams) / len(grams)
    return [k for k in self._slots if overlap(k) > 0.5]

  def _MissingKeyMsg(self, name: str) -> str:
    near = self._Neighbours(name)
    if near:
      return '%s (did you mean: [%s])' % (name, ','.join(sorted(near)))
    return '%s (keys are %s)' % (name, sorted(self._slots))

  # ------------------------------------------------------------ definition --
  def Define(self, name: str, default_value: Any, description: str) -> None:
    """Adds a new parameter. Name must match ^[a-z_][a-z0-9_]*$."""
    if self._immutable:
      raise TypeError('This Params instance is immutable.')
    assert name is not None and isinstance(name, str) and _NAME_RE.match(name), (
        'Invalid param name: %r' % name)
    if name in self._slots:
      raise AttributeError('Parameter %s is already defined' % name)
    self._slots[name] = _Slot(name, default_value, description, default_value)

  def ParamIsSet(self, key: str) -> bool:
    """True if the (possibly nested, dotted) parameter has a non-None value; raises
    AttributeError if it does not exist (ref :478)."""
    return self.Get(key) is not None

  def MergeCommonKeysFrom(self, other: 'Params') -> 'Params':
    """Copies the value of every key both params define (ref :368)."""
    return CopyFieldsTo(other, self, ignore_unknown_keys=True)

  def Freeze(self) -> None:
    object.__setattr__(self, '_immutable', True)

  def IsImmutable(self) -> bool:
    return self._immutable

  def Copy(self):
    return self._CopyInto(type(self)())

  def _CopyInto(self, dst):
    object.__setattr__(dst, '_slots',
                       {k: s.Clone() for k, s in self._slots.items()})
    object.__setattr__(dst, '_immutable', False)
    return dst

  # -------------------------------------------------------- dotted get/set --
  def _Walk(self, dotted: str) -> Tuple['Params', str]:
    """Resolves all but the last path segment; supports `a.b[2].c`."""
    cur = self
    parts = dotted.split('.')
    for i, part in enumerate(parts[:-1]):
      m = re.match(r'^(.+)\[(.+)\]$', part)
      idx = None
      if m:
        part, idx = m.group(1), int(m.group(2))
      try:
        nxt = cur._slots[part].value
      except KeyError:
        raise AttributeError('.'.join(parts[:i + 1])) from None
      if idx is not None:
        nxt = nxt[idx]
      if not isinstance(nxt, Params):
        raise AssertionError('Cannot introspect %s for %s' %
                             (type(nxt), '.'.join(parts[:i + 1])))
      cur = nxt
    return cur, parts[-1]

  def Set(self, **kwargs):
    """Sets (possibly dotted) keys; returns self for chaining."""
    if self._immutable:
      raise TypeError('This Params instance is immutable: %s' % self)
    for name, value in kwargs.items():
      owner, key = self._Walk(name)
      if key not in owner._slots:
        raise AttributeError(self._MissingKeyMsg(name))
      # NB: like the reference, only the root's immutability is enforced so
      # text overrides can reach frozen leaf specs (e.g. params_init.scale).
      owner._slots[key].value = value
    return self

  def Get(self, name: str):
    owner, key = self._Walk(name)
    m = re.match(r'^(.+)\[(.+)\]$', key)
    try:
      if m:
        return owner._slots[m.group(1)].value[int(m.group(2))]
      return owner._slots[key].value
    except KeyError:
      raise AttributeError(self._MissingKeyMsg(name)) from None

  def Delete(self, *names):
    if self._immutable:
      raise TypeError('This Params instance is immutable.')
    for name in names:
      owner, key = self._Walk(name)
      if key not in owner._slots:
        raise AttributeError(self._MissingKeyMsg(name))
      del owner._slots[key]
    return self

  def IterParams(self) -> Iterator[Tuple[str, Any]]:
    for k, s in self._slots.items():
      yield k, s.value

  def GetKeys(self) -> List[str]:
    return sorted(self._slots)

  def GetDescription(self, name: str) -> str:
    return self._slots[name].doc

  # ----------------------------------------------------------------- visit --
  def Visit(self, visit_fn: Callable[[str, Any], None],
            enter_fn: Optional[Callable[[str, Any], bool]] = None,
            exit_fn: Optional[Callable[[str, Any], None]] = None) -> None:
    """Depth-first traversal with reference-compatible key spelling.

    Keys: `a.b` for nested Params, `a[3]` for sequence items, `a[k]` for dict
    items / dataclass & namedtuple fields / (name, Params) pair lists.
    """
    enter_fn = enter_fn or (lambda k, v: True)
    exit_fn = exit_fn or (lambda k, v: None)

    def sub(key, sub_key):
      return '%s[%s]' % (key, sub_key)

    def go(key, val):
      if isinstance(val, Params):
        if enter_fn(key, val):
          for k in sorted(val._slots):
            go((key + '.' + k) if key else k, val._slots[k].value)
          exit_fn(key, val)
      elif isinstance(val, dict):
        if enter_fn(key, val):
          for k in val:
            go(sub(key, k), val[k])
          exit_fn(key, val)
      elif dataclasses.is_dataclass(val) and not isinstance(val, type):
        if enter_fn(key, val):
          for f in dataclasses.fields(val):
            go(sub(key, f.name), getattr(val, f.name))
          exit_fn(key, val)
      elif _IsNamedTuple(val):
        if enter_fn(key, val):
          for f in val._fields:
            go(sub(key, f), getattr(val, f))
          exit_fn(key, val)
      elif isinstance(val, (list, tuple)) and val and all(
          isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], str) and
          isinstance(x[1], Params) for x in val):
        if enter_fn(key, val):
          for name, item in val:
            go(sub(key, name), item)
          exit_fn(key, val)
      elif isinstance(val, (list, range, tuple)):
        if enter_fn(key, val):
          for i, item in enumerate(val):
            go(sub(key, i), item)
          exit_fn(key, val)
      else:
        visit_fn(key, val)

    go('', self)

  # ------------------------------------------------------------ text codec --
  @staticmethod
  def _Repr(val):
    """Structure-preserving representation (scalars/str stay themselves)."""
    if isinstance(val, Params):
      return _OrderedReprDict(
          (k, Params._Repr(v)) for k, v in val.IterParams())
    if isinstance(val, dict):
      return _OrderedReprDict((k, Params._Repr(v)) for k, v in val.items())
    if isinstance(val, np.ndarray):
      return np.array2string(val, separator=', ')
    if dataclasses.is_dataclass(val) and not isinstance(val, type):
      return _OrderedReprDict(
          (k, Params._Repr(v)) for k, v in val.__dict__.items())
    if _IsNamedTuple(val):
      return _OrderedReprDict(
          (k, Params._Repr(v)) for k, v in val._asdict().items())
    if isinstance(val, (list, tuple)):
      return type(val)([Params._Repr(v) for v in val])
    if isinstance(val, (int, float, bool, str, enum.Enum)):
      return val
    if _IsDtype(val):
      return DtypeName(val)
    if isinstance(val, np.dtype) or (isinstance(val, type) and
                                     issubclass(val, np.generic)):
      return np.dtype(val).name
    if inspect.isclass(val) or inspect.isroutine(val):
      mod = inspect.getmodule(val)
      return 'type/%s/%s' % (mod.__name__ if mod else '?', val.__name__)
    return type(val).__name__

  @staticmethod
  def _Encode(val) -> str:
    """Value -> its one-leaf text representation."""
    r = Params._Repr(val)
    if isinstance(val, str):
      return _Quote(val)
    return str(r)

  def ToText(self, include_types: bool = False, separator: str = ':'):
    """Sorted `key : value` lines — the params.txt format."""
    rows: Dict[str, str] = {}
    types: Dict[str, str] = {}

    def enter(key, val):
      if isinstance(val, Params):
        return True
      if isinstance(val, (list, tuple)) and not _IsNamedTuple(val) and val:
        if all(isinstance(x, Params) for x in val):
          return True
        if all(isinstance(x, tuple) and len(x) == 2 and
               isinstance(x[0], str) and isinstance(x[1], Params) for x in val):
          return True
      if isinstance(val, dict) and val and all(
          isinstance(k, str) and isinstance(v, Params)
          for k, v in val.items()):
        return True
      if isinstance(val, (list, tuple, dict)) or _IsNamedTuple(val) or (
          dataclasses.is_dataclass(val) and not isinstance(val, type)):
        rows[key] = Params._Encode(val)
        types[key] = type(val).__name__
      return False

    def visit(key, val):
      rows[key] = Params._Encode(val)
      types[key] = type(val).__name__

    self.Visit(visit, enter_fn=enter)
    text = ''
    for k in sorted(rows):
      text += '%s %s %s\n' % (k, separator, rows[k])
    if include_types:
      return text, types
    return text

  def ToTextWithTypes(self) -> str:
    text, types = self.ToText(include_types=True)
    text += '\n\n'
    for k in sorted(types):
      text += '%s : %s\n' % (k, types[k])
    return text

  # ------------------------------------------------------------------ proto --
  # Wire-compatible with `lingvo/core/hyperparams.proto` (Hyperparam / HyperparamValue;
  # reference `ToProto` :529, `FromProto` :611) without generated code: a serialized
  # `Hyperparam` message as bytes, plus the protobuf text format for `params.pbtxt`.
  def ToProto(self) -> bytes:
    """Serialized `tensorflow.lingvo.Hyperparam` message."""
    return _HyperparamMsg(self)

  def ToProtoText(self) -> str:
    """`params.pbtxt`: the protobuf text format of `ToProto()`."""
    return _HyperparamText(self, 0)

  @classmethod
  def FromProto(cls, data: bytes) -> 'Params':
    """Rebuilds a Params tree from `ToProto()` bytes (classes are resolved by import
    path; an `InstantiableParams` comes back bound to its class)."""
    return _ParamsFromMsg(data)

  def FromTextWithTypes(self, text: str) -> None:
    body, type_block = text.split('\n\n\n')
    types = {}
    for row in type_block.split('\n'):
      if not row.strip():
        continue
      k, v = row.split(':')
      types[k.strip()] = v.strip()
    self.FromText(body, type_overrides=types)

  def FromText(self, text: str, type_overrides: Optional[Dict[str, str]] = None):
    """Parses `key : value` lines produced by ToText (or written by hand)."""
    if self._immutable:
      raise TypeError('This Params instance is immutable.')
    type_overrides = type_overrides or {}
    pending: Dict[str, str] = {}
    key = None
    quote = None
    buf: List[str] = []
    for raw in text.split('\n'):
      if key is not None:  # inside a multi-line string
        buf.append(raw)
        if _TerminalQuote(raw, quote):
          pending[key] = '\n'.join(buf)
          key, quote, buf = None, None, []
        continue
      line = raw.strip()
      if not line or line.startswith('#'):
        continue
      parts = re.split(r'\s*[:=]\s*', line, maxsplit=1)
      if len(parts) != 2:
        raise ValueError('Line {} is not in <key>:<value> format'.format(raw))
      k, v = parts[0].strip(), parts[1].strip()
      if v and v[0] in '"\'' and not (len(v) > 1 and _TerminalQuote(v[1:], v[0])):
        key, quote, buf = k, v[0], [v]
        continue
      pending[k] = v
    if key is not None:
      raise ValueError('Unterminated multi-line string for key %s' % key)

    for k, v in pending.items():
      old = self.Get(k)
      tname = type_overrides.get(k, type(old).__name__)
      self.Set(**{k: self._Decode(k, v, old, tname)})

  @staticmethod
  def _Decode(key: str, text: str, old, tname: str):
    """Text → value, typed by the current value's type (or an override)."""
    if isinstance(old, bool) or tname == 'bool':
      return bool(text) and text not in ('False', 'false')
    if (isinstance(old, int) and tname == 'int') or tname == 'int':
      return int(text)
    if isinstance(old, float) or tname == 'float':
      return float(text)
    if _IsDtype(old) or tname == 'dtype':
      return DtypeFromName(text)
    if isinstance(old, np.dtype):
      return np.dtype(text)
    if isinstance(old, enum.Enum) or '.' in text and tname not in (
        'str', 'NoneType', 'list', 'tuple', 'dict') and not isinstance(
            old, (str, type(None), list, tuple, dict)):
      cls_name, _, member = text.rpartition('.')
      if isinstance(old, enum.Enum):
        if cls_name and cls_name != type(old).__name__:
          raise ValueError('Expected enum of class %s but got %s' %
                           (type(old).__name__, cls_name))
        return type(old)[member]
    if tname in ('list', 'tuple', 'dict') or isinstance(old, (list, tuple, dict)) \
        and not isinstance(old, str):
      if text == 'NoneType':
        return None
      if tname == 'str':
        return _Unquote(text)
      val = ast.literal_eval(text)
      if tname == 'tuple' or isinstance(old, tuple):
        val = tuple(val) if isinstance(val, (list, tuple)) else val
      return val
    if text.startswith('type/'):
      _, mod, name = text.split('/', 2)
      if mod in sys.modules:
        m = sys.modules[mod]
      else:
        m = importlib.import_module(mod)
      return getattr(m, name)
    if isinstance(old, str) or tname == 'str':
      val = _Unquote(text)
      # Lists stored as strings: re-parse when the literal looks like one.
      if tname != 'str' and re.match(r'^\s*[\[\(].*[\]\)]\s*$', val):
        try:
          return ast.literal_eval(val)
        except (ValueError, SyntaxError):
          pass
      return val
    if old is None or tname == 'NoneType':
      if text in ('NoneType', 'None'):
        return None
      if text in ('True', 'False', 'true', 'false'):
        return text in ('True', 'true')
      try:
        return ast.literal_eval(text)
      except (ValueError, SyntaxError):
        if text in _TORCH_DTYPES:
          return _TORCH_DTYPES[text]
        return _Unquote(text)
    if inspect.isclass(old) or inspect.isroutine(old):
      if text == 'NoneType':
        return None
      raise ValueError('Cannot parse %s for key %s' % (text, key))
    raise ValueError('Failed to read a parameter: %s : %s (type %s)' %
                     (key, text, tname))

  def ToNestedDict(self) -> dict:
    out = {}
    for k, v in self.IterParams():
      out[k] = v.ToNestedDict() if isinstance(v, Params) else v
    return out

  def TextDiff(self, other: 'Params') -> str:
    """Human-readable diff of two Params trees (reference :1036)."""

    def diff(a, b, indent):
      text = ''
      keys = sorted(set(a.GetKeys()) | set(b.GetKeys()))
      for k in keys:
        ina, inb = k in a, k in b
        if ina and inb:
          va, vb = a.Get(k), b.Get(k)
          if isinstance(va, Params) and isinstance(vb, Params):
            sub = diff(va, vb, indent + '  ')
            if sub:
              text += '%s%s:\n%s' % (indent, k, sub)
          elif va != vb:
            text += '%s> %s: %s\n%s< %s: %s\n' % (indent, k, va, indent, k, vb)
        elif ina:
          text += '%s> %s: %s\n' % (indent, k, a.Get(k))
        else:
          text += '%s< %s: %s\n' % (indent, k, b.Get(k))
      return text

    return diff(self, other, '')


class InstantiableParams(Params):
  """Params bound to a class; `Instantiate()` constructs `cls(params, **kw)`."""

  def __init__(self, cls=None):
    super().__init__()
    self.Define('cls', cls, 'Cls that this param object is associated with.')

  def Instantiate(self, **args):
    assert self.cls is not None
    return self.cls(self, **args)

  def Copy(self):
    return self._CopyInto(type(self)(self.cls))


def CopyFieldsTo(from_p: Params, to_p: Params,
                 skip: Optional[List[str]] = None, ignore_unknown_keys: bool = False) -> Params:
  """Copies fields from `from_p` to `to_p` (reference :197). `ignore_unknown_keys`: only the
  keys both define are copied; otherwise an unknown key is an error."""
  skip = [skip] if isinstance(skip, str) else list(skip or [])
  skip.append('cls')
  for n, p in from_p.IterParams():
    if n in skip or (ignore_unknown_keys and n not in to_p):
      continue
    if isinstance(p, Params):
      to_p.Set(**{n: p.Copy()})
    else:
      to_p.Set(**{n: p})
  return to_p


def CopyFieldsSubsetTo(from_p: Params, to_p: Params, fields_to_set) -> Params:
  """Copies only the named fields (a string or list of strings) (reference :234)."""
  if not isinstance(fields_to_set, (list, tuple)):
    fields_to_set = [fields_to_set]
  for n, p in from_p.IterParams():
    if n == 'cls' or n not in fields_to_set:
      continue
    to_p.Set(**{n: p.Copy() if isinstance(p, Params) else p})
  return to_p


# -------------------------------------------------------------- hyperparams.proto ----
# HyperparamValue oneof field numbers
_PV_PARAM, _PV_LIST, _PV_TUPLE, _PV_DICT, _PV_TYPE, _PV_DTYPE, _PV_STR, _PV_BOOL, _PV_INT, \
    _PV_FLOAT, _PV_PROTO, _PV_ENUM, _PV_NTUPLE, _PV_REPR, _PV_SYMBOLIC = range(1, 16)


def _TypeName(val) -> str:
  mod = inspect.getmodule(val)
  return '%s/%s' % (mod.__name__ if mod else '?', getattr(val, '__qualname__', val.__name__))


def _ValueMsg(val) -> bytes:
  """One `HyperparamValue` (empty ⇒ None)."""
  from lingvo_b200.utils import protowire as pw  # pylint: disable=g-import-not-at-top
  if val is None:
    return b''
  if isinstance(val, Params):
    return pw.f_msg(_PV_PARAM, _HyperparamMsg(val))
  if _IsNamedTuple(val):
    body = pw.f_string(1, _TypeName(type(val)))
    for v in val:
      body += pw.f_msg(2, _ValueMsg(v))
    return pw.f_msg(_PV_NTUPLE, body)
  if isinstance(val, (list, tuple)):
    body = b''.join(pw.f_msg(1, _ValueMsg(v)) for v in val)
    return pw.f_msg(_PV_LIST if isinstance(val, list) else _PV_TUPLE, body)
  if isinstance(val, dict):
    body = b''
    for k in val:
      entry = pw.f_string(1, str(k)) + pw.f_msg(2, _ValueMsg(val[k]))
      body += pw.f_msg(1, entry)
    return pw.f_msg(_PV_DICT, body)
  if isinstance(val, bool):
    return pw.f_bool(_PV_BOOL, val)
  if isinstance(val, enum.Enum):
    return pw.f_msg(_PV_ENUM, pw.f_string(1, _TypeName(type(val))) + pw.f_string(2, val.name))
  if isinstance(val, (int, np.integer)):
    return pw.f_varint(_PV_INT, int(val))
  if isinstance(val, (float, np.floating)):
    # float_val is a 32-bit float in the schema; keep full precision when it matters
    if float(np.float32(val)) == float(val):
      return pw.f_float(_PV_FLOAT, float(val))
    return pw.f_string(_PV_REPR, repr(float(val)))
  if isinstance(val, str):
    return pw.f_string(_PV_STR, val)
  if _IsDtype(val):
    return pw.f_string(_PV_DTYPE, DtypeName(val))
  if isinstance(val, np.dtype) or (isinstance(val, type) and issubclass(val, np.generic)):
    return pw.f_string(_PV_DTYPE, 'np.' + np.dtype(val).name)
  if inspect.isclass(val) or inspect.isroutine(val):
    return pw.f_string(_PV_TYPE, _TypeName(val))
  if dataclasses.is_dataclass(val):
    body = pw.f_string(1, _TypeName(type(val)))
    for f in dataclasses.fields(val):
      body += pw.f_msg(2, _ValueMsg(getattr(val, f.name)))
    return pw.f_msg(_PV_NTUPLE, body)
  return pw.f_string(_PV_REPR, repr(val))


def _HyperparamMsg(params: 'Params') -> bytes:
  from lingvo_b200.utils import protowire as pw  # pylint: disable=g-import-not-at-top
  out = b''
  for name, val in sorted(params.IterParams()):
    out += pw.f_msg(1, pw.f_string(1, name) + pw.f_msg(2, _ValueMsg(val)))
  return out


def _ResolveType(path: str):
  mod_name, _, qual = path.partition('/')
  obj = importlib.import_module(mod_name)
  for part in qual.split('.'):
    obj = getattr(obj, part)
  return obj


def _ValueFromMsg(data: bytes):
  from lingvo_b200.utils import protowire as pw  # pylint: disable=g-import-not-at-top
  import struct  # pylint: disable=g-import-not-at-top
  fields = list(pw.parse(data))
  if not fields:
    return None
  field, _, v = fields[0]
  if field == _PV_PARAM:
    return _ParamsFromMsg(v)
  if field in (_PV_LIST, _PV_TUPLE):
    items = [_ValueFromMsg(x) for f, _, x in pw.parse(v) if f == 1]
    return items if field == _PV_LIST else tuple(items)
  if field == _PV_DICT:
    out = {}
    for f, _, entry in pw.parse(v):
      d = pw.parse_dict(entry)
      out[d[1][0].decode('utf-8')] = _ValueFromMsg(d.get(2, [b''])[0])
    return out
  if field == _PV_TYPE:
    return _ResolveType(v.decode('utf-8'))
  if field == _PV_DTYPE:
    name = v.decode('utf-8')
    if name.startswith('np.'):
      return np.dtype(name[3:])
    try:
      return DtypeFromName(name)
    except ValueError:
      return np.dtype(name)
  if field == _PV_STR:
    return v.decode('utf-8')
  if field == _PV_BOOL:
    return bool(v)
  if field == _PV_INT:
    return pw.to_signed64(v)
  if field == _PV_FLOAT:
    return float(struct.unpack('<f', v)[0])
  if field == _PV_ENUM:
    d = pw.parse_dict(v)
    return _ResolveType(d[1][0].decode('utf-8'))[d[2][0].decode('utf-8')]
  if field == _PV_NTUPLE:
    d = pw.parse_dict(v)
    cls = _ResolveType(d[1][0].decode('utf-8'))
    return cls(*[_ValueFromMsg(x) for x in d.get(2, [])])
  if field == _PV_REPR:
    text = v.decode('utf-8')
    try:
      return ast.literal_eval(text)
    except (ValueError, SyntaxError):
      return text
  raise ValueError('unsupported HyperparamValue field %d' % field)


def _ParamsFromMsg(data: bytes) -> 'Params':
  from lingvo_b200.utils import protowire as pw  # pylint: disable=g-import-not-at-top
  items = {}
  for f, _, entry in pw.parse(data):
    if f != 1:
      continue
    d = pw.parse_dict(entry)
    items[d[1][0].decode('utf-8')] = _ValueFromMsg(d.get(2, [b''])[0])
  if 'cls' in items and items['cls'] is not None:
    out = InstantiableParams(items.pop('cls'))
  else:
    items.pop('cls', None)
    out = Params()
  for k, v in items.items():
    out.Define(k, v, '')
  return out


def _TextValue(val, indent: int) -> str:
  pad = '  ' * indent
  if val is None:
    return ''
  esc = lambda t: t.replace('\\', '\\\\').replace('"', '\\"').replace('\n', '\\n')
  if isinstance(val, Params):
    return '%sparam_val {\n%s%s}\n' % (pad, _HyperparamText(val, indent + 1), pad)
  if _IsNamedTuple(val) or (dataclasses.is_dataclass(val) and not isinstance(val, type)):
    vals = list(val) if _IsNamedTuple(val) else [
        getattr(val, f.name) for f in dataclasses.fields(val)]
    body = '%s  type: "%s"\n' % (pad, _TypeName(type(val)))
    for v in vals:
      body += '%s  items {\n%s%s  }\n' % (pad, _TextValue(v, indent + 2), pad)
    return '%snamed_tuple_val {\n%s%s}\n' % (pad, body, pad)
  if isinstance(val, (list, tuple)):
    kind = 'list_val' if isinstance(val, list) else 'tuple_val'
    body = ''.join('%s  items {\n%s%s  }\n' % (pad, _TextValue(v, indent + 2), pad)
                   for v in val)
    return '%s%s {\n%s%s}\n' % (pad, kind, body, pad)
  if isinstance(val, dict):
    body = ''
    for k in val:
      body += '%s  items {\n%s    key: "%s"\n%s    value {\n%s%s    }\n%s  }\n' % (
          pad, pad, esc(str(k)), pad, _TextValue(val[k], indent + 3), pad, pad)
    return '%sdict_val {\n%s%s}\n' % (pad, body, pad)
  if isinstance(val, bool):
    return '%sbool_val: %s\n' % (pad, 'true' if val else 'false')
  if isinstance(val, enum.Enum):
    return '%senum_val {\n%s  type: "%s"\n%s  name: "%s"\n%s}\n' % (
        pad, pad, _TypeName(type(val)), pad, val.name, pad)
  if isinstance(val, (int, np.integer)):
    return '%sint_val: %d\n' % (pad, int(val))
  if isinstance(val, (float, np.floating)):
    return '%sfloat_val: %r\n' % (pad, float(val))
  if isinstance(val, str):
    return '%sstring_val: "%s"\n' % (pad, esc(val))
  if _IsDtype(val):
    return '%sdtype_val: "%s"\n' % (pad, DtypeName(val))
  if isinstance(val, np.dtype) or (isinstance(val, type) and issubclass(val, np.generic)):
    return '%sdtype_val: "np.%s"\n' % (pad, np.dtype(val).name)
  if inspect.isclass(val) or inspect.isroutine(val):
    return '%stype_val: "%s"\n' % (pad, _TypeName(val))
  return '%sstring_repr_val: "%s"\n' % (pad, esc(repr(val)))


def _HyperparamText(params: 'Params', indent: int) -> str:
  pad = '  ' * indent
  out = ''
  for name, val in sorted(params.IterParams()):
    out += '%sitems {\n%s  key: "%s"\n%s  value {\n%s%s  }\n%s}\n' % (
        pad, pad, name, pad, _TextValue(val, indent + 2), pad, pad)
  return out
