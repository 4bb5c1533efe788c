"""TF V2 "tensor bundle" checkpoint reader/writer.

The reference checkpoint layout (`<prefix>.index` + `<prefix>.data-?????-of-
?????`) must be preserved, so this module implements the format
natively:

* data shard: raw little-endian tensor bytes back to back;
* index: a leveldb-style SSTable (`tensorflow/core/lib/io/table_builder.cc`
  format: prefix-compressed blocks with restart arrays, 5-byte block trailer
  = compression byte + masked crc32c, metaindex + index blocks, 48-byte
  footer with magic 0xdb4775248b80fb57) mapping `""` → BundleHeaderProto and
  each tensor name → BundleEntryProto{dtype, shape, shard_id, offset, size,
  crc32c}.

The hot paths (crc32c, block encode) use the native extension when present.
"""

from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from lingvo_b200.utils import protowire as pw
from lingvo_b200.utils import tfrecord

_MAGIC = 0xdb4775248b80fb57
_RESTART_INTERVAL = 16
_BLOCK_SIZE = 256 * 1024

try:
  import ml_dtypes  # noqa
  _BF16 = np.dtype(ml_dtypes.bfloat16)
except Exception:  # pylint: disable=broad-except
  _BF16 = None

# tensorflow DataType enum
_DT = {
    np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3,
    np.dtype(np.uint8): 4, np.dtype(np.int16): 5, np.dtype(np.int8): 6,
    np.dtype(np.int64): 9, np.dtype(np.bool_): 10, np.dtype(np.uint16): 17,
    np.dtype(np.float16): 19, np.dtype(np.uint32): 22, np.dtype(np.uint64): 23,
}
_DT_BFLOAT16 = 14
_DT_STRING = 7
_DT_INV = {v: k for k, v in _DT.items()}


class BFloat16Array:
  """bf16 payload carried as uint16 bits (numpy has no native bf16)."""

  def __init__(self, bits: np.ndarray):
    self.bits = np.ascontiguousarray(bits, dtype=np.uint16)

  @property
  def shape(self):
    return self.bits.shape


def _Varint64(n):
  return pw.varint(n)


class _BlockBuilder:

  def __init__(self):
    self.buf = bytearray()
    self.restarts = [0]
    self.counter = 0
    self.last_key = b''

  def Add(self, key: bytes, value: bytes):
    shared = 0
    if self.counter < _RESTART_INTERVAL:
      m = min(len(self.last_key), len(key))
      while shared < m and self.last_key[shared] == key[shared]:
        shared += 1
    else:
      self.restarts.append(len(self.buf))
      self.counter = 0
    non_shared = len(key) - shared
    self.buf += pw.varint(shared) + pw.varint(non_shared) + pw.varint(len(value))
    self.buf += key[shared:] + value
    self.last_key = key
    self.counter += 1

  def Finish(self) -> bytes:
    out = bytes(self.buf)
    out += b''.join(struct.pack('<I', r) for r in self.restarts)
    out += struct.pack('<I', len(self.restarts))
    return out

  def Size(self):
    return len(self.buf) + 4 * len(self.restarts) + 4

  def Empty(self):
    return not self.buf


def _WriteBlock(f, contents: bytes) -> Tuple[int, int]:
  offset = f.tell()
  trailer_type = b'\x00'
  crc = tfrecord.masked_crc32c(contents + trailer_type)
  f.write(contents)
  f.write(trailer_type + struct.pack('<I', crc))
  return offset, len(contents)


def _Handle(offset, size) -> bytes:
  return _Varint64(offset) + _Varint64(size)


def WriteTable(path: str, items: Iterable[Tuple[bytes, bytes]]):
  """Writes sorted (key, value) pairs as an SSTable."""
  with open(path, 'wb') as f:
    index = _BlockBuilder()
    block = _BlockBuilder()
    last_key = b''
    for key, value in items:
      assert key >= last_key, 'keys must be sorted'
      block.Add(key, value)
      last_key = key
      if block.Size() >= _BLOCK_SIZE:
        off, size = _WriteBlock(f, block.Finish())
        index.Add(last_key, _Handle(off, size))
        block = _BlockBuilder()
    if not block.Empty():
      off, size = _WriteBlock(f, block.Finish())
      index.Add(last_key, _Handle(off, size))
    meta_off, meta_size = _WriteBlock(f, _BlockBuilder().Finish())
    idx_off, idx_size = _WriteBlock(f, index.Finish())
    footer = _Handle(meta_off, meta_size) + _Handle(idx_off, idx_size)
    footer += b'\x00' * (40 - len(footer))
    footer += struct.pack('<Q', _MAGIC)
    f.write(footer)


def _ParseBlock(data: bytes) -> List[Tuple[bytes, bytes]]:
  (num_restarts,) = struct.unpack('<I', data[-4:])
  limit = len(data) - 4 - 4 * num_restarts
  out = []
  pos = 0
  key = b''
  while pos < limit:
    shared, pos = pw.read_varint(data, pos)
    non_shared, pos = pw.read_varint(data, pos)
    vlen, pos = pw.read_varint(data, pos)
    key = key[:shared] + data[pos:pos + non_shared]
    pos += non_shared
    out.append((key, data[pos:pos + vlen]))
    pos += vlen
  return out


def ReadTable(path: str) -> List[Tuple[bytes, bytes]]:
  with open(path, 'rb') as f:
    data = f.read()
  if len(data) < 48:
    raise IOError('%s is too short to be an sstable' % path)
  footer = data[-48:]
  (magic,) = struct.unpack('<Q', footer[40:])
  if magic != _MAGIC:
    raise IOError('%s: bad table magic number' % path)
  pos = 0
  _, pos = pw.read_varint(footer, pos)
  _, pos = pw.read_varint(footer, pos)
  idx_off, pos = pw.read_varint(footer, pos)
  idx_size, pos = pw.read_varint(footer, pos)

  def block(off, size):
    contents = data[off:off + size]
    ctype = data[off + size]
    if ctype != 0:
      raise IOError('compressed sstable blocks are not supported')
    return contents

  out = []
  for _, handle in _ParseBlock(block(idx_off, idx_size)):
    off, p = pw.read_varint(handle, 0)
    size, p = pw.read_varint(handle, p)
    out.extend(_ParseBlock(block(off, size)))
  return out


def _ShapeProto(shape) -> bytes:
  return b''.join(pw.f_msg(2, pw.f_varint(1, int(d))) for d in shape)


def _EntryProto(dtype_enum, shape, shard_id, offset, size, crc) -> bytes:
  out = pw.f_varint(1, dtype_enum)
  out += pw.f_msg(2, _ShapeProto(shape))
  if shard_id:
    out += pw.f_varint(3, shard_id)
  if offset:
    out += pw.f_varint(4, offset)
  out += pw.f_varint(5, size)
  out += pw.f_fixed32(6, crc)
  return out


def _HeaderProto(num_shards: int) -> bytes:
  version = pw.f_varint(1, 1)  # VersionDef.producer = 1
  return pw.f_varint(1, num_shards) + pw.f_msg(3, version)


def DataPath(prefix: str, shard: int, num_shards: int) -> str:
  return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


class BundleWriter:
  """Writes `{name: ndarray}` as `<prefix>.index` + data shard(s).

  Single-process use: `BundleWriter(prefix)` → `Add…` → `Finish()`.
  Sharded use (one writer per rank, reference `saver.py:168-194` save-then-merge): rank r
  constructs `BundleWriter(prefix, shard_id=r, num_shards=W)`, adds the tensors it owns and
  calls `FinishShard()`, which returns the index entries of its shard; one process then
  calls `MergeShardIndex(prefix, [entries…], W)`.
  """

  def __init__(self, prefix: str, shard_id: int = 0, num_shards: int = 1):
    self._prefix = prefix
    self._shard, self._num_shards = int(shard_id), int(num_shards)
    os.makedirs(os.path.dirname(prefix) or '.', exist_ok=True)
    self._data_path = DataPath(prefix, self._shard, self._num_shards)
    self._tmp_data = self._data_path + '.tempstate'
    self._f = open(self._tmp_data, 'wb')
    self._entries: Dict[str, bytes] = {}
    self._offset = 0

  def Add(self, name: str, value):
    if isinstance(value, BFloat16Array):
      bits = np.ascontiguousarray(value.bits)
      raw = memoryview(bits).cast('B') if bits.size else b''
      dt, shape = _DT_BFLOAT16, value.bits.shape
    else:
      arr = np.asarray(value, order="C")
      if arr.dtype.kind in 'OUS':
        raise TypeError('string tensors are not supported: %s' % name)
      if _BF16 is not None and arr.dtype == _BF16:
        dt = _DT_BFLOAT16
      else:
        if arr.dtype not in _DT:
          raise TypeError('unsupported dtype %s for %s' % (arr.dtype, name))
        dt = _DT[arr.dtype]
      # zero-copy view of the tensor bytes (no `tobytes()` duplicate of multi-GB tensors)
      raw = memoryview(arr).cast('B') if arr.size and arr.dtype != bool else arr.tobytes()
      shape = arr.shape
    crc = tfrecord.masked_crc32c(raw)
    self._f.write(raw)
    nbytes = raw.nbytes if isinstance(raw, memoryview) else len(raw)
    self._entries[name] = _EntryProto(dt, shape, self._shard, self._offset, nbytes, crc)
    self._offset += nbytes

  def FinishShard(self) -> Dict[str, bytes]:
    """Commits this shard's data file; returns its `{name: BundleEntryProto bytes}`."""
    self._f.flush()
    os.fsync(self._f.fileno())
    self._f.close()
    os.replace(self._tmp_data, self._data_path)
    return dict(self._entries)

  def Finish(self):
    assert self._num_shards == 1, 'sharded writers use FinishShard + MergeShardIndex'
    MergeShardIndex(self._prefix, [self.FinishShard()], 1)


def MergeShardIndex(prefix: str, shard_entries, num_shards: int):
  """Writes `<prefix>.index` over the entries of all shards (atomic rename)."""
  merged: Dict[str, bytes] = {}
  for entries in shard_entries:
    for name, proto in entries.items():
      if name in merged:
        raise ValueError('tensor %s written by more than one shard' % name)
      merged[name] = proto
  items = [(b'', _HeaderProto(num_shards))]
  for name in sorted(merged, key=lambda s: s.encode('utf-8')):
    items.append((name.encode('utf-8'), merged[name]))
  tmp_index = prefix + '.index.tempstate'
  WriteTable(tmp_index, items)
  os.replace(tmp_index, prefix + '.index')


# -- dim-0 slices of one logical tensor spread over shards (expert-parallel variables) -----
_SLICE_RE = re.compile(r'^(.*)/__slice_(\d+)_(\d+)_of_(\d+)$')


def SliceKey(name: str, lo: int, hi: int, total: int) -> str:
  return '%s/__slice_%d_%d_of_%d' % (name, lo, hi, total)


def ParseSliceKey(key: str):
  m = _SLICE_RE.match(key)
  if not m:
    return None
  return m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))


class BundleReader:
  """Reads tensors of a bundle written by TF or by `BundleWriter`."""

  def __init__(self, prefix: str):
    self._prefix = prefix
    if not os.path.exists(prefix + '.index'):
      raise FileNotFoundError(prefix + '.index')
    self._entries: Dict[str, Dict[int, list]] = {}
    self._num_shards = 1
    for key, value in ReadTable(prefix + '.index'):
      if key == b'':
        hdr = pw.parse_dict(value)
        self._num_shards = hdr.get(1, [1])[0]
        continue
      self._entries[key.decode('utf-8')] = pw.parse_dict(value)
    self._files = {}

  def Keys(self) -> List[str]:
    return sorted(self._entries)

  def Has(self, name: str) -> bool:
    return name in self._entries

  def ShapeAndDtype(self, name: str):
    e = self._entries[name]
    dt = e.get(1, [1])[0]
    shape = []
    for sp in e.get(2, []):
      for dim in pw.parse_dict(sp).get(2, []):
        shape.append(pw.to_signed64(pw.parse_dict(dim).get(1, [0])[0]))
    return tuple(shape), dt

  def _File(self, shard: int):
    if shard not in self._files:
      self._files[shard] = open(DataPath(self._prefix, shard,
                                         self._num_shards), 'rb')
    return self._files[shard]

  def Read(self, name: str, check_crc: bool = False):
    e = self._entries[name]
    shape, dt = self.ShapeAndDtype(name)
    shard = e.get(3, [0])[0]
    offset = e.get(4, [0])[0]
    size = e.get(5, [0])[0]
    f = self._File(shard)
    f.seek(offset)
    raw = f.read(size)
    if check_crc and 6 in e:
      (crc,) = struct.unpack('<I', e[6][0])
      if crc != tfrecord.masked_crc32c(raw):
        raise IOError('crc mismatch for tensor %s' % name)
    if dt == _DT_BFLOAT16:
      bits = np.frombuffer(raw, dtype=np.uint16).reshape(shape)
      return BFloat16Array(bits)
    if dt == _DT_STRING:
      raise TypeError('string tensor %s is not supported' % name)
    return np.frombuffer(raw, dtype=_DT_INV[dt]).reshape(shape).copy()

  def LogicalKeys(self) -> List[str]:
    """Keys with dim-0 slice entries folded back into their logical tensor names."""
    out = set()
    for k in self._entries:
      sl = ParseSliceKey(k)
      out.add(sl[0] if sl else k)
    return sorted(out)

  def _Slices(self, name: str):
    out = []
    for k in self._entries:
      sl = ParseSliceKey(k)
      if sl and sl[0] == name:
        out.append((sl[1], sl[2], sl[3], k))
    return sorted(out)

  def HasLogical(self, name: str) -> bool:
    return name in self._entries or bool(self._Slices(name))

  def LogicalShape(self, name: str):
    if name in self._entries:
      return self.ShapeAndDtype(name)[0]
    sl = self._Slices(name)
    shape = self.ShapeAndDtype(sl[0][3])[0]
    return (sl[0][2],) + tuple(shape[1:])

  def ReadRange(self, name: str, lo: Optional[int] = None, hi: Optional[int] = None):
    """Rows [lo, hi) of dim 0 of logical tensor `name` (whole tensor by default), whether
    it was saved whole or as per-rank dim-0 slices."""
    if name in self._entries:
      full = self.Read(name)
      if lo is None:
        return full
      if isinstance(full, BFloat16Array):
        return BFloat16Array(full.bits[lo:hi])
      return full[lo:hi]
    slices = self._Slices(name)
    if not slices:
      raise KeyError(name)
    total = slices[0][2]
    lo = 0 if lo is None else lo
    hi = total if hi is None else hi
    parts, cursor = [], lo
    for s_lo, s_hi, _, key in slices:
      if s_hi <= cursor or s_lo >= hi:
        continue
      if s_lo > cursor:
        raise IOError('slices of %s do not cover row %d' % (name, cursor))
      part = self.Read(key)
      bits = part.bits if isinstance(part, BFloat16Array) else part
      parts.append(bits[cursor - s_lo:min(hi, s_hi) - s_lo])
      cursor = min(hi, s_hi)
    if cursor < hi:
      raise IOError('slices of %s do not cover rows up to %d' % (name, hi))
    cat = np.concatenate(parts, axis=0) if len(parts) > 1 else parts[0]
    if isinstance(part, BFloat16Array):
      return BFloat16Array(np.ascontiguousarray(cat))
    return np.ascontiguousarray(cat)

  def ReadAll(self) -> Dict[str, np.ndarray]:
    return {k: self.Read(k) for k in self.Keys()}

  def Close(self):
    for f in self._files.values():
      f.close()
    self._files = {}
