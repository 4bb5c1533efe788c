"""TFRecord framing (length / masked-crc32c) reader & writer.

Format: uint64 length | uint32 masked_crc(length) | data | uint32
masked_crc(data). Used by the event-file writer, tensor-bundle checkpoints
and the record yielder (`tfrecord:` file type, reference
`record_yielder.cc:275-381`). A C++ implementation (`ops/csrc/native_io.cpp`)
is used when the extension is built; this module is the portable fallback.
"""

import struct
from typing import Iterator

_TABLE = []


def _MakeTable():
  poly = 0x82F63B78
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ poly if c & 1 else c >> 1
    _TABLE.append(c)


_MakeTable()


_NATIVE_CRC = None


def _NativeCrc():
  """`_H.crc32c_buffer` (hardware CRC32C, zero-copy, GIL released) or False."""
  global _NATIVE_CRC
  if _NATIVE_CRC is None:
    try:
      from lingvo_b200 import ops  # pylint: disable=g-import-not-at-top
      _NATIVE_CRC = getattr(ops.host(), 'crc32c_buffer', False)
    except Exception:  # pylint: disable=broad-except
      _NATIVE_CRC = False
  return _NATIVE_CRC


def crc32c(data, crc: int = 0) -> int:
  """CRC32C of a bytes-like / buffer object (numpy arrays included)."""
  fn = _NativeCrc()
  if fn:
    return fn(data, crc)
  if not isinstance(data, (bytes, bytearray)):
    data = bytes(memoryview(data).cast('B'))
  c = crc ^ 0xFFFFFFFF
  tbl = _TABLE
  for b in data:
    c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
  c = crc32c(data)
  return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc32c(masked: int) -> int:
  rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
  return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


class TFRecordWriter:

  def __init__(self, path: str, mode: str = 'wb'):
    self._f = open(path, mode)

  def write(self, data: bytes) -> None:
    hdr = struct.pack('<Q', len(data))
    self._f.write(hdr)
    self._f.write(struct.pack('<I', masked_crc32c(hdr)))
    self._f.write(data)
    self._f.write(struct.pack('<I', masked_crc32c(data)))

  def flush(self):
    self._f.flush()

  def close(self):
    self._f.close()

  def __enter__(self):
    return self

  def __exit__(self, *a):
    self.close()


def ReadRecords(path: str, check_crc: bool = False) -> Iterator[bytes]:
  opener = open
  if path.endswith('.gz'):
    import gzip
    opener = gzip.open
  with opener(path, 'rb') as f:
    while True:
      hdr = f.read(8)
      if len(hdr) < 8:
        return
      (n,) = struct.unpack('<Q', hdr)
      (hcrc,) = struct.unpack('<I', f.read(4))
      if check_crc and hcrc != masked_crc32c(hdr):
        raise IOError('corrupted record header in %s' % path)
      data = f.read(n)
      if len(data) < n:
        raise IOError('truncated record in %s' % path)
      (dcrc,) = struct.unpack('<I', f.read(4))
      if check_crc and dcrc != masked_crc32c(data):
        raise IOError('corrupted record in %s' % path)
      yield data
