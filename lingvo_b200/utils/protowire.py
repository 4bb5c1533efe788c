"""Minimal protobuf wire-format encoder/decoder (no generated code).

Enough to write TensorBoard `Event`/`Summary` protos, tensor-bundle
`BundleHeaderProto`/`BundleEntryProto`, tf.Example records and the
hyperparams/inference-graph protos.
"""

import struct
from typing import Dict, Iterator, List, Tuple, Union


def varint(n: int) -> bytes:
  if n < 0:
    n += 1 << 64
  out = bytearray()
  while True:
    b = n & 0x7F
    n >>= 7
    if n:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def key(field: int, wire: int) -> bytes:
  return varint((field << 3) | wire)


def f_varint(field: int, v: int) -> bytes:
  return key(field, 0) + varint(int(v))


def f_bool(field: int, v: bool) -> bytes:
  return key(field, 0) + varint(1 if v else 0)


def f_fixed64(field: int, v: int) -> bytes:
  return key(field, 1) + struct.pack('<Q', v)


def f_double(field: int, v: float) -> bytes:
  return key(field, 1) + struct.pack('<d', v)


def f_fixed32(field: int, v: int) -> bytes:
  return key(field, 5) + struct.pack('<I', v)


def f_float(field: int, v: float) -> bytes:
  return key(field, 5) + struct.pack('<f', v)


def f_bytes(field: int, v: Union[bytes, str]) -> bytes:
  if isinstance(v, str):
    v = v.encode('utf-8')
  return key(field, 2) + varint(len(v)) + v


f_msg = f_bytes
f_string = f_bytes


def f_packed_varint(field: int, vals) -> bytes:
  body = b''.join(varint(int(v)) for v in vals)
  return key(field, 2) + varint(len(body)) + body


def f_packed_float(field: int, vals) -> bytes:
  body = struct.pack('<%df' % len(vals), *vals)
  return key(field, 2) + varint(len(body)) + body


def read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
  result = 0
  shift = 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def parse(buf: bytes) -> Iterator[Tuple[int, int, Union[int, bytes]]]:
  """Yields (field, wire_type, value) triples of one message."""
  pos = 0
  n = len(buf)
  while pos < n:
    k, pos = read_varint(buf, pos)
    field, wire = k >> 3, k & 7
    if wire == 0:
      v, pos = read_varint(buf, pos)
    elif wire == 1:
      v = buf[pos:pos + 8]
      pos += 8
    elif wire == 2:
      ln, pos = read_varint(buf, pos)
      v = buf[pos:pos + ln]
      pos += ln
    elif wire == 5:
      v = buf[pos:pos + 4]
      pos += 4
    else:
      raise ValueError('unsupported wire type %d' % wire)
    yield field, wire, v


def parse_dict(buf: bytes) -> Dict[int, List]:
  out: Dict[int, List] = {}
  for f, _, v in parse(buf):
    out.setdefault(f, []).append(v)
  return out


def to_signed64(v: int) -> int:
  return v - (1 << 64) if v >= (1 << 63) else v


def as_double(v) -> float:
  """Value of a fixed64 field (8 raw bytes from `parse`) as a double; numbers pass through."""
  if isinstance(v, (bytes, bytearray)):
    return struct.unpack('<d', v)[0] if len(v) == 8 else struct.unpack('<f', v)[0]
  return float(v)


def as_float(v) -> float:
  if isinstance(v, (bytes, bytearray)):
    return struct.unpack('<f', v)[0] if len(v) == 4 else struct.unpack('<d', v)[0]
  return float(v)


def parse_packed_varints(buf: bytes) -> List[int]:
  out, pos = [], 0
  while pos < len(buf):
    v, pos = read_varint(buf, pos)
    out.append(v)
  return out


def f_packed_double(field: int, vals) -> bytes:
  payload = b''.join(struct.pack('<d', float(v)) for v in vals)
  return f_bytes(field, payload)
