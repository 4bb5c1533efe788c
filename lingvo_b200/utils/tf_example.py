"""Minimal `tf.train.Example` codec (wire-compatible, no TensorFlow).

Example{1: Features{1: map<string, Feature>}}; Feature is a oneof of
BytesList(1){1: bytes*}, FloatList(2){1: packed float}, Int64List(3){1: packed varint}.
"""

from __future__ import annotations

import struct
from typing import Dict

import numpy as np

from lingvo_b200.utils import protowire as pw


def ParseExample(record: bytes) -> Dict[str, np.ndarray]:
  out = {}
  ex = pw.parse_dict(record)
  for feats in ex.get(1, []):
    for entry in pw.parse_dict(feats).get(1, []):
      kv = pw.parse_dict(entry)
      name = kv[1][0].decode('utf-8')
      feature = pw.parse_dict(kv[2][0]) if 2 in kv else {}
      if 3 in feature:                                    # Int64List
        vals = []
        for field, wire, v in pw.parse(feature[3][0]):
          if wire == 2:
            pos = 0
            while pos < len(v):
              x, pos = pw.read_varint(v, pos)
              vals.append(pw.to_signed64(x))
          else:
            vals.append(pw.to_signed64(v))
        out[name] = np.asarray(vals, np.int64)
      elif 2 in feature:                                  # FloatList
        vals = []
        for field, wire, v in pw.parse(feature[2][0]):
          if wire == 2:
            vals.extend(struct.unpack('<%df' % (len(v) // 4), v))
          else:
            vals.append(struct.unpack('<f', struct.pack('<I', v))[0])
        out[name] = np.asarray(vals, np.float32)
      elif 1 in feature:                                  # BytesList
        out[name] = np.asarray(
            [v for _, _, v in pw.parse(feature[1][0])], dtype=object)
      else:
        out[name] = np.asarray([], np.float32)
  return out


def MakeExample(features: Dict[str, object]) -> bytes:
  entries = b''
  for name, val in features.items():
    arr = np.asarray(val)
    if arr.dtype.kind in 'iu':
      feat = pw.f_bytes(3, pw.f_packed_varint(1, [int(v) & ((1 << 64) - 1) for v in arr.reshape(-1)]))
    elif arr.dtype.kind == 'f':
      feat = pw.f_bytes(2, pw.f_packed_float(1, [float(v) for v in arr.reshape(-1)]))
    else:
      items = val if isinstance(val, (list, tuple)) else [val]
      feat = pw.f_bytes(1, b''.join(pw.f_bytes(1, v) for v in items))
    entries += pw.f_bytes(1, pw.f_bytes(1, name) + pw.f_bytes(2, feat))
  return pw.f_bytes(1, entries)
