"""Decoder-job output helpers (ref `lingvo/core/decoder_lib.py`)."""
import pickle

import numpy as np
import torch

from lingvo_b200.utils import protowire as pw


def WriteKeyValuePairs(filename, key_value_pairs):
  with open(filename, 'wb') as f:
    pickle.dump(key_value_pairs, f, protocol=pickle.HIGHEST_PROTOCOL)


def SerializeOutputs(nmap) -> bytes:
  """Record{fields: map<string, bytes(npy)>} of the flattened NestedMap."""
  import io
  out = b''
  for key, value in nmap.FlattenItems():
    arr = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
    buf = io.BytesIO()
    np.save(buf, arr, allow_pickle=False)
    out += pw.f_bytes(1, pw.f_bytes(1, key) + pw.f_bytes(2, buf.getvalue()))
  return out


def DeserializeOutputs(serialized: bytes):
  import io
  out = {}
  for entry in pw.parse_dict(serialized).get(1, []):
    kv = pw.parse_dict(entry)
    out[kv[1][0].decode()] = np.load(io.BytesIO(kv[2][0]), allow_pickle=False)
  return out
