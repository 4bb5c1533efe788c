"""TensorBoard event-file writer/reader (`events.out.tfevents.*`).

Reference: `tf.summary.FileWriter` use in `base_runner.py:653` /
`summary_utils.py`. Event{wall_time=1,step=2,file_version=3,summary=5};
Summary.Value{tag=1, simple_value=2, image=4, histo=5, tensor=8(text)}.
"""

import os
import socket
import threading
import time
from typing import Dict, Iterator, List, Tuple

import numpy as np

from lingvo_b200.utils import protowire as pw
from lingvo_b200.utils import tfrecord


def _Event(step=None, summary=None, file_version=None, wall_time=None):
  out = pw.f_double(1, wall_time if wall_time is not None else time.time())
  if step is not None:
    out += pw.f_varint(2, int(step))
  if file_version is not None:
    out += pw.f_string(3, file_version)
  if summary is not None:
    out += pw.f_msg(5, summary)
  return out


def ScalarValue(tag: str, value: float) -> bytes:
  return pw.f_msg(1, pw.f_string(1, tag) + pw.f_float(2, float(value)))


def HistogramValue(tag: str, values) -> bytes:
  v = np.asarray(values, dtype=np.float64).reshape(-1)
  if v.size == 0:
    v = np.zeros([1])
  counts, edges = np.histogram(v, bins=30)
  h = (pw.f_double(1, float(v.min())) + pw.f_double(2, float(v.max())) +
       pw.f_double(3, float(v.size)) + pw.f_double(4, float(v.sum())) +
       pw.f_double(5, float((v * v).sum())))
  lim = b''.join(__import__('struct').pack('<d', e) for e in edges[1:])
  bkt = b''.join(__import__('struct').pack('<d', float(c)) for c in counts)
  h += pw.key(6, 2) + pw.varint(len(lim)) + lim
  h += pw.key(7, 2) + pw.varint(len(bkt)) + bkt
  return pw.f_msg(1, pw.f_string(1, tag) + pw.f_msg(5, h))


def TextValue(tag: str, text: str) -> bytes:
  # TensorProto{dtype=1(DT_STRING=7), string_val=8}; metadata plugin 'text'.
  tensor = pw.f_varint(1, 7) + pw.f_bytes(8, text.encode('utf-8'))
  plugin = pw.f_msg(1, pw.f_string(1, 'text'))
  return pw.f_msg(1, pw.f_string(1, tag) + pw.f_msg(8, tensor) +
                  pw.f_msg(9, plugin))


def EncodePng(image) -> bytes:
  """Minimal PNG encoder (zlib only): uint8 `[H, W]` (gray), `[H, W, 3]` (RGB) or
  `[H, W, 4]` (RGBA); float inputs in [0, 1] are scaled to 0..255."""
  import struct
  import zlib
  a = np.asarray(image)
  if a.dtype != np.uint8:
    a = (np.clip(a.astype(np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
  if a.ndim == 3 and a.shape[2] == 1:
    a = a[:, :, 0]
  assert a.ndim == 2 or (a.ndim == 3 and a.shape[2] in (3, 4)), a.shape
  color = {2: 0, 3: 2, 4: 6}[2 if a.ndim == 2 else a.shape[2]]
  h, w = a.shape[:2]
  rows = a.reshape(h, -1)
  raw = np.concatenate([np.zeros((h, 1), np.uint8), rows], 1).tobytes()   # filter 0 per row

  def Chunk(kind, data):
    body = kind + data
    return struct.pack('>I', len(data)) + body + struct.pack('>I', zlib.crc32(body) & 0xffffffff)

  return (b'\x89PNG\r\n\x1a\n' + Chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, color, 0, 0, 0))
          + Chunk(b'IDAT', zlib.compress(raw, 6)) + Chunk(b'IEND', b''))


def DecodePng(data: bytes):
  """Inverse of `EncodePng` for the PNGs this module writes (8-bit, filter 0)."""
  import struct
  import zlib
  assert data[:8] == b'\x89PNG\r\n\x1a\n'
  pos, idat, hdr = 8, b'', None
  while pos < len(data):
    n, kind = struct.unpack('>I4s', data[pos:pos + 8])
    body = data[pos + 8:pos + 8 + n]
    if kind == b'IHDR':
      hdr = struct.unpack('>IIBBBBB', body)
    elif kind == b'IDAT':
      idat += body
    pos += 12 + n
  w, h, depth, color = hdr[:4]
  assert depth == 8
  ch = {0: 1, 2: 3, 6: 4}[color]
  raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
  assert not raw[:, 0].any(), 'only filter type 0 is supported'
  out = raw[:, 1:].reshape(h, w, ch)
  return out[:, :, 0] if ch == 1 else out


def ImageValue(tag: str, image) -> bytes:
  """Summary.Value{tag=1, image=4{height=1, width=2, colorspace=3, encoded_image_string=4}};
  `image` is an array (encoded here) or ready-made PNG bytes."""
  if isinstance(image, (bytes, bytearray)):
    png = bytes(image)
    import struct
    w, h = struct.unpack('>II', png[16:24])
    colorspace = {0: 1, 2: 3, 6: 4}.get(png[25], 3)
  else:
    a = np.asarray(image)
    h, w = a.shape[:2]
    colorspace = 1 if a.ndim == 2 or a.shape[2] == 1 else a.shape[2]
    png = EncodePng(a)
  img = (pw.f_varint(1, int(h)) + pw.f_varint(2, int(w)) + pw.f_varint(3, int(colorspace)) +
         pw.f_bytes(4, png))
  return pw.f_msg(1, pw.f_string(1, tag) + pw.f_msg(4, img))


class EventFileWriter:
  """Thread-safe append-only event writer."""

  def __init__(self, logdir: str, filename_suffix: str = ''):
    os.makedirs(logdir, exist_ok=True)
    fname = 'events.out.tfevents.%010d.%s%s' % (int(time.time()),
                                               socket.gethostname(),
                                               filename_suffix)
    self._path = os.path.join(logdir, fname)
    self._w = tfrecord.TFRecordWriter(self._path)
    self._lock = threading.Lock()
    self._w.write(_Event(file_version='brain.Event:2', step=0))
    self._w.flush()

  @property
  def path(self):
    return self._path

  def add_scalars(self, scalars: Dict[str, float], step: int):
    body = b''.join(ScalarValue(k, v) for k, v in scalars.items())
    self.add_summary_bytes(body, step)

  def add_scalar(self, tag: str, value: float, step: int):
    self.add_summary_bytes(ScalarValue(tag, value), step)

  def add_histogram(self, tag: str, values, step: int):
    self.add_summary_bytes(HistogramValue(tag, values), step)

  def add_text(self, tag: str, text: str, step: int):
    self.add_summary_bytes(TextValue(tag, text), step)

  def add_image(self, tag: str, image, step: int):
    self.add_summary_bytes(ImageValue(tag, image), step)

  def add_summary_bytes(self, summary: bytes, step: int):
    with self._lock:
      self._w.write(_Event(step=step, summary=summary))

  def flush(self):
    with self._lock:
      self._w.flush()

  def close(self):
    with self._lock:
      self._w.close()


def ReadScalars(path: str) -> Iterator[Tuple[int, str, float]]:
  """Yields (step, tag, simple_value) from an event file."""
  import struct
  for rec in tfrecord.ReadRecords(path):
    ev = pw.parse_dict(rec)
    step = pw.to_signed64(ev.get(2, [0])[0]) if 2 in ev else 0
    for s in ev.get(5, []):
      for val in pw.parse_dict(s).get(1, []):
        d = pw.parse_dict(val)
        if 1 in d and 2 in d:
          yield step, d[1][0].decode('utf-8'), struct.unpack('<f', d[2][0])[0]


def ReadImages(path: str) -> Iterator[Tuple[int, str, np.ndarray]]:
  """Yields (step, tag, decoded image) for every image summary of an event file."""
  for rec in tfrecord.ReadRecords(path):
    ev = pw.parse_dict(rec)
    step = pw.to_signed64(ev.get(2, [0])[0]) if 2 in ev else 0
    for s in ev.get(5, []):
      for val in pw.parse_dict(s).get(1, []):
        d = pw.parse_dict(val)
        if 1 in d and 4 in d:
          img = pw.parse_dict(d[4][0])
          yield step, d[1][0].decode('utf-8'), DecodePng(img[4][0])
