"""TensorBoard event-file writer/reader (`events.out.tfevents.*`).

Reference: `tf.summary.FileWriter` use in `base_runner.py:653` /
`summary_utils.py`. Event{wall_time=1,step=2,file_version=3,summary=5};
Summary.Value{tag=1, simple_value=2, histo=5, tensor=8(text)}.
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
