"""Summaries, step-rate tracking and model analysis.

Reference `lingvo/core/summary_utils.py`: `scalar/histogram/image/text` gated
by `cluster.add_summary` (:38-95), `StepRateTracker` (:393-429),
`ModelAnalysis` (:432-510). Here summaries are collected into a thread-local
`SummaryCollector` (host scalars) that runners flush to an `EventFileWriter`.
"""

from __future__ import annotations

import threading
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from lingvo_b200.core import cluster_factory


class _Collectors(threading.local):

  def __init__(self):
    super().__init__()
    self.stack: List['SummaryCollector'] = []


_COLLECTORS = _Collectors()


class SummaryCollector:
  """Context that captures summaries emitted by layers during FProp/BProp."""

  def __init__(self):
    self.scalars: Dict[str, Any] = {}
    self.histograms: Dict[str, Any] = {}
    self.texts: Dict[str, str] = {}
    self.images: Dict[str, Any] = {}

  def __enter__(self):
    _COLLECTORS.stack.append(self)
    return self

  def __exit__(self, *args):
    _COLLECTORS.stack.pop()

  def Resolve(self) -> Dict[str, float]:
    """Host-side float values (one sync for all device scalars)."""
    out = {}
    dev = {k: v for k, v in self.scalars.items() if isinstance(v, torch.Tensor)}
    if dev:
      stacked = torch.stack([v.detach().float().reshape(()) for v in dev.values()])
      for k, v in zip(dev.keys(), stacked.cpu().tolist()):
        out[k] = v
    for k, v in self.scalars.items():
      if k not in out:
        out[k] = float(v)
    return out

  def WriteTo(self, writer, step: int):
    vals = self.Resolve()
    if vals:
      writer.add_scalars(vals, step)
    for k, v in self.histograms.items():
      arr = v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else v
      writer.add_histogram(k, arr, step)
    for k, v in self.texts.items():
      writer.add_text(k, v, step)
    for k, v in self.images.items():
      for i, img in enumerate(_ToImages(v)):
        writer.add_image('%s/image/%d' % (k, i) if len(_ToImages(v)) > 1 else k + '/image',
                         img, step)


def _ToImages(v, max_outputs=3):
  """PNG bytes pass through; arrays / tensors `[H, W]`, `[H, W, C]` or `[B, H, W, C]` with
  values in [0, 1] become a list of images."""
  if isinstance(v, (bytes, bytearray)):
    return [bytes(v)]
  if isinstance(v, (list, tuple)):
    return [x for e in v for x in _ToImages(e)][:max_outputs]
  a = v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
  if a.ndim == 4:
    return [a[i] for i in range(min(a.shape[0], max_outputs))]
  return [a]


def _Current() -> Optional[SummaryCollector]:
  if not _COLLECTORS.stack:
    return None
  if not cluster_factory.Current().add_summary:
    return None
  return _COLLECTORS.stack[-1]


def _ShouldAddSummary() -> bool:
  return _Current() is not None


def scalar(name: str, value, **kwargs):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.scalars[name] = value


def scalar_input_stats(*args, **kwargs):  # pylint: disable=invalid-name
  scalar(*args, **kwargs)


def histogram(name: str, tensor):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.histograms[name] = tensor


def text(name: str, value: str):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.texts[name] = value


def image(name: str, tensor, **kwargs):  # pylint: disable=invalid-name
  c = _Current()
  if c is not None:
    c.images[name] = tensor


# the *_v2 entry points of the reference (:69-95) share the collector here
scalar_v2, histogram_v2, image_v2, text_v2 = scalar, histogram, image, text   # pylint: disable=invalid-name


def GetTensorName(tensor, name_eager=None, i_eager=None):
  """Tensors are anonymous in eager execution: `<name>_<i>` (ref :513)."""
  del tensor
  return '%s_%d' % (name_eager, i_eager) if name_eager is not None else 'tensor'


def SequenceLength(padding):
  """Non-padded length `[batch]` (int) of `[batch, seqlen]` 0/1 paddings (ref :97)."""
  return torch.round((1.0 - padding.float()).sum(1)).to(torch.int32).reshape(padding.shape[0])


def TrimPaddingAndPlotSequence(fig, axes, seq_matrix, seq_len, **kwargs):
  """Plot function: image of `seq_matrix[:, :seq_len]` (`(dim, time)`) (ref :114)."""
  from lingvo_b200.core import plot  # pylint: disable=g-import-not-at-top
  plot.AddImage(fig, axes, np.asarray(seq_matrix)[:, :int(seq_len)], **kwargs)


def TrimPaddingAndPlotAttention(fig, axes, atten_matrix, src_len, tgt_len, transcript=None,
                                **kwargs):
  """Plot function: `atten_matrix[:tgt_len, :src_len]` with a fixed 0..1 colour scale and the
  transcript under the source axis (ref :129)."""
  from lingvo_b200.core import plot  # pylint: disable=g-import-not-at-top
  plot.AddImage(fig, axes, np.asarray(atten_matrix)[:int(tgt_len), :int(src_len)],
                clim=(0, 1), **kwargs)
  if transcript is not None:
    if isinstance(transcript, np.ndarray):
      transcript = ' '.join(str(t) for t in transcript[:int(src_len)])
    axes.set_xlabel(plot.ToUnicode(transcript), size='x-small', wrap=True)


def _HeatmapPng(matrix, scale=4):
  """Dependency-free rendering of a `[rows, cols]` matrix with values in [0, 1]: dark =
  high (the `bone_r` look), every cell `scale`×`scale` pixels."""
  from lingvo_b200.utils import tfevents  # pylint: disable=g-import-not-at-top
  m = 1.0 - np.clip(np.asarray(matrix, np.float64), 0.0, 1.0)
  return tfevents.EncodePng(np.kron(m, np.ones((scale, scale))))


def AddAttentionSummaryBatchMajor(name, attention_tensors, src_paddings, tgt_paddings,
                                  transcripts=None, max_outputs=3):
  """Image summaries of attention matrices `[batch, target_len, source_len]`, trimmed to the
  non-padded lengths, plus the `average_normalized_entropy` scalar that drops as attention
  sharpens (ref :195). `src_paddings` / `tgt_paddings`: one tensor or one per attention
  tensor. With matplotlib the figure has titles / axes / transcripts; without, each matrix
  is rendered as a plain heat-map PNG."""
  def AsList(x):
    x = x if isinstance(x, list) else [x]
    if len(x) not in (1, len(attention_tensors)):
      raise ValueError('Bad length of paddings list {}'.format(len(x)))
    return x

  src_paddings, tgt_paddings = AsList(src_paddings), AsList(tgt_paddings)
  pick = lambda xs, i: xs[0 if len(xs) == 1 else i]
  for i, atten in enumerate(attention_tensors):
    want = list(pick(tgt_paddings, i).shape[:2]) + [pick(src_paddings, i).shape[1]]
    assert list(atten.shape[:3]) == want, (list(atten.shape), want)
  c = _Current()
  if c is None:
    return
  from lingvo_b200.core import plot  # pylint: disable=g-import-not-at-top
  src_lens = [SequenceLength(p) for p in src_paddings]
  tgt_lens = [SequenceLength(p) for p in tgt_paddings]
  fig = plot.MatplotlibFigureSummary(name + '/Attention', max_outputs=max_outputs,
                                     gridspec_kwargs={'hspace': 0.3})
  fallback = []
  for n, atten in enumerate(attention_tensors):
    a = atten.detach().float()
    max_entropy = torch.log(pick(src_lens, n).float()).reshape(-1, 1, 1)
    entropy = -a * torch.log(a + 1e-10) / max_entropy
    scalar(name + '/Attention/average_normalized_entropy/%d' % n, entropy.mean())
    args = [a.cpu().numpy(), pick(src_lens, n).cpu().numpy(), pick(tgt_lens, n).cpu().numpy()]
    if transcripts is not None and n == 0:
      args.append(np.asarray(transcripts))
    fig.AddSubplot(args, TrimPaddingAndPlotAttention, title=GetTensorName(atten, name, n),
                   xlabel='Input', ylabel='Output')
    fallback.append(args)
  pngs = fig.Finalize()
  if pngs is None:
    pngs = []
    for b in range(min(max_outputs, fallback[0][0].shape[0])):
      rows = [m[b][:int(tl[b]), :int(sl[b])] for m, sl, tl, *_ in fallback]
      width = max(r.shape[1] for r in rows)
      rows = [np.pad(r, ((0, 1), (0, width - r.shape[1]))) for r in rows]   # 1 blank row between
      pngs.append(_HeatmapPng(np.concatenate(rows, 0)[:-1]))
  c.images[name + '/Attention'] = list(pngs)


def AddAttentionSummary(name, attention_tensors, src_paddings, tgt_paddings, transcripts=None,
                        max_outputs=3):
  """Time-major twin: attention `[target_len, batch, source_len]`, paddings `[len, batch]`."""
  tr = lambda ps: [p.t() for p in (ps if isinstance(ps, list) else [ps])]
  AddAttentionSummaryBatchMajor(name, [a.transpose(0, 1) for a in attention_tensors],
                                tr(src_paddings), tr(tgt_paddings), transcripts, max_outputs)


def PrepareSequenceForPlot(tensor, padding, name):
  """`[batch, time, …]` → (`[batch, dim, time]` with trailing dims flattened, lengths)."""
  del name
  b, t = tensor.shape[:2]
  return tensor.reshape(b, t, -1).transpose(1, 2), SequenceLength(padding)


def PlotSequenceFeatures(plots, name, **kwargs):
  """Stack of per-example feature images, one row per (tensor, seq_len) pair (ref :347)."""
  c = _Current()
  if c is None:
    return
  from lingvo_b200.core import plot  # pylint: disable=g-import-not-at-top
  fig = plot.MatplotlibFigureSummary(name, figsize=(8, len(plots) * 3.5))
  for i, (tensor, seq_len) in enumerate(plots):
    fig.AddSubplot([tensor.detach().float().cpu().numpy(), seq_len.cpu().numpy()],
                   TrimPaddingAndPlotSequence, title=GetTensorName(tensor, name, i), **kwargs)
  pngs = fig.Finalize()
  if pngs is None:
    pngs = []
    for b in range(min(3, plots[0][0].shape[0])):
      rows = []
      for tensor, seq_len in plots:
        m = tensor[b, :, :int(seq_len[b])].detach().float().cpu().numpy()
        lo, hi = float(m.min()), float(m.max())
        rows.append((m - lo) / (hi - lo) if hi > lo else np.zeros_like(m))
      width = max(r.shape[1] for r in rows)
      pngs.append(_HeatmapPng(np.concatenate(
          [np.pad(r, ((0, 0), (0, width - r.shape[1]))) for r in rows], 0)))
  c.images[name] = list(pngs)


class StatsCounter:
  """A named monotone counter kept as a (non-trainable, checkpointed) int64 variable whose
  pre-increment value is reported as a scalar summary (ref :367)."""

  def __init__(self, name):
    from lingvo_b200.core import py_utils  # pylint: disable=g-import-not-at-top
    self._name = name
    self._var = py_utils.CreateVariable(
        name, py_utils.WeightParams([], py_utils.WeightInit.Constant(0), torch.int64),
        trainable=False)

  @property
  def var(self):
    return self._var

  def Value(self):
    return self._var.detach().clone()

  def IncBy(self, delta):
    """Adds `delta`; returns the new value (a 0-d int64 tensor)."""
    scalar(self._name, self._var.detach().clone())
    with torch.no_grad():
      self._var.add_(torch.as_tensor(delta, device=self._var.device).to(torch.int64))
    return self._var.detach().clone()


def AddNormSummary(name: str, vs_gs) -> Tuple[torch.Tensor, torch.Tensor]:
  """Returns (var_norm, grad_norm) and emits both as scalars."""
  from lingvo_b200.core import py_utils
  leaves = [vg for vg in vs_gs.Flatten() if isinstance(vg, py_utils.VarGrad)]
  vn = torch.sqrt(py_utils.SumSquared([vg.var.detach() for vg in leaves]))
  gn = torch.sqrt(py_utils.SumSquared([vg.grad for vg in leaves]))
  scalar('var_norm/' + name, vn)
  scalar('grad_norm/' + name, gn)
  return vn, gn


def CollectVarHistogram(vs_gs):
  from lingvo_b200.core import py_utils
  for vg in vs_gs.Flatten():
    if isinstance(vg, py_utils.VarGrad):
      n = getattr(vg.var, 'var_name', 'var')
      histogram('var_hist/' + n, vg.var)
      histogram('grad_hist/' + n, vg.grad)


class StepRateTracker:
  """steps/sec & examples/sec over a sliding window (reference :393-429)."""

  def __init__(self):
    self._first_step = -1
    self._time_steps: List[Tuple[float, int, float]] = []

  def ComputeStepRate(self, current_steps: int, total_examples: float):
    if self._time_steps:
      total_examples += self._time_steps[-1][-1]
    else:
      self._first_step = current_steps
    self._time_steps.append((time.time(), current_steps, total_examples))
    # Keep a window of ~1000 steps.
    t_n, s_n, e_n = self._time_steps[-1]
    i = 0
    while i + 1 < len(self._time_steps) and s_n - self._time_steps[i + 1][1] > 1000:
      i += 1
    self._time_steps = self._time_steps[i:]
    (t0, s0, e0), (t1, s1, e1) = self._time_steps[0], self._time_steps[-1]
    rate = example_rate = 0.0
    if t1 > t0 + 1e-9:
      elapsed = t1 - t0
      rate = (s1 - s0) / elapsed
      example_rate = (e1 - e0) / elapsed
    return rate, example_rate, total_examples


def ModelAnalysis(model) -> Tuple[str, int]:
  """Table of variables (shape, size, dtype) + total (reference :432-510)."""
  rows = []
  total = 0
  for key, v in model.vars.FlattenItems():
    n = int(np.prod(v.shape)) if v.dim() else 1
    total += n
    rows.append((getattr(v, 'var_name', key), tuple(v.shape), n,
                 str(v.dtype).split('.')[-1]))
  w_name = max([len(r[0]) for r in rows] + [4])
  w_shape = max([len(str(r[1])) for r in rows] + [5])
  lines = ['%-*s  %-*s  %-12s %s' % (w_name, 'name', w_shape, 'shape', 'size',
                                    'dtype')]
  for r in rows:
    lines.append('%-*s  %-*s  %-12d %s' % (w_name, r[0], w_shape, str(r[1]),
                                          r[2], r[3]))
  lines.append('')
  lines.append('=' * 30)
  lines.append('total #params: %10d' % total)
  lines.append('')
  return '\n'.join(lines), total
