"""Matplotlib figure summaries (ref `lingvo/core/plot.py`): render numpy data into image
summaries (`MatplotlibFigureSummary`, `Image`, `Scatter`, `Curve`). Matplotlib is optional;
without it the helpers return None and summaries are skipped."""
import io

import numpy as np

try:
  import matplotlib  # pylint: disable=g-import-not-at-top
  matplotlib.use('Agg')
  from matplotlib import pyplot as plt  # pylint: disable=g-import-not-at-top
  _HAS_MPL = True
except Exception:  # pylint: disable=broad-except
  plt = None
  _HAS_MPL = False


def ToUnicode(text):
  return text.decode('utf-8') if isinstance(text, bytes) else text


def AddPlot(unused_fig, axes, data, title=u'', xlabel=u'', ylabel=u'', fontsize='small',
            xlim=None, ylim=None, suppress_xticks=False, suppress_yticks=False):
  axes.plot(data)
  axes.set_title(ToUnicode(title), size=fontsize)
  axes.set_xlabel(ToUnicode(xlabel), size=fontsize)
  axes.set_ylabel(ToUnicode(ylabel), size=fontsize)
  if xlim:
    axes.set_xlim(xlim)
  if ylim:
    axes.set_ylim(ylim)
  if suppress_xticks:
    axes.set_xticks([])
  if suppress_yticks:
    axes.set_yticks([])


def AddImage(fig, axes, data, cmap='bone_r', clim=None, show_colorbar=True, title=u'',
             xlabel=u'', ylabel=u'', fontsize='small', origin='lower', suppress_xticks=False,
             suppress_yticks=False, aspect='auto', vmin=None, vmax=None):
  image = axes.imshow(data, cmap=cmap, origin=origin, aspect=aspect, interpolation='nearest',
                      vmin=vmin, vmax=vmax)
  if show_colorbar:
    fig.colorbar(image, ax=axes)
  if clim is not None:
    image.set_clim(clim)
  axes.set_title(ToUnicode(title), size=fontsize)
  axes.set_xlabel(ToUnicode(xlabel), size=fontsize)
  axes.set_ylabel(ToUnicode(ylabel), size=fontsize)
  if suppress_xticks:
    axes.set_xticks([])
  if suppress_yticks:
    axes.set_yticks([])


def AddScatterPlot(unused_fig, axes, xs, ys, title=u'', xlabel=u'', ylabel=u'',
                   fontsize='small', xlim=None, ylim=None, **kwargs):
  axes.scatter(xs, ys, **kwargs)
  axes.set_title(ToUnicode(title), size=fontsize)
  axes.set_xlabel(ToUnicode(xlabel), size=fontsize)
  axes.set_ylabel(ToUnicode(ylabel), size=fontsize)
  if xlim:
    axes.set_xlim(xlim)
  if ylim:
    axes.set_ylim(ylim)


def FigureToPng(fig):
  buf = io.BytesIO()
  fig.savefig(buf, format='png')
  plt.close(fig)
  return buf.getvalue()


class MatplotlibFigureSummary:
  """Collects subplots and renders one PNG per batch element."""

  def __init__(self, name, figsize=(8, 10), max_outputs=3, subplot_grid_shape=None,
               gridspec_kwargs=None, plot_func=AddImage, shared_subplot_kwargs=None):
    self._name, self._figsize, self._max = name, figsize, max_outputs
    self._grid, self._plot_func = subplot_grid_shape, plot_func
    self._shared = shared_subplot_kwargs or {}
    self._subplots = []

  def AddSubplot(self, tensor_list, plot_func=None, **kwargs):
    merged = dict(self._shared)
    merged.update(kwargs)
    self._subplots.append((tensor_list, plot_func or self._plot_func, merged))

  def Finalize(self):
    """→ list of PNG bytes (one per example) or None without matplotlib."""
    if not _HAS_MPL or not self._subplots:
      return None
    n = min(self._max, min(len(np.asarray(t[0][0])) for t in self._subplots))
    grid = self._grid or (len(self._subplots), 1)
    out = []
    for i in range(n):
      fig = plt.figure(figsize=self._figsize)
      for k, (tensors, fn, kw) in enumerate(self._subplots):
        ax = fig.add_subplot(grid[0], grid[1], k + 1)
        fn(fig, ax, *[np.asarray(t)[i] for t in tensors], **kw)
      out.append(FigureToPng(fig))
    return out


def Image(name, figsize, image, setter=None, **kwargs):
  if not _HAS_MPL:
    return None
  fig = plt.figure(figsize=figsize)
  ax = fig.add_subplot(1, 1, 1)
  AddImage(fig, ax, np.asarray(image), **kwargs)
  if setter:
    setter(fig, ax)
  return FigureToPng(fig)


def Custom(name, figsize, setter):
  """One-axes figure drawn entirely by `setter(fig, axes)` → PNG bytes (None without
  matplotlib)."""
  del name
  if not _HAS_MPL:
    return None
  fig = plt.figure(figsize=figsize)
  ax = fig.add_subplot(1, 1, 1)
  setter(fig, ax)
  return FigureToPng(fig)


def Scatter(name, figsize, xs, ys, setter=None, **kwargs):
  if not _HAS_MPL:
    return None
  fig = plt.figure(figsize=figsize)
  ax = fig.add_subplot(1, 1, 1)
  AddScatterPlot(fig, ax, np.asarray(xs), np.asarray(ys), **kwargs)
  if setter:
    setter(fig, ax)
  return FigureToPng(fig)


def Curve(name, figsize, xs, ys, setter=None, **kwargs):
  if not _HAS_MPL:
    return None
  fig = plt.figure(figsize=figsize)
  ax = fig.add_subplot(1, 1, 1)
  ax.plot(np.asarray(xs), np.asarray(ys), **kwargs)
  if setter:
    setter(fig, ax)
  return FigureToPng(fig)


def _AddMultiCurveRowPlots(fig, axes, data, length, x_label_override=None, row_labels=None,
                           title=u'', xlabel=u'', ylabel=u'', fontsize='small'):
  """One line per row of `data [rows, time]`, cut at `length` (ref :497)."""
  del fig
  colors = ['b-', 'r-', 'g-', 'm-', 'y-']
  for row in range(data.shape[0]):
    label = row_labels[row] if row_labels else '{}'.format(row)
    axes.plot(data[row, :int(length)], colors[row % len(colors)], label=label)
  axes.set_xlim([0, int(length)])
  axes.legend()
  axes.set_title(ToUnicode(title), size=fontsize)
  if x_label_override is not None:
    axes.set_xlabel(ToUnicode(x_label_override), size='x-small', wrap=True)
  else:
    axes.set_xlabel(ToUnicode(xlabel), size=fontsize)
  axes.set_ylabel(ToUnicode(ylabel), size=fontsize)


def MultiCurveData(tensors, paddings, labels):
  """Stacks the non-None `[batch, length]` tensors (zeroed under their paddings) into
  `[batch, rows, length]` and returns (data, per-example max length, row labels)."""
  tensors = [None if t is None else np.asarray(t, np.float64) for t in tensors]
  if not isinstance(paddings, (list, tuple)):
    paddings = [paddings] * len(tensors)
  paddings = [np.asarray(p, np.float64) for p in paddings]
  max_lengths = np.zeros(paddings[0].shape[0], np.int32)
  data, row_labels = [], []
  for t, l, p in zip(tensors, labels, paddings):
    max_lengths = np.maximum(max_lengths, np.round((1.0 - p).sum(1)).astype(np.int32))
    if t is not None:
      data.append(t * (1.0 - p))
      row_labels.append(l)
  return np.stack(data, 1), max_lengths, row_labels


def AddMultiCurveSubplot(fig, tensors, paddings, labels, xlabels=None, **kwargs):
  """Adds to the `MatplotlibFigureSummary` `fig` a subplot with one labelled curve per tensor
  (`[batch, length]` each; `paddings` one `[batch, length]` array or one per tensor),
  optionally with a per-example x label (ref :456)."""
  data, max_lengths, row_labels = MultiCurveData(tensors, paddings, labels)
  args = [data, max_lengths]
  if xlabels is not None:
    args.append(np.asarray(xlabels))
  fig.AddSubplot(args, plot_func=_AddMultiCurveRowPlots, row_labels=row_labels, **kwargs)


def FigureToSummary(name, fig):
  """A matplotlib figure → serialized Summary with one `<name>/image` PNG value (ref :341)."""
  from lingvo_b200.utils import tfevents  # pylint: disable=g-import-not-at-top
  return tfevents.ImageValue('%s/image' % name, FigureToPng(fig))
