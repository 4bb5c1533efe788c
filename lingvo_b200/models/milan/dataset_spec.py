"""Dataset descriptions (ref `lingvo/tasks/milan/dataset_spec.py`).

A `DatasetSpec` knows how to `Read` a split into an iterator of feature dicts and how to
`Label` example pairs drawn from it; `TFRecordDatasetSpec` reads sharded TFRecord files of
`tf.train.Example`s through the native record yielders.
"""

from __future__ import annotations

import abc
import glob

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.models.milan import constants
from lingvo_b200.models.milan import labels as label_lib
from lingvo_b200.utils import tf_example


class Metadata:
  """Static facts about a dataset (ref :37)."""

  def __init__(self, features, modality_to_features=None, modality_batch_shapes=None):
    self.features = dict(features)
    self.modality_to_features = dict(modality_to_features or {})
    self.modality_batch_shapes = dict(modality_batch_shapes or self._InferDefaultBatchShapes())

  def _InferDefaultBatchShapes(self):
    """[None] + the leading item dims of a modality's first feature."""
    out = {}
    for modality, names in self.modality_to_features.items():
      names = [names] if isinstance(names, str) else list(names)
      shape = self.features[names[0]][0]
      out[modality] = (None,) + tuple(shape[:1])
    return out


class DatasetSpec(metaclass=abc.ABCMeta):
  """ref :57."""

  @abc.abstractmethod
  def Read(self, split, batch_size=None, shuffle=False, seed=0, **kwargs):
    """→ iterator of feature dicts (unbatched) or of stacked batches."""

  @abc.abstractmethod
  def Label(self, pairs: label_lib.ExamplePairs):
    """→ int labels `[query_batch, result_batch, …]`."""

  @property
  @abc.abstractmethod
  def meta(self) -> Metadata:
    """Dataset metadata."""


class FileBasedDatasetSpec(DatasetSpec):
  """A dataset stored as files per split (ref :87)."""

  def __init__(self, split_paths, schema, label_fn, metadata=None, reader_fn=None):
    self._split_paths = dict(split_paths)
    self._schema = dict(schema)
    self._label_fn = label_fn
    self._meta = metadata or Metadata(schema)
    self._reader_fn = reader_fn

  def _Files(self, split):
    if split not in self._split_paths:
      raise ValueError('Unknown split %r; have %s' % (split, sorted(self._split_paths)))
    pats = self._split_paths[split]
    pats = [pats] if isinstance(pats, str) else list(pats)
    files = sorted(f for p in pats for f in (glob.glob(p) or [p]))
    return files

  def _ParseRecord(self, record):
    raw = tf_example.ParseExample(record)
    out = {}
    for name, (shape, dtype) in self._schema.items():
      if name not in raw:
        raise KeyError('feature %s missing from record (has %s)' % (name, sorted(raw)))
      v = raw[name]
      if dtype is bytes:
        v = list(v)
        out[name] = v if shape else v[0]
      else:
        out[name] = np.asarray(v, dtype).reshape(shape)
    return out

  def Read(self, split, batch_size=None, shuffle=False, seed=0, num_epochs=None, **kwargs):
    del kwargs
    files = self._Files(split)
    pattern = 'tfrecord:' + ','.join(files)
    train = split == constants.Split.TRAIN
    epochs = num_epochs if num_epochs is not None else (0 if train else 1)
    h = ops.host()
    if shuffle:
      y = h.basic_record_yielder(pattern, seed=seed + 1, bufsize=4096, parallelism=2,
                                 num_epochs=epochs)
    else:
      y = h.sequential_record_yielder(pattern, epochs if epochs else -1)
    def _Examples():
      while True:
        rec = y.next()
        if rec is None:
          return
        yield self._ParseRecord(rec[0])
    if not batch_size:
      return _Examples()
    def _Batches():
      buf = []
      for ex in _Examples():
        buf.append(ex)
        if len(buf) == batch_size:
          yield _Stack(buf)
          buf = []
      if buf and not train:
        yield _Stack(buf)
    return _Batches()

  def Label(self, pairs):
    return self._label_fn(pairs)

  @property
  def meta(self):
    return self._meta


def _Stack(examples):
  out = {}
  for k in examples[0]:
    vals = [e[k] for e in examples]
    out[k] = vals if isinstance(vals[0], (bytes, str, list)) else np.stack(vals)
  return out


class TFRecordDatasetSpec(FileBasedDatasetSpec):
  """TFRecord-backed dataset (ref :223)."""

  def __init__(self, split_paths, schema, label_fn, **kwargs):
    super().__init__(split_paths, schema, label_fn, **kwargs)
