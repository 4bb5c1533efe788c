"""Common `tf.train.Example` layout for image-caption(-like) data (ref
`lingvo/tasks/milan/common_schema.py`).

A schema is a dict `feature name → (shape, dtype)`; `dataset_spec.TFRecordDatasetSpec`
parses records against it with the in-repo protobuf wire codec.
"""

from __future__ import annotations

import numpy as np


def _Feature(shape, dtype=np.float32):
  return (tuple(shape), dtype)


def ImageFeatures(images_per_example=1):
  """Encoded image bytes and ids (ref :31)."""
  n = images_per_example
  return {
      'image/encoded': _Feature([n], bytes),
      'image/id': _Feature([n], np.int64),
  }


def TextFeatures(captions_per_example=1, bert_embeddings_shape=None):
  """Caption strings / ids and optional pre-computed BERT features (ref :47)."""
  n = captions_per_example
  feats = {
      'text/captions': _Feature([n], bytes),
      'text/id': _Feature([n], np.int64),
  }
  if bert_embeddings_shape is not None:
    max_len, dim = bert_embeddings_shape
    feats['text/bert/lengths'] = _Feature([n], np.int64)
    feats['text/bert/embeddings'] = _Feature([n, max_len, dim], np.float32)
  return feats


def AudioFeatures(mfcc_shape=None, cpc8k_shape=None):
  """Optional audio features (ref :80)."""
  feats = {}
  if mfcc_shape is not None:
    feats['audio/mfcc'] = _Feature(mfcc_shape, np.float32)
    feats['audio/mfcc/lengths'] = _Feature([], np.int64)
  if cpc8k_shape is not None:
    feats['audio/cpc8k/features'] = _Feature(cpc8k_shape, np.float32)
    feats['audio/cpc8k/lengths'] = _Feature([], np.int64)
  return feats
