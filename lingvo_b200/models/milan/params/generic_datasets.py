"""Generic image-text TFRecord datasets (ref
`lingvo/tasks/milan/params/generic_datasets.py`)."""

import json
import logging
import os

from lingvo_b200.core import hyperparams
from lingvo_b200.models.milan import common_schema
from lingvo_b200.models.milan import constants
from lingvo_b200.models.milan import dataset_spec
from lingvo_b200.models.milan import labels as label_lib
from lingvo_b200.models.milan import utils


def _SimpleImageCaptionDatasetLabeler(image_id_feature):
  """Positives on the diagonal; off-diagonal pairs sharing an image id (duplicates and
  co-captions) are dropped from the loss rather than used as negatives (ref :30)."""
  return label_lib.ExamplePairLabeler(drop_pairs_that_match=image_id_feature)


class ImageTextTFRecords(dataset_spec.TFRecordDatasetSpec):
  """See `common_schema` for the on-disk format (ref :60)."""

  @classmethod
  def Params(cls):
    p = hyperparams.InstantiableParams(cls)
    p.Define('split_paths', {constants.Split.TRAIN: 'train-*', constants.Split.DEV: 'dev-*',
                             constants.Split.TEST: 'test-*'}, 'split → file pattern.')
    p.Define('data_dir', '', 'Base directory of relative split paths.')
    p.Define('bert_max_length', 48, 'Token length of the pre-computed BERT features.')
    p.Define('bert_dim', 768, 'BERT feature dim.')
    return p

  @classmethod
  def ParamsFromEnv(cls, environment_variable='MILAN_DATASET_CONFIG_JSON',
                    die_if_unset=False):
    cfg = os.getenv(environment_variable)
    if cfg is not None:
      return cls.Params().Set(**json.loads(cfg))
    msg = '%s: set %s to configure the dataset.' % (cls.__name__, environment_variable)
    if die_if_unset:
      raise ValueError(msg)
    logging.warning(msg)
    return cls.Params().Set(data_dir='/please-set-%s-to-configure-dataset' % environment_variable)

  def __init__(self, params):
    self.params = params
    paths = {k: os.path.join(params.data_dir, v) for k, v in params.split_paths.items()}
    schema = common_schema.ImageFeatures()
    schema.update(common_schema.TextFeatures(
        bert_embeddings_shape=(params.bert_max_length, params.bert_dim)))
    # one image / caption per example: squeeze the item dim for the single-item labeler
    self._squeeze = ('image/id', 'text/id', 'text/bert/lengths', 'text/bert/embeddings',
                     'image/encoded', 'text/captions')
    super().__init__(paths, schema,
                     _SimpleImageCaptionDatasetLabeler(image_id_feature='image/id'))

  def _ParseRecord(self, record):
    ex = super()._ParseRecord(record)
    for k in self._squeeze:
      if k in ex:
        ex[k] = ex[k][0]
    if 'text/bert/embeddings' in ex:
      import torch
      t = utils.PadOrTrimDimension(torch.as_tensor(ex['text/bert/embeddings']),
                                   self.params.bert_max_length, axis=-2)
      ex['text/bert/token_features'] = t.numpy()
    return ex
