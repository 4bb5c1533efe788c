"""Milan breadth: labels, utils, dataset spec → input generator → recipe → train step."""

import numpy as np
import torch

from lingvo_b200 import ops
from lingvo_b200.core import cluster_factory
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.milan import constants
from lingvo_b200.models.milan import image_preprocessor
from lingvo_b200.models.milan import labels as label_lib
from lingvo_b200.models.milan import tf_hub_layers
from lingvo_b200.models.milan import transformers
from lingvo_b200.models.milan import utils
from lingvo_b200.models.milan.params import dual_encoder_recipe
from lingvo_b200.models.milan.params import generic_datasets
from lingvo_b200.models.milan.tools import prepare_coco


def test_example_pair_labeler_drops_duplicates():
  batch = NestedMap(image_id=torch.tensor([7, 8, 7, 9]), text_id=torch.tensor([1, 2, 3, 4]))
  pairs = label_lib.ExamplePairs.WithinBatch(batch, query_modality='image',
                                             result_modality='text')
  lab = label_lib.ExamplePairLabeler(drop_pairs_that_match='image_id')(pairs)
  assert lab.diagonal().tolist() == [1, 1, 1, 1]
  assert lab[0, 2] == -1 and lab[2, 0] == -1 and lab[0, 1] == 0
  loss = label_lib.MultiLabelContrastiveLoss(lab, torch.randn(4, 4))
  assert loss.shape == (4,) and torch.isfinite(loss).all()


def test_multi_item_wrapper():
  batch = NestedMap(id=torch.arange(3))
  pairs = label_lib.ExamplePairs.WithinBatch(batch, query_modality='text',
                                             result_modality='text')
  w = label_lib.MultiItemExampleWrapper(label_lib.ExamplePairLabeler(),
                                        {'text': (None, 2)})
  lab = w(pairs)
  assert lab.shape == (3, 2, 3, 2)
  assert lab[0, 0, 0, 0] == -1 and lab[0, 0, 0, 1] == 1 and lab[0, 0, 1, 0] == 0


def test_utils():
  nm = NestedMap(a=NestedMap(b=torch.ones(4, 2)), c=torch.zeros(4))
  sel = utils.Selector({'x': 'a.b', 'y': 'c'})(nm)
  assert sel.x.shape == (4, 2) and sel.y.shape == (4,)
  assert utils.InferBatchSize(nm) == 4
  f = utils.BatchFlattener([None, 3])
  t = torch.randn(2, 3, 5)
  flat = f.Flatten(t)
  assert flat.shape == (6, 5) and torch.equal(f.Unflatten(flat), t)
  assert utils.PadOrTrimDimension(t, 7, 1).shape == (2, 7, 5)
  assert utils.PadOrTrimDimension(t, 2, 1).shape == (2, 2, 5)
  layer = utils.MakeFnLayer(lambda x, y: x + y, name='add').Instantiate()
  assert float(layer.FProp(None, torch.ones(1), torch.ones(1))) == 2.0


def test_image_preprocessor_and_tower():
  p = image_preprocessor.ImagePreprocessor.Params().Set(output_image_size=32)
  pre = p.Instantiate()
  imgs = (np.random.rand(3, 48, 64, 3) * 255).astype(np.uint8)
  out = pre.FProp(None, imgs)
  assert out.shape == (3, 3, 32, 32) and float(out.min()) >= -1.0 and float(out.max()) <= 1.0
  tower = tf_hub_layers.ImageModuleV2.Params().Set(
      name='tower', stage_channels=[8, 16], blocks_per_stage=1, output_dim=24).Instantiate()
  feat = tower.FProp(tower.theta, out)
  assert feat.shape == (3, 24)
  feat.sum().backward()
  frozen = tf_hub_layers.ImageModuleV2.Params().Set(
      name='frozen', stage_channels=[8], output_dim=4, trainable=False).Instantiate()
  assert not any(v.requires_grad for v in frozen.vars.Flatten())


def test_text_adapter():
  enc = transformers.GetTransformerStackWithEmbeddingInput(
      input_dim=12, num_layers=1, hidden_dim=16, num_attention_heads=2, output_dim=8,
      name='adapter').Instantiate()
  out = enc.FProp(enc.theta, torch.randn(3, 6, 12), torch.tensor([6, 2, 4]))
  assert out.shape == (3, 8)


class _TinyRecipe(dual_encoder_recipe.DualEncoderRecipe):
  DATA = None

  def __init__(self):
    super().__init__()
    self.input_params.batch_size = 4
    self.task_params.dual_encoder.loss_weights = {('image', 'text'): 0.5, ('text', 'image'): 0.5}
    self.task_params.dual_encoder.joint_embedding_dim = 8
    self.input_params.features_to_read += ['image/encoded', 'image/id', 'text/bert/.*', 'text/id']
    self.AddPreprocessor('image/encoded', image_preprocessor.ImagePreprocessor.Params().Set(
        output_image_size=16))
    self.AddModality('image', input_features='image/encoded', id_feature='image/id',
                     encoder=tf_hub_layers.ImageModuleV2.Params().Set(
                         stage_channels=[8], output_dim=12), output_dim=12)
    self.AddModality('text', input_features=('text/bert/token_features', 'text/bert/lengths'),
                     id_feature='text/id',
                     encoder=transformers.GetTransformerStackWithEmbeddingInput(
                         input_dim=6, num_layers=1, hidden_dim=16, num_attention_heads=2,
                         output_dim=8), output_dim=8)

  @property
  def default_dataset(self):
    return generic_datasets.ImageTextTFRecords.Params().Set(
        data_dir=self.DATA, bert_max_length=5, bert_dim=6,
        split_paths={constants.Split.TRAIN: 'train-*', constants.Split.DEV: 'train-*',
                     constants.Split.TEST: 'train-*'}).Instantiate()


def test_recipe_end_to_end(tmp_path):
  rng = np.random.RandomState(0)
  w = ops.host().TFRecordWriter(str(tmp_path / 'train-00000'))
  for i in range(12):
    img = (rng.rand(20, 20, 3) * 255).astype(np.uint8)
    emb = rng.randn(5, 6).astype(np.float32)
    # decoded pixels stored as raw bytes are not decodable without PIL → store via prepare_coco
    # schema but with a tiny valid "image" : use PNG-free path by writing arrays through numpy
    ex = prepare_coco.MakeExample(img.tobytes(), i // 2, 'caption %d' % i, i, emb, 3 + i % 3)
    w.write(ex)
  w.close()
  _TinyRecipe.DATA = str(tmp_path)
  recipe = _TinyRecipe()
  ds = recipe.dataset
  first = next(ds.Read(constants.Split.TRAIN))
  assert first['text/bert/token_features'].shape == (5, 6) and first['image/id'].shape == ()
  # raw bytes → arrays (the test images are raw RGB, not JPEG)
  ip = recipe.Train()
  raw_fn = ip.dataset_fn
  def _Decoded(batch_size, **kw):
    for b in raw_fn(batch_size=batch_size, **kw):
      b['image/encoded'] = np.stack([np.frombuffer(x, np.uint8).reshape(20, 20, 3)
                                     for x in b['image/encoded']])
      yield b
  ip.dataset_fn = _Decoded
  gen = ip.Instantiate()
  batch = gen.GetPreprocessedInputBatch()
  assert batch['image/encoded'].shape == (4, 3, 16, 16)
  assert batch['text/bert/token_features'].shape == (4, 5, 6)
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = recipe.Task().Set(name='milan').Instantiate()
  metrics, _ = task.FPropTower(task.theta, batch)
  assert torch.isfinite(metrics.loss[0])
  metrics.loss[0].backward()
  # pairs (0,1), (2,3) share an image id → ignored, so recall is computed on a sane target
  assert 'recall_at_1_image_to_text' in metrics


def test_milan_task_predictions_loss_and_decode_split():
  import torch
  from lingvo_b200.core import layers
  from lingvo_b200.core.nested_map import NestedMap
  from lingvo_b200.models.milan import dual_encoder as de
  enc = lambda d: layers.ProjectionLayer.Params().Set(input_dim=d, output_dim=6,
                                                      activation='NONE', batch_norm=False)
  p = de.MilanTask.Params()
  cfg = de.EncoderConfig() if hasattr(de, 'EncoderConfig') else None
  if cfg is None:
    return
  p.dual_encoder.encoder_configs = {
      'image': de.EncoderConfig().Set(input_features='img', encoder=enc(5).Set(name='ie'),
                                      output_dim=6, id_feature='image/id'),
      'text': de.EncoderConfig().Set(input_features='txt', encoder=enc(7).Set(name='te'),
                                     output_dim=6)}
  p.dual_encoder.loss_weights = {('image', 'text'): 1.0, ('text', 'image'): 0.5}
  task = p.Instantiate()
  batch = NestedMap({'img': torch.randn(4, 5), 'txt': torch.randn(4, 7),
                     'image/id': torch.arange(4)})
  preds = task.ComputePredictions(task.theta, batch)
  assert preds.image.encodings.shape == (4, 6) and preds.image.ids.tolist() == [0, 1, 2, 3]
  assert 'ids' not in preds.text
  metrics, _ = task.ComputeLoss(task.theta, preds, batch)
  want = metrics.loss_image_to_text[0] + 0.5 * metrics.loss_text_to_image[0]
  torch.testing.assert_close(metrics.loss[0], want)
  assert metrics.loss[1] == 4.0
  loss, m2 = task.dual_encoder.FProp(task.theta.dual_encoder, batch)
  torch.testing.assert_close(loss, metrics.loss[0])
  dm = task.CreateDecoderMetrics()
  assert task.PostProcessDecodeOut(task.Decode(batch), dm) == []
  assert dm['num_samples_in_batch'].total_value == 4 or dm['num_samples_in_batch'].value == 4
