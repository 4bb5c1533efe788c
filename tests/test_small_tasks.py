"""Punctuator + Milan tasks on CPU."""

import numpy as np
import pytest
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import layers
from lingvo_b200.core import optimizer
from lingvo_b200.core import schedule
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.milan import dual_encoder
from lingvo_b200.models.milan import labels as label_lib
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.punctuator import input_generator as punct_inp
from lingvo_b200.models.punctuator import model as punct_model


def test_multi_label_contrastive_loss():
  logits = torch.tensor([[2.0, 0.0, 0.0], [0.0, 1.0, 3.0]])
  labels = torch.tensor([[1., 0, -1], [0, 1, 1]])
  loss = label_lib.MultiLabelContrastiveLoss(labels, logits)
  want0 = torch.log(torch.exp(torch.tensor(2.)) + 1) - 2.0
  want1 = torch.logsumexp(logits[1], 0) - torch.logsumexp(logits[1, 1:], 0)
  torch.testing.assert_close(loss, torch.stack([want0, want1]))


def test_punctuator_pipeline_and_training(tmp_path):
  text = tmp_path / 'train.txt'
  lines = ['Hello, world.', 'The cat sat; the dog ran!', 'Yes? No.', 'It is, however, fine.'] * 30
  text.write_text('\n'.join(lines))
  inp = punct_inp.PunctuatorInput.Params().Set(name='inp', bucket_upper_bound=[40],
                                               bucket_batch_limit=[8])
  inp.file_datasource.file_pattern = str(text)
  inp.file_datasource.shuffle_buffer_size = 16
  inp.tokenizer = tokenizers.AsciiTokenizer.Params()
  p = base_config.SetupRNMTParams(
      punct_model.RNMTModel.Params(), name='punct', vocab_size=76, embedding_dim=16,
      hidden_dim=16, num_heads=2, num_encoder_layers=1, num_decoder_layers=2,
      learning_rate=5e-3, l2_regularizer_weight=None, lr_warmup_steps=1, lr_decay_start=10**6,
      lr_decay_end=2 * 10**6, lr_min=0.5, ls_uncertainty=0.0, atten_dropout_prob=0.0,
      residual_dropout_prob=0.0, adam_beta2=0.98, adam_epsilon=1e-6)
  p.input = inp
  p.train.lr_schedule = schedule.Constant.Params()
  task = p.Instantiate()
  batch = task.input.GetPreprocessedInputBatch()
  src = task.input.IdsToStrings(batch.src.ids, (1 - batch.src.paddings).sum(1).long())
  assert all(s == s.lower() and ',' not in s and '.' not in s for s in src)
  l0 = float(task.TrainStep()[0]['log_pplx'][0])
  for _ in range(25):
    m, _ = task.TrainStep()
  assert float(m['log_pplx'][0]) < l0
  # string-in / string-out inference subgraph (ref punctuator/model.py:37)
  sub = task.Inference()
  assert list(sub) == ['default']
  out = sub['default'](['hello world', 'yes no'])
  k = out.topk_scores.shape[1]
  assert len(out.topk_decoded) == 2 and all(len(h) == k for h in out.topk_decoded)
  assert all(isinstance(s, str) for h in out.topk_decoded for s in h)
  assert out.src_ids.shape[0] == 2 and out.topk_ids.shape[:2] == (2, k)
  assert (out.topk_scores[:, :-1] >= out.topk_scores[:, 1:] - 1e-5).all()   # best first
  assert isinstance(task.Punctuate('hello world'), str)


class _PairInput(base_input_generator.BaseInputGenerator):

  def _InputBatch(self):
    g = torch.Generator().manual_seed(int(torch.randint(0, 1 << 30, ()).item()))
    z = torch.randn(16, 6, generator=g)
    return NestedMap(image=z @ self._a, text=z @ self._b)

  def __init__(self, params):
    super().__init__(params)
    g = torch.Generator().manual_seed(0)
    self._a = torch.randn(6, 12, generator=g)
    self._b = torch.randn(6, 10, generator=g)


def test_milan_dual_encoder_learns_alignment():
  cfg_i = dual_encoder.EncoderConfig().Set(
      input_features='image', output_dim=8,
      encoder=layers.FCLayer.Params().Set(input_dim=12, output_dim=8, activation='NONE'))
  cfg_t = dual_encoder.EncoderConfig().Set(
      input_features='text', output_dim=8,
      encoder=layers.FCLayer.Params().Set(input_dim=10, output_dim=8, activation='NONE'))
  p = dual_encoder.MilanTask.Params()
  p.input = _PairInput.Params().Set(name='pairs', batch_size=16)
  p.dual_encoder.Set(encoder_configs={'image': cfg_i, 'text': cfg_t}, joint_embedding_dim=6,
                     loss_weights={('image', 'text'): 0.5, ('text', 'image'): 0.5},
                     initial_temperature=0.2)
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 2e-2
  p.train.lr_schedule = schedule.Constant.Params()
  task = p.Instantiate()
  first = None
  for i in range(80):
    m, _ = task.TrainStep()
    if first is None:
      first = float(m['loss'][0])
  assert float(m['loss'][0]) < 0.6 * first
  assert float(m['recall_at_1_image_to_text'][0]) > 0.5


class _BlobInput(base_input_generator.BaseInputGenerator):
  """Two Gaussian blobs rendered as tiny 4×4 single-channel images."""

  def _InputBatch(self):
    y = torch.randint(0, 2, (32,))
    x = torch.randn(32, 4, 4, 1) * 0.5 + (y.float() * 2 - 1).reshape(32, 1, 1, 1)
    return NestedMap(data=x, label=y, weight=torch.ones(32))


def test_image_classifier_v2_trains_and_serves():
  from lingvo_b200.models.image import classifier
  torch.manual_seed(0)
  assert float(classifier.TopKAccuracy(
      2, torch.tensor([[0.1, 0.9, 0.5], [0.8, 0.1, 0.3]]), torch.tensor([2, 1]), torch.ones(2))) == 0.5
  p = classifier.ModelV2.Params().Set(name='v2', label_smoothing=0.1)
  p.extract = layers.FCLayer.Params().Set(name='fc', input_dim=1, output_dim=3, activation='TANH')
  p.softmax = layers.SimpleFullSoftmax.Params().Set(name='softmax', input_dim=4 * 4 * 3, num_classes=2)
  p.input = _BlobInput.Params().Set(name='blobs', batch_size=32)
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 1e-2
  p.train.lr_schedule = schedule.Constant.Params()
  task = p.Instantiate()
  for _ in range(60):
    m, _ = task.TrainStep()
  assert float(m['accuracy'][0]) > 0.95 and float(m['error'][0]) < 0.05
  sub = task.Inference()['default']
  out = sub(torch.ones(3, 4, 4, 1))
  assert out.prediction.tolist() == [1, 1, 1] and out.probs.shape == (3, 2)
  torch.testing.assert_close(out.probs.sum(-1), torch.ones(3))
