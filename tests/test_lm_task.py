"""LM task end-to-end on CPU: text files → native batcher → RnnLm / TransformerLm."""

import numpy as np
import pytest
import torch

from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core import tokenizers
from lingvo_b200.models.lm import input_generator as lm_inp
from lingvo_b200.models.lm import layers as lm_layers
from lingvo_b200.models.lm import model as lm_model


@pytest.fixture(scope='module')
def corpus(tmp_path_factory):
  d = tmp_path_factory.mktemp('lm')
  rng = np.random.RandomState(0)
  words = ['the', 'cat', 'sat', 'on', 'a', 'mat', 'dog', 'ran']
  with open(d / 'train.txt', 'w') as f:
    for _ in range(400):
      f.write(' '.join(rng.choice(words, rng.randint(2, 7))) + '\n')
  return 'text:' + str(d / 'train.txt')


def _Input(corpus):
  p = lm_inp.LmInput.Params().Set(
      name='inp', file_pattern=corpus, bucket_upper_bound=[20, 40],
      bucket_batch_limit=[16, 8], file_buffer_size=64, file_parallelism=1,
      num_batcher_threads=2, target_max_length=40)
  p.tokenizer = tokenizers.AsciiTokenizer.Params()
  return p


@pytest.mark.parametrize('kind', ['rnn', 'transformer'])
def test_lm_trains(corpus, kind):
  p = lm_model.LanguageModel.Params().Set(name='lm_task', input=_Input(corpus))
  if kind == 'rnn':
    p.lm = lm_layers.RnnLm.CommonParams(vocab_size=76, emb_dim=16, num_layers=2,
                                        rnn_dims=16)
  else:
    p.lm = lm_layers.TransformerLm.CommonParams(
        model_dim=16, hidden_dim=32, num_heads=2, num_layers=2, vocab_size=76,
        residual_dropout_prob=0.0)
  p.lm.params_init = py_utils.WeightInit.Xavier(1.0) if kind == 'transformer' else \
      py_utils.WeightInit.Uniform(0.1)
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 3e-3
  p.train.l2_regularizer_weight = None
  task = p.Instantiate()
  losses = []
  for _ in range(25):
    metrics, _ = task.TrainStep()
    losses.append(float(metrics['log_pplx'][0]))
  assert losses[0] > 3.0                       # ≈ ln(76) at init
  assert min(losses[-5:]) < losses[0] - 0.5, losses
  assert float(metrics['num_predictions'][0]) > 0
