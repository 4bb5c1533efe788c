"""Punctuator codelab configs (ref `lingvo/tasks/punctuator/params/codelab.py`)."""

import os

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.core import tokenizers
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.punctuator import input_generator
from lingvo_b200.models.punctuator import model


class BrownCorpusWPM(base_model_params.SingleTaskModelParams):
  """Brown-corpus text with a 16k word-piece vocabulary (ref :28)."""

  _DATADIR = os.environ.get('LINGVO_B200_PUNCTUATOR', '/tmp/punctuator_data')
  _VOCAB_FILE = os.path.join(os.path.dirname(__file__), 'brown_corpus_wpm.16000.vocab')
  _VOCAB_SIZE = 16000

  def _Input(self, name, is_eval):
    p = input_generator.PunctuatorInput.Params()
    p.file_datasource.file_pattern = os.path.join(self._DATADIR, name)
    p.tokenizer = tokenizers.WpmTokenizer.Params().Set(
        vocab_filepath=self._VOCAB_FILE, vocab_size=self._VOCAB_SIZE)
    p.source_max_length = 40
    p.target_max_length = 40
    p.bucket_upper_bound = [10, 20, 30, 60, 120]
    p.bucket_batch_limit = [16] * 4 + [4] if is_eval else [512, 256, 160, 80, 40]
    return p

  def Train(self):
    p = self._Input('train.txt', False)
    p.num_samples = 51094
    return p

  def Test(self):
    p = self._Input('test.txt', True)
    p.num_samples = 1000
    return p


@model_registry.RegisterSingleTaskModel
class RNMTModel(BrownCorpusWPM):
  """RNMT+ punctuator (ref :89)."""

  def Task(self):
    p = base_config.SetupRNMTParams(
        model.RNMTModel.Params(), name='punctuator_rnmt', vocab_size=self._VOCAB_SIZE,
        embedding_dim=128, hidden_dim=512, num_heads=4, num_encoder_layers=2,
        num_decoder_layers=2, learning_rate=1e-4, l2_regularizer_weight=1e-5,
        lr_warmup_steps=500, lr_decay_start=400000, lr_decay_end=1200000, lr_min=0.5,
        ls_uncertainty=0.1, atten_dropout_prob=0.3, residual_dropout_prob=0.3,
        adam_beta2=0.98, adam_epsilon=1e-6)
    p.eval.samples_per_summary = 2466
    return p
