"""WMT'14 XEnDec experiments (arXiv 2106.04060; ref
`lingvo/tasks/mt/params/xendec/wmt14_en_de.py`)."""

import os

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.mt import input_generator
from lingvo_b200.models.mt import model


@model_registry.RegisterSingleTaskModel
class WmtEnDeXEnDec(base_model_params.SingleTaskModelParams):
  """En→De with crossover training (ref :29)."""

  DATADIR = os.environ.get('LINGVO_B200_WMT14_XENDEC', '/tmp/wmt14ende/')
  DATATRAIN = 'train.xendec.tfrecord-?????-of-?????'
  DATADEV = 'newstest2013.xendec.tfrecord-00000-of-00001'
  DATATEST = 'newstest2014.xendec.tfrecord-00000-of-00001'
  VOCAB = 'wpm32k.vocab'
  PACKED_INPUT = False
  vocab_size = 32000

  # model
  num_heads, model_dim, hidden_dim = 8, 512, 2048
  residual_dropout_prob, input_dropout_prob = 0.1, 0.1
  atten_dropout_prob, relu_dropout_prob = 0.0, 0.0
  # crossover data
  source_mask_ratio, source_mask_ratio_beta = -1, '2,6'
  mask_word_id, pad_id = 5, 6
  mask_words_ratio, permutation_distance = 0.5, 3
  # objective
  loss_mix_weight = loss_clean_weight = loss_mono_weight = 1.0
  use_prob_cl, use_prob_drop, atten_drop = True, False, 0.2
  # optimisation
  batch_size_ratio, learning_rate, warmup_steps = 1, 1.0, 4000
  num_samples = 4506303

  def Train(self):
    p = input_generator.NmtDoubleInput.Params()
    p.Set(file_random_seed=0, file_parallelism=64, file_buffer_size=10000000,
          natural_order_model=True, num_samples=self.num_samples)
    p.bucket_upper_bound = [8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112,
                            128, 160, 192, 224, 256]
    limits = [512, 409, 341, 292, 256, 204, 170, 146, 128, 102, 85, 73, 64, 51, 42, 36, 32,
              25, 21, 18, 16]
    p.bucket_batch_limit = [max(int(b * self.batch_size_ratio), 1) for b in limits]
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, self.DATATRAIN)
    p.tokenizer.token_vocab_filepath = os.path.join(self.DATADIR, self.VOCAB)
    p.tokenizer.vocab_size = self.vocab_size
    p.Set(source_mask_ratio=self.source_mask_ratio,
          source_mask_ratio_beta=self.source_mask_ratio_beta,
          mask_word_id=self.mask_word_id, pad_id=self.pad_id,
          mask_words_ratio=self.mask_words_ratio,
          permutation_distance=self.permutation_distance, packed_input=self.PACKED_INPUT,
          vocab_file=os.path.join(self.DATADIR, self.VOCAB))
    return p

  def _EvalParams(self, fname, n):
    p = input_generator.NmtInput.Params()
    p.tokenizer.vocab_size = self.vocab_size
    p.tokenizer.token_vocab_filepath = os.path.join(self.DATADIR, self.VOCAB)
    p.Set(file_random_seed=27182818, file_parallelism=1, file_buffer_size=1,
          bucket_upper_bound=[10, 14, 19, 26, 36, 50, 70, 98, 137, 300],
          bucket_batch_limit=[16] * 8 + [4] * 2, num_samples=n,
          file_pattern='tfrecord:' + os.path.join(self.DATADIR, fname))
    return p

  def Dev(self):
    return self._EvalParams(self.DATADEV, 3000)       # newstest2013

  def Test(self):
    return self._EvalParams(self.DATATEST, 3003)      # newstest2014

  def Task(self):
    p = base_config.SetupXEnDecTransformerParams(
        model.TransformerXEnDecModel.Params(), name='transformer',
        vocab_size=self.vocab_size, model_dim=self.model_dim, hidden_dim=self.hidden_dim,
        num_heads=self.num_heads, num_layers=6,
        residual_dropout_prob=self.residual_dropout_prob,
        input_dropout_prob=self.input_dropout_prob,
        atten_dropout_prob=self.atten_dropout_prob, relu_dropout_prob=self.relu_dropout_prob,
        learning_rate=self.learning_rate, warmup_steps=self.warmup_steps)
    p.Set(loss_mix_weight=self.loss_mix_weight, loss_clean_weight=self.loss_clean_weight,
          loss_mono_weight=self.loss_mono_weight, use_prob_cl=self.use_prob_cl,
          atten_drop=self.atten_drop, use_prob_drop=self.use_prob_drop)
    p.train.save_keep_checkpoint_every_n_hours = 1.0 / 6
    p.decoder.beam_search.length_normalization = 0.6
    p.decoder.beam_search.beam_size = 4
    return p


@model_registry.RegisterSingleTaskModel
class WmtDeEnXEnDec(WmtEnDeXEnDec):
  """De→En (ref :171)."""

  DATADIR = os.environ.get('LINGVO_B200_WMT14_XENDEC_DEEN', '/tmp/wmt14deen')
