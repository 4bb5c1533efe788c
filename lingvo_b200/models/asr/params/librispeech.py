"""Librispeech 960h configs (ref `lingvo/tasks/asr/params/librispeech.py`)."""

import os

import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.core import layers
from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core import schedule
from lingvo_b200.core import tokenizers
from lingvo_b200.models.asr import input_generator
from lingvo_b200.models.asr import model


@model_registry.RegisterSingleTaskModel
class Librispeech960Base(base_model_params.SingleTaskModelParams):
  """Grapheme LAS baseline (ref :29)."""

  DATADIR = os.environ.get('LINGVO_B200_LIBRISPEECH', '/tmp/librispeech')

  def _CommonInputParams(self, is_eval):
    p = input_generator.AsrInput.Params()
    p.frame_size = 80
    p.append_eos_frame = True
    p.pad_to_max_seq_length = False
    p.file_random_seed = 0
    p.file_buffer_size = 10000
    p.file_parallelism = 16
    if is_eval:
      p.source_max_length = 3600
      p.bucket_upper_bound = [639, 1062, 1275, 1377, 1449, 1506, 1563, 3600]
    else:
      p.source_max_length = 3000
      p.bucket_upper_bound = [639, 1062, 1275, 1377, 1449, 1506, 1563, 1710]
    p.bucket_batch_limit = [96, 48, 48, 48, 48, 48, 48, 48]
    return p

  def SetBucketSizes(self, params, bucket_upper_bound, bucket_batch_limit):
    params.bucket_upper_bound = bucket_upper_bound
    params.bucket_batch_limit = bucket_batch_limit
    return params

  def Train(self):
    p = self._CommonInputParams(False)
    p.file_datasource = None
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'train/train.tfrecords-*')
    p.num_samples = 281241
    return p

  def Dev(self):
    p = self._CommonInputParams(True)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'devtest/dev-clean.tfrecords-00000-of-00001')
    p.num_samples = 2703
    return p

  def Devother(self):
    p = self._CommonInputParams(True)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'devtest/dev-other.tfrecords-00000-of-00001')
    p.num_samples = 2864
    return p

  def Test(self):
    p = self._CommonInputParams(True)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'devtest/test-clean.tfrecords-00000-of-00001')
    p.num_samples = 2620
    return p

  def Testother(self):
    p = self._CommonInputParams(True)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'devtest/test-other.tfrecords-00000-of-00001')
    p.num_samples = 2939
    return p

  def Task(self):
    p = model.AsrModel.Params()
    p.name = 'librispeech'
    tp = p.train
    tp.learning_rate = 1e-3
    tp.lr_schedule = schedule.ContinuousSchedule.Params().Set(
        start_step=50000, half_life_steps=100000, min=0.01)
    p.vn.global_vn = True
    tp.vn_start_step = 20000
    tp.vn_std = 0.075
    tp.l2_regularizer_weight = 1e-6
    tp.clip_gradient_norm_to_value = 1.0
    tp.grad_norm_to_clip_to_zero = 100.0
    tp.tpu_steps_per_loop = 20

    ep = p.encoder
    ep.input_shape = [None, None, 80, 1]
    ep.lstm_cell_size = 1024
    ep.num_lstm_layers = 4
    ep.conv_filter_shapes = [(3, 3, 1, 32), (3, 3, 32, 32)]
    ep.conv_filter_strides = [(2, 2), (2, 2)]
    ep.cnn_tpl.params_init = py_utils.WeightInit.Gaussian(0.001)
    ep.num_conv_lstm_layers = 0
    ep.use_specaugment = True

    dp = p.decoder
    dp.source_dim = 2048
    dp.emb_dim = 96
    dp.emb.vocab_size = 76
    dp.emb.max_num_shards = 1
    dp.rnn_cell_dim = 1024
    dp.rnn_layers = 2
    dp.attention.hidden_dim = 128
    dp.softmax.num_classes = 76
    dp.target_seq_len = 620
    dp.label_smoothing = layers.UniformLabelSmoother.Params().Set(
        num_classes=76, uncertainty=0.1)
    dp.beam_search.num_hyps_per_beam = 8
    return p


@model_registry.RegisterSingleTaskModel
class Librispeech960Grapheme(Librispeech960Base):
  """Grapheme targets (ref :156)."""

  GRAPHEME_TARGET_SEQUENCE_LENGTH = 620
  GRAPHEME_VOCAB_SIZE = 76

  def InitializeTokenizer(self, params):
    params.tokenizer = tokenizers.AsciiTokenizer.Params()
    params.tokenizer.vocab_size = self.GRAPHEME_VOCAB_SIZE
    params.target_max_length = self.GRAPHEME_TARGET_SEQUENCE_LENGTH
    return params

  def _CommonInputParams(self, is_eval):
    return self.InitializeTokenizer(super()._CommonInputParams(is_eval))


class _StaticShapeMixin:
  """Fixed-shape batches (ref `…TpuV2` variants, :217/:310): every training batch is
  padded to the longest bucket with a constant batch size, so the whole step has one
  shape signature and can be replayed from a CUDA graph."""

  def _CommonInputParams(self, is_eval):
    p = super()._CommonInputParams(is_eval)
    if not is_eval:
      p.pad_to_max_seq_length = True
      p.bucket_batch_limit = [48] * len(p.bucket_upper_bound)
      p.source_max_length = p.bucket_upper_bound[-1]
    return p

  def Task(self):
    p = super().Task()
    p.encoder.pad_steps = 0
    p.decoder.emb.max_num_shards = 1
    return p


@model_registry.RegisterSingleTaskModel
class Librispeech960GraphemeTpuV2(_StaticShapeMixin, Librispeech960Grapheme):
  """Static-shape grapheme model (ref :217)."""


@model_registry.RegisterSingleTaskModel
class Librispeech960Wpm(Librispeech960Base):
  """16k word-piece targets (ref :239)."""

  WPM_SYMBOL_TABLE_FILEPATH = os.path.join(
      os.path.dirname(__file__), '..', 'wpm_16k_librispeech.vocab')
  WPM_TARGET_SEQUENCE_LENGTH = 140
  WPM_VOCAB_SIZE = 16328
  EMBEDDING_DIMENSION = 96
  NUM_TRAINING_WORKERS = 8

  def _CommonInputParams(self, is_eval):
    p = super()._CommonInputParams(is_eval)
    p.tokenizer = tokenizers.WpmTokenizer.Params().Set(
        vocab_filepath=self.WPM_SYMBOL_TABLE_FILEPATH, vocab_size=self.WPM_VOCAB_SIZE)
    p.target_max_length = self.WPM_TARGET_SEQUENCE_LENGTH
    return p

  def Task(self):
    p = super().Task()
    dp = p.decoder
    dp.emb.vocab_size = self.WPM_VOCAB_SIZE
    dp.emb_dim = self.EMBEDDING_DIMENSION
    dp.softmax.num_classes = self.WPM_VOCAB_SIZE
    dp.target_seq_len = self.WPM_TARGET_SEQUENCE_LENGTH
    dp.label_smoothing.num_classes = self.WPM_VOCAB_SIZE
    return p


@model_registry.RegisterSingleTaskModel
class Librispeech960WpmTpuV2(_StaticShapeMixin, Librispeech960Wpm):
  """Static-shape word-piece model (ref :310)."""


@model_registry.RegisterSingleTaskModel
class Librispeech960ConformerWpm(Librispeech960Wpm):
  """Conformer encoder (17 × D=512, 8 heads, kernel 32, relative attention) + the LAS
  word-piece decoder: BASELINE.json config #4 ("asr.librispeech Conformer encoder bf16").
  The reference ships the block (`core/conformer_layer.py:471`) but registers no model
  with it; hyper-parameters follow Conformer-L of the paper."""

  CONFORMER_DIM = 512
  CONFORMER_LAYERS = 17
  CONFORMER_HEADS = 8
  CONFORMER_KERNEL = 32

  def Task(self):
    from lingvo_b200.models.asr import encoder as asr_encoder
    p = super().Task()
    old = p.encoder
    p.encoder = asr_encoder.ConformerEncoder.Params().Set(
        name='enc', model_dim=self.CONFORMER_DIM, num_layers=self.CONFORMER_LAYERS,
        num_heads=self.CONFORMER_HEADS, kernel_size=self.CONFORMER_KERNEL,
        use_specaugment=True, dropout_prob=0.1, input_shape=list(old.input_shape))
    p.encoder.fprop_dtype = torch.bfloat16
    p.decoder.source_dim = self.CONFORMER_DIM
    tp = p.train
    tp.learning_rate = 1e-3
    tp.lr_schedule = schedule.TransformerSchedule.Params().Set(
        warmup_steps=10000, model_dim=self.CONFORMER_DIM)
    tp.optimizer = optimizer.Adam.Params().Set(beta1=0.9, beta2=0.98, epsilon=1e-9)
    tp.vn_std = 0.0
    return p


@model_registry.RegisterSingleTaskModel
class Librispeech960ConformerWpmTpuV2(_StaticShapeMixin, Librispeech960ConformerWpm):
  """Static-shape variant: one shape signature ⇒ whole-step CUDA-graph replay."""


@model_registry.RegisterSingleTaskModel
class Librispeech960ConformerTiny(Librispeech960Grapheme):
  """2-block Conformer for tests / smoke runs."""

  def Task(self):
    from lingvo_b200.models.asr import encoder as asr_encoder
    p = super().Task()
    old = p.encoder
    p.encoder = asr_encoder.ConformerEncoder.Params().Set(
        name='enc', model_dim=32, num_layers=2, num_heads=2, kernel_size=8,
        dropout_prob=0.0, input_shape=list(old.input_shape))
    p.decoder.source_dim = 32
    p.decoder.rnn_cell_dim = 32
    p.decoder.attention.hidden_dim = 16
    p.decoder.emb_dim = 16
    return p
