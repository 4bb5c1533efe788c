"""WMT'14 En→De configs (ref `lingvo/tasks/mt/params/wmt14_en_de.py`)."""

import os

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.mt import input_generator
from lingvo_b200.models.mt import model


@model_registry.RegisterSingleTaskModel
class WmtEnDeTransformerBase(base_model_params.SingleTaskModelParams):
  """Transformer-base on word-piece tf.Example records (ref :27)."""

  DATADIR = os.environ.get('LINGVO_B200_WMT14', '/tmp/wmt14/wpm/')
  VOCAB_SIZE = 32000

  def _CommonInputParams(self, is_eval):
    p = input_generator.NmtInput.Params()
    p.tokenizer.vocab_size = self.VOCAB_SIZE
    p.tokenizer.token_vocab_filepath = os.path.join(self.DATADIR, 'wpm-ende.voc')
    if is_eval:
      p.file_random_seed = 27182818
      p.file_parallelism = 1
      p.file_buffer_size = 1
      p.bucket_upper_bound = [10, 14, 19, 26, 36, 50, 70, 98, 137, 200]
      p.bucket_batch_limit = [16] * 8 + [4] * 2
    else:
      p.file_random_seed = 0
      p.file_parallelism = 16
      p.file_buffer_size = 10000000
      p.bucket_upper_bound = [8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96]
      p.bucket_batch_limit = [512, 409, 341, 292, 256, 204, 170, 146, 128, 102, 85, 73, 64,
                              51, 42]
    return p

  def Train(self):
    p = self._CommonInputParams(False)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'train.tfrecords-*')
    p.num_samples = 4492447
    return p

  def Dev(self):
    p = self._CommonInputParams(True)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'dev.tfrecords')
    p.num_samples = 3000
    return p

  def Test(self):
    p = self._CommonInputParams(True)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, 'test.tfrecords')
    p.num_samples = 2737
    return p

  def Task(self):
    p = base_config.SetupTransformerParams(
        model.TransformerModel.Params(), name='wmt14_en_de_transformer_base',
        vocab_size=self.VOCAB_SIZE, model_dim=512, hidden_dim=2048, num_heads=8,
        num_layers=6, residual_dropout_prob=0.1, input_dropout_prob=0.1,
        learning_rate=3.0, warmup_steps=40000)
    p.eval.samples_per_summary = 7500
    return p


@model_registry.RegisterSingleTaskModel
class WmtEnDeTransformerSmall(WmtEnDeTransformerBase):
  """Small Transformer for quick experiments (ref :100)."""

  def Task(self):
    p = base_config.SetupTransformerParams(
        model.TransformerModel.Params(), name='wmt14_en_de_transformer_small',
        vocab_size=self.VOCAB_SIZE, model_dim=64, hidden_dim=128, num_heads=2,
        num_layers=2, residual_dropout_prob=0.1, input_dropout_prob=0.1,
        learning_rate=3.0, warmup_steps=40000)
    p.eval.samples_per_summary = 7500
    return p


@model_registry.RegisterSingleTaskModel
class WmtEnDeRNMT(WmtEnDeTransformerBase):
  """RNMT+ (ref :141)."""

  def _CommonInputParams(self, is_eval):
    p = super()._CommonInputParams(is_eval)
    if is_eval:
      return base_config.InitTestDatasetParams(self.VOCAB_SIZE, p)
    return base_config.InitTrainDatasetParams(self.VOCAB_SIZE, p)

  def Task(self):
    p = base_config.SetupRNMTParams(
        model.RNMTModel.Params(), name='wmt14_en_de_rnmtplus_base',
        vocab_size=self.VOCAB_SIZE, embedding_dim=1024, hidden_dim=1024, num_heads=4,
        num_encoder_layers=6, num_decoder_layers=8, learning_rate=1e-4,
        l2_regularizer_weight=1e-5, lr_warmup_steps=500, lr_decay_start=400000,
        lr_decay_end=1200000, lr_min=0.5, ls_uncertainty=0.1, atten_dropout_prob=0.3,
        residual_dropout_prob=0.3, adam_beta2=0.98, adam_epsilon=1e-6)
    p.eval.samples_per_summary = 7500
    return p


@model_registry.RegisterSingleTaskModel
class WmtEnDeTransformerBigGPipe(WmtEnDeTransformerBase):
  """Transformer-big (d = 1024, ff = 4096, 16 heads, 6 + 6 layers, shared 32k word pieces)
  as one GPipe pipeline of `GPUS` cells — BASELINE.json config #5. The reference registers
  no MT GPipe model; this assembles `layers_with_gpipe.GPipeTransformerStack` (:576) the
  way `lm.one_billion_wds.OneBWdsGPipeTransformerWPM` does for the LM."""

  MODEL_DIM = 1024
  HIDDEN_DIM = 4096
  NUM_HEADS = 16
  NUM_LAYERS = 6
  GPUS = 4
  NUM_MICRO_BATCHES = 8

  def _CommonInputParams(self, is_eval):
    p = super()._CommonInputParams(is_eval)
    if not is_eval:
      # one bucket, fixed batch: every micro-batch has the same shape (static pipeline links)
      p.bucket_upper_bound = [96]
      p.bucket_batch_limit = [128]
      p.pad_to_max_seq_length = True
      p.source_max_length = 96
      p.target_max_length = 96
    return p

  def Task(self):
    from lingvo_b200.core import layers_with_gpipe
    from lingvo_b200.core import optimizer
    from lingvo_b200.core import schedule
    p = model.GPipeTransformerModel.Params().Set(name='wmt14_en_de_transformer_big_gpipe')
    st = p.stack
    st.Set(name='stack', model_dim=self.MODEL_DIM, num_encoder_layers=self.NUM_LAYERS,
           num_decoder_layers=self.NUM_LAYERS, use_pipelined_embeddings=True,
           num_splits=self.GPUS, splits=self.GPUS, num_micro_batches=self.NUM_MICRO_BATCHES)
    st.emb_tpl.Set(vocab_size=self.VOCAB_SIZE, model_dim=self.MODEL_DIM,
                   input_dropout_prob=0.1, max_seq_len=1024)
    st.softmax_tpl.Set(num_classes=self.VOCAB_SIZE, input_dim=self.MODEL_DIM)
    for tpl in (st.encoder_tpl, st.decoder_tpl):
      tpl.tr_atten_tpl.num_attention_heads = self.NUM_HEADS
      tpl.tr_atten_tpl.residual_dropout_prob = 0.1
      tpl.tr_fflayer_tpl.hidden_dim = self.HIDDEN_DIM
      tpl.tr_fflayer_tpl.residual_dropout_prob = 0.1
    p.train.Set(learning_rate=3.0, optimizer=optimizer.Adam.ParamsB(),
                clip_gradient_norm_to_value=0.0, grad_norm_to_clip_to_zero=0.0,
                lr_schedule=schedule.TransformerSchedule.Params().Set(
                    warmup_steps=40000, worker_replicas=1, model_dim=self.MODEL_DIM))
    p.eval.samples_per_summary = 7500
    return p


@model_registry.RegisterSingleTaskModel
class WmtEnDeTransformerGPipeTiny(WmtEnDeTransformerBigGPipe):
  """2-cell toy variant for tests."""
  MODEL_DIM = 32
  HIDDEN_DIM = 64
  NUM_HEADS = 2
  NUM_LAYERS = 1
  GPUS = 2
  NUM_MICRO_BATCHES = 2
  VOCAB_SIZE = 64
