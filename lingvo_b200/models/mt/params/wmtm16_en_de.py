"""WMT'16 multimodal-task En→De captions, text only (ref
`lingvo/tasks/mt/params/wmtm16_en_de.py`): a 29k-sentence toy that reaches >30 BLEU in
a few thousand steps — the quick sanity benchmark of the MT stack."""

import os

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.mt import input_generator
from lingvo_b200.models.mt import model


@model_registry.RegisterSingleTaskModel
class WmtCaptionEnDeTransformer(base_model_params.SingleTaskModelParams):
  """2-layer Transformer on a 2k word-piece vocabulary (ref :27)."""

  DATADIR = os.environ.get('LINGVO_B200_WMTM16', '/tmp/wmtm16/wpm/')
  VOCAB_SIZE = 2000
  VOCAB_FILE = 'wpm-ende-2k.voc'

  def _CommonInputParams(self, is_eval):
    p = input_generator.NmtInput.Params()
    if is_eval:
      p.Set(file_random_seed=27182818, file_parallelism=1, file_buffer_size=1,
            bucket_upper_bound=[10, 14, 19, 26, 36, 50, 70, 98, 137, 200],
            bucket_batch_limit=[16] * 8 + [4] * 2)
    else:
      p.Set(file_random_seed=0, file_parallelism=1, file_buffer_size=29000,
            bucket_upper_bound=[14, 17, 20, 24, 29, 35, 45, 75],
            bucket_batch_limit=[292, 240, 204, 170, 141, 117, 91, 54])
    p.tokenizer.vocab_size = self.VOCAB_SIZE
    p.tokenizer.token_vocab_filepath = os.path.join(self.DATADIR, self.VOCAB_FILE)
    return p

  def _Split(self, is_eval, fname, n):
    p = self._CommonInputParams(is_eval)
    p.file_pattern = 'tfrecord:' + os.path.join(self.DATADIR, fname)
    p.num_samples = n
    return p

  def Train(self):
    return self._Split(False, 'train.tfrecords', 29000)

  def Dev(self):
    return self._Split(True, 'val.tfrecords', 1014)

  def Test(self):
    return self._Split(True, 'test.tfrecords', 1000)

  def Task(self):
    p = base_config.SetupTransformerParams(
        model.TransformerModel.Params(), name='wmt14_en_de_transformer_base',
        vocab_size=self.VOCAB_SIZE, model_dim=256, hidden_dim=512, num_heads=2,
        num_layers=2, residual_dropout_prob=0.2, input_dropout_prob=0.2,
        learning_rate=1.0, warmup_steps=1000)
    p.eval.samples_per_summary = 7500
    p.train.save_interval_seconds = 60
    p.train.max_steps = 12000
    return p


@model_registry.RegisterSingleTaskModel
class WmtCaptionEnDeTransformerCloudTpu(WmtCaptionEnDeTransformer):
  """Static-shape variant (ref :95): every batch padded to the longest bucket with a
  fixed batch size — one shape signature, CUDA-graph friendly."""

  def _CommonInputParams(self, is_eval):
    p = super()._CommonInputParams(is_eval)
    p.pad_to_max_seq_length = True
    p.source_max_length = p.bucket_upper_bound[-1]
    p.bucket_batch_limit = [16] * len(p.bucket_batch_limit)
    return p

  def Task(self):
    p = super().Task()
    for emb in (p.encoder.token_emb, p.decoder.token_emb):
      if 'max_num_shards' in emb:          # sharded EmbeddingLayer templates only
        emb.max_num_shards = 1
    return p
