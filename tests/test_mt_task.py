"""MT task on CPU: tf.Example records → NmtInput → Transformer / RNMT, train on a
copy task, then beam-search decode."""

import numpy as np
import pytest
import torch

from lingvo_b200 import ops
from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core import schedule
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.mt import input_generator
from lingvo_b200.models.mt import model as mt_model
from lingvo_b200.utils import tf_example

VOCAB = 16


@pytest.fixture(scope='module')
def records(tmp_path_factory):
  d = tmp_path_factory.mktemp('mt')
  rng = np.random.RandomState(0)
  w = ops.host().TFRecordWriter(str(d / 'train.tfrecords-00000'))
  for _ in range(300):
    n = rng.randint(3, 8)
    src = rng.randint(3, VOCAB, n)
    ex = tf_example.MakeExample({
        'source_id': src, 'source_padding': np.zeros(n, np.float32),
        'target_id': np.concatenate([[1], src]), 'target_padding': np.zeros(n + 1, np.float32),
        'target_label': np.concatenate([src, [2]]), 'target_weight': np.ones(n + 1, np.float32)})
    w.write(ex)
  w.close()
  vocab = d / 'vocab.txt'
  vocab.write_text('\n'.join(['<unk>', '<s>', '</s>'] + ['w%d' % i for i in range(3, VOCAB)]))
  return 'tfrecord:' + str(d / 'train.tfrecords-*'), str(vocab)


def test_tf_example_roundtrip():
  ex = tf_example.MakeExample({'a': np.asarray([1, -2, 300]), 'b': np.asarray([0.5, 2.0]),
                               'c': [b'xy']})
  out = tf_example.ParseExample(ex)
  assert out['a'].tolist() == [1, -2, 300] and out['b'].tolist() == [0.5, 2.0]
  assert out['c'][0] == b'xy'


def _Input(records):
  pattern, vocab = records
  p = input_generator.NmtInput.Params().Set(
      name='inp', file_pattern=pattern, bucket_upper_bound=[10], bucket_batch_limit=[16],
      file_buffer_size=64, file_parallelism=1, num_batcher_threads=2)
  p.tokenizer.token_vocab_filepath = vocab
  p.tokenizer.load_token_ids_from_vocab = False
  p.tokenizer.vocab_size = VOCAB
  return p


def test_transformer_mt_learns_copy_and_decodes(records):
  p = base_config.SetupTransformerParams(
      mt_model.TransformerModel.Params(), name='mt', vocab_size=VOCAB, model_dim=32,
      hidden_dim=64, num_heads=2, num_layers=2, learning_rate=3e-3, warmup_steps=1,
      residual_dropout_prob=0.0, label_smoothing_uncertainty=0.0)
  p.input = _Input(records)
  p.train.lr_schedule = schedule.Constant.Params()
  p.train.optimizer = optimizer.Adam.Params()
  p.decoder.target_seq_len = 10
  p.decoder.beam_search.num_hyps_per_beam = 2
  p.decoder.beam_search.sync_every = 1
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  task = p.Instantiate()
  losses = []
  for _ in range(150):
    m, _ = task.TrainStep()
    losses.append(float(m['log_pplx'][0]))
  assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])
  batch = task.input.GetPreprocessedInputBatch()
  out = task.Decode(batch)
  b = batch.src.ids.shape[0]
  assert out.topk_ids.shape[0] == b * 2
  dm = task.CreateDecoderMetrics()
  task.PostProcessDecodeOut(out, dm)
  assert dm['corpus_bleu'].value > 0.3, dm['corpus_bleu'].value


def test_rnmt_trains(records):
  p = base_config.SetupRNMTParams(
      mt_model.RNMTModel.Params(), name='rnmt', vocab_size=VOCAB, embedding_dim=16,
      hidden_dim=16, num_heads=2, num_encoder_layers=2, num_decoder_layers=2,
      learning_rate=5e-3, l2_regularizer_weight=None, lr_warmup_steps=1,
      lr_decay_start=100000, lr_decay_end=200000, lr_min=0.5, ls_uncertainty=0.0,
      atten_dropout_prob=0.0, residual_dropout_prob=0.0, adam_beta2=0.98, adam_epsilon=1e-6)
  p.input = _Input(records)
  p.train.lr_schedule = schedule.Constant.Params()
  p.decoder.target_seq_len = 10
  p.decoder.beam_search.num_hyps_per_beam = 2
  task = p.Instantiate()
  losses = []
  for _ in range(40):
    m, _ = task.TrainStep()
    losses.append(float(m['log_pplx'][0]))
  assert min(losses[-5:]) < losses[0] - 0.2, (losses[0], losses[-5:])
  out = task.Decode(task.input.GetPreprocessedInputBatch())
  assert out.topk_scores.shape[1] == 2


def test_experimental_decode_program_overlaps_postprocessing(records, tmp_path):
  """`ExperimentalDecodeProgram` (reference program.py:1807): device decode of batch i+1
  overlaps host post-processing of batch i; same artefacts as `DecodeProgram`."""
  import os
  import pickle
  from lingvo_b200.core import base_model
  from lingvo_b200.core import program
  task_p = base_config.SetupTransformerParams(
      mt_model.TransformerModel.Params(), name='mt', vocab_size=VOCAB, model_dim=16,
      hidden_dim=32, num_heads=2, num_layers=1, learning_rate=1e-3, warmup_steps=1,
      residual_dropout_prob=0.0, label_smoothing_uncertainty=0.0)
  task_p.input = _Input(records)
  task_p.decoder.target_seq_len = 6
  task_p.decoder.beam_search.num_hyps_per_beam = 2
  task_p.decoder.beam_search.sync_every = 1
  cfg = base_model.SingleTaskModel.Params(task_p)
  outs = {}
  for cls in (program.DecodeProgram, program.ExperimentalDecodeProgram):
    logdir = str(tmp_path / cls.__name__)
    pp = cls.Params().Set(name='decode', task=cfg, logdir=logdir, dataset_name='Test',
                          steps_per_loop=3)
    prog = pp.Instantiate()
    prog.BuildTpuSubgraph()
    assert prog.Run() is False
    files = [f for f in os.listdir(prog._program_dir) if f.startswith('decoder_out_')]
    assert len(files) == 1
    with open(os.path.join(prog._program_dir, files[0]), 'rb') as f:
      outs[cls.__name__] = pickle.load(f)
    assert 'corpus_bleu' in prog.last_metrics
    # second run on the same step is skipped through the decode-status cache
    assert prog.Run() is False
  assert len(outs['DecodeProgram']) == len(outs['ExperimentalDecodeProgram']) > 0
