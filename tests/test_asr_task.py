"""ASR task on CPU: frontend, SpecAugment, encoder→LAS decoder train + decode + WER."""

import numpy as np
import pytest
import torch

from lingvo_b200 import ops
from lingvo_b200.core import optimizer
from lingvo_b200.core import schedule
from lingvo_b200.core import spectrum_augmenter
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.asr import decoder_utils
from lingvo_b200.models.asr import frontend
from lingvo_b200.models.asr import input_generator
from lingvo_b200.models.asr import model as asr_model
from lingvo_b200.utils import tf_example


def test_edit_distance():
  assert decoder_utils.EditDistance('the cat sat', 'the cat sat')[3] == 0
  ins, subs, dels, tot = decoder_utils.EditDistance('a b c d', 'a x c')
  assert tot == 2 and subs == 1 and dels == 1
  assert decoder_utils.ComputeWer(['hello there'], ['hello world'])[0] == (1, 2)


def test_mel_frontend_shapes_and_tone():
  p = frontend.MelAsrFrontend.Params().Set(num_bins=40, noise_scale=0.0)
  fe = p.Instantiate()
  t = torch.arange(16000).float() / 16000
  pcm = torch.stack([torch.sin(2 * np.pi * 440 * t), torch.sin(2 * np.pi * 3000 * t)]) * 1000
  out = fe.FPropDefaultTheta(NestedMap(src_inputs=pcm, paddings=torch.zeros(2, 16000)))
  assert out.src_inputs.shape[0] == 2 and out.src_inputs.shape[2] == 40
  assert out.src_inputs.shape[1] == out.paddings.shape[1] == (16000 - 401) // 160 + 1
  # the 3 kHz tone peaks in a higher mel bin than the 440 Hz tone
  peak = out.src_inputs[:, 10:-10, :, 0].mean(1).argmax(-1)
  assert int(peak[1]) > int(peak[0])


def test_spec_augment_masks_only_in_training():
  p = spectrum_augmenter.SpectrumAugmenter.Params().Set(
      name='aug', freq_mask_max_bins=5, freq_mask_count=2, time_mask_max_frames=10,
      time_mask_count=2, random_seed=3)
  aug = p.Instantiate()
  x = torch.ones(3, 50, 20, 1)
  pad = torch.zeros(3, 50)
  y, _ = aug.FPropDefaultTheta(x, pad)
  frac = float((y == 0).float().mean())
  assert 0.0 < frac < 0.8
  # masks are rectangular: a masked frequency bin is masked for every frame
  col = (y[0, :, :, 0] == 0).all(0)
  row = (y[0, :, :, 0] == 0).all(1)
  assert bool(col.any() or row.any())
  aug2 = p.Copy().Set(is_inference=True)
  from lingvo_b200.core import cluster_factory
  with cluster_factory.SetEval(True):
    y2, _ = aug.FPropDefaultTheta(x, pad)
  assert torch.equal(y2, x)


@pytest.fixture(scope='module')
def asr_records(tmp_path_factory):
  d = tmp_path_factory.mktemp('asr')
  rng = np.random.RandomState(0)
  w = ops.host().TFRecordWriter(str(d / 'train.tfrecords-00000'))
  words = {'aa': 0, 'bb': 1, 'cc': 2}
  for _ in range(200):
    seq = rng.choice(list(words), rng.randint(1, 4))
    frames = []
    for wd in seq:          # each word = 6 frames with a distinctive feature pattern
      f = np.zeros((6, 8), np.float32)
      f[:, words[wd] * 2:words[wd] * 2 + 2] = 1.0
      frames.append(f + 0.05 * rng.randn(6, 8).astype(np.float32))
    frames = np.concatenate(frames)
    w.write(tf_example.MakeExample({
        'uttid': [b'u'], 'transcript': [' '.join(seq).encode()], 'frames': frames.reshape(-1)}))
  w.close()
  return 'tfrecord:' + str(d / 'train.tfrecords-*')


def test_asr_trains_and_decodes(asr_records):
  inp = input_generator.AsrInput.Params().Set(
      name='inp', file_pattern=asr_records, frame_size=8, bucket_upper_bound=[40],
      bucket_batch_limit=[16], file_buffer_size=32, file_parallelism=1,
      num_batcher_threads=2, target_max_length=16)
  inp.tokenizer = tokenizers.AsciiTokenizer.Params()
  p = asr_model.AsrModel.Params().Set(name='asr', input=inp)
  ep = p.encoder
  ep.input_shape = [None, None, 8, 1]
  ep.conv_filter_shapes = [(3, 3, 1, 4)]
  ep.conv_filter_strides = [(2, 2)]
  ep.num_cnn_layers = 1
  ep.lstm_cell_size = 16
  ep.num_lstm_layers = 1
  ep.pad_steps = 0
  dp = p.decoder
  dp.source_dim = 32
  dp.emb_dim = 8
  dp.emb.vocab_size = 76
  dp.emb.max_num_shards = 1
  dp.rnn_cell_dim = 24
  dp.rnn_layers = 2
  dp.attention.hidden_dim = 16
  dp.softmax.num_classes = 76
  dp.target_seq_len = 14
  dp.beam_search.num_hyps_per_beam = 2
  dp.beam_search.sync_every = 2
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 1e-2
  p.train.lr_schedule = schedule.Constant.Params()
  p.train.vn_std = 0.0
  p.train.l2_regularizer_weight = None
  task = p.Instantiate()
  losses = []
  for _ in range(60):
    m, _ = task.TrainStep()
    losses.append(float(m['log_pplx'][0]))
  assert min(losses[-5:]) < 0.6 * losses[0], (losses[0], losses[-5:])
  batch = task.input.GetPreprocessedInputBatch()
  out = task.Decode(batch)
  dm = task.CreateDecoderMetrics()
  kv = task.PostProcessDecodeOut(out, dm)
  assert len(kv) == batch.src.src_inputs.shape[0]
  assert 0.0 <= dm['wer'].value <= 3.0


def test_conformer_asr_model_trains(asr_records):
  """BASELINE config #4: Conformer encoder (core/conformer_layer.ConformerLayer) + LAS
  decoder through the regular AsrModel task."""
  from lingvo_b200.models.asr import encoder as asr_encoder
  inp = input_generator.AsrInput.Params().Set(
      name='inp', file_pattern=asr_records, frame_size=8, bucket_upper_bound=[40],
      bucket_batch_limit=[16], file_buffer_size=32, file_parallelism=1,
      num_batcher_threads=2, target_max_length=16)
  inp.tokenizer = tokenizers.AsciiTokenizer.Params()
  p = asr_model.AsrModel.Params().Set(name='asr', input=inp)
  p.encoder = asr_encoder.ConformerEncoder.Params().Set(
      name='enc', input_shape=[None, None, 8, 1], conv_filter_shapes=[(3, 3, 1, 4)],
      conv_filter_strides=[(2, 2)], model_dim=32, num_layers=2, num_heads=2, kernel_size=4,
      dropout_prob=0.0)
  dp = p.decoder
  dp.source_dim = 32
  dp.emb_dim = 8
  dp.emb.vocab_size = 76
  dp.emb.max_num_shards = 1
  dp.rnn_cell_dim = 24
  dp.rnn_layers = 2
  dp.attention.hidden_dim = 16
  dp.softmax.num_classes = 76
  dp.target_seq_len = 14
  dp.beam_search.num_hyps_per_beam = 2
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 5e-3
  p.train.lr_schedule = schedule.Constant.Params()
  p.train.vn_std = 0.0
  p.train.l2_regularizer_weight = None
  task = p.Instantiate()
  assert len(task.encoder.blocks) == 2
  losses = []
  for _ in range(50):
    m, _ = task.TrainStep()
    losses.append(float(m['log_pplx'][0].detach()))
  assert min(losses[-5:]) < 0.75 * losses[0], (losses[0], losses[-5:])
  out = task.Decode(task.input.GetPreprocessedInputBatch())
  assert out.topk_ids.shape[0] > 0


def test_registered_conformer_configs_instantiate():
  from lingvo_b200 import model_registry
  import lingvo_b200.models.asr.params.librispeech  # noqa: F401
  from lingvo_b200.core import py_utils
  cfg = model_registry.GetParams('asr.librispeech.Librispeech960ConformerWpm', 'Train')
  enc = cfg.task.encoder
  assert enc.num_layers == 17 and enc.model_dim == 512 and enc.num_heads == 8
  with py_utils.StubVariablesScope('meta'):
    task = cfg.task.Instantiate()
  n = sum(v.numel() for v in task.encoder.vars.Flatten())
  assert 1.0e8 < n < 1.4e8, n           # Conformer-L encoder ≈ 118 M parameters
