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


# --------------------------------------------------------------- decoder features (r2) --
def _Decoder(**kw):
  from lingvo_b200.models.asr import decoder as asr_decoder
  p = asr_decoder.AsrDecoder.Params().Set(
      name='dec', source_dim=6, emb_dim=4, rnn_cell_dim=8, rnn_layers=2, residual_start=2,
      target_seq_len=6, **kw)
  p.emb.vocab_size = 12
  p.emb.max_num_shards = 1
  p.attention.hidden_dim = 5
  p.softmax.num_classes = 12
  p.random_seed = 99
  dec = p.Instantiate()
  dec.InstantiateVariables()
  return dec


def _DecInputs(b=3, t=5, s=7):
  g = torch.Generator().manual_seed(1)
  enc = NestedMap(encoded=torch.randn(s, b, 6, generator=g), padding=torch.zeros(s, b))
  ids = torch.randint(1, 12, (b, t), generator=g)
  pad = torch.zeros(b, t)
  pad[0, 3:] = 1
  tgt = NestedMap(ids=ids, labels=torch.roll(ids, -1, 1), paddings=pad, weights=1 - pad)
  return enc, tgt


def test_asr_decoder_step_plan_equals_sequence_plan():
  """Teacher forcing: unrolling with SingleDecodeStep gives the same logits / loss as the
  hoisted whole-sequence plan (same cells, same weights)."""
  dec = _Decoder()
  enc, tgt = _DecInputs()
  seq = dec.ComputePredictions(dec.theta, enc, tgt)
  assert 'logits' not in seq
  dyn = dec.ComputePredictionsDynamic(dec.theta, enc, tgt)
  torch.testing.assert_close(dyn.softmax_input, seq.softmax_input, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(dyn.attention.probs, seq.attention.probs, atol=1e-5, rtol=1e-5)
  m_seq, ps_seq = dec.ComputeLoss(dec.theta, seq, tgt)
  m_dyn, ps_dyn = dec.ComputeLoss(dec.theta, dyn, tgt)
  assert float(m_seq['loss'][0]) == pytest.approx(float(m_dyn['loss'][0]), rel=1e-5)
  torch.testing.assert_close(ps_seq.loss, ps_dyn.loss, atol=1e-5, rtol=1e-5)
  assert ps_seq.loss.shape == (3,)
  assert set(m_seq) >= {'loss', 'log_pplx', 'token_normed_prob',
                        'fraction_of_correct_next_step_preds', 'loss/logits'}
  assert float(m_seq['token_normed_prob'][0]) == pytest.approx(
      float(torch.exp(-m_seq['log_pplx'][0])), rel=1e-6)


def test_asr_decoder_loss_variants():
  import torch.nn.functional as F
  from lingvo_b200.core import layers
  from lingvo_b200.models.asr import decoder as asr_decoder
  enc, tgt = _DecInputs()
  base = _Decoder()
  pred = base.ComputePredictions(base.theta, enc, tgt)
  logits = base._ComputeLogits(base.theta, pred.softmax_input).transpose(0, 1)
  nll = F.cross_entropy(logits.reshape(-1, 12), tgt.labels.reshape(-1),
                        reduction='none').reshape(3, 5)
  w = tgt.weights
  m, ps = base.ComputeLoss(base.theta, pred, tgt)
  assert float(m['loss'][0]) == pytest.approx(float((nll * w).sum() / w.sum()), rel=1e-4)
  torch.testing.assert_close(ps.loss, (nll * w).sum(1), atol=1e-4, rtol=1e-4)
  # per-sequence averaging + length normalisation
  seq = _Decoder(per_token_avg_loss=False, token_normalized_per_seq_loss=True)
  m2, ps2 = seq.ComputeLoss(seq.theta, NestedMap(logits=logits), tgt)
  want = (nll * w).sum(1) / (w.sum(1) + 0.001)
  torch.testing.assert_close(ps2.loss, want, atol=1e-4, rtol=1e-4)
  assert float(m2['loss'][0]) == pytest.approx(float(want.mean()), rel=1e-4)
  assert float(m2['loss'][1]) == 3.0
  # focal loss down-weights confident tokens
  foc = _Decoder(focal_loss_gamma=2.0)
  _, ps3 = foc.ComputeLoss(foc.theta, NestedMap(logits=logits), tgt)
  p_t = torch.exp(-nll)
  torch.testing.assert_close(ps3.loss, (((1 - p_t) ** 2) * nll * w).sum(1), atol=1e-4,
                             rtol=1e-4)
  # label smoothing and explicit target distributions go through the soft-label path
  ls = _Decoder(label_smoothing=layers.UniformLabelSmoother.Params().Set(uncertainty=0.1))
  m4, _ = ls.ComputeLoss(ls.theta, NestedMap(logits=logits), tgt)
  lp = F.log_softmax(logits, -1)
  probs = torch.full((3, 5, 12), 0.1 / 11)
  probs.scatter_(-1, tgt.labels.unsqueeze(-1), 0.9)
  want4 = (-(probs * lp).sum(-1) * w).sum() / w.sum()
  assert float(m4['loss'][0]) == pytest.approx(float(want4), rel=1e-3)
  t2 = NestedMap(tgt)
  t2.probs = probs
  m5, _ = base.ComputeLoss(base.theta, NestedMap(logits=logits), t2)
  assert float(m5['loss'][0]) == pytest.approx(float(want4), rel=1e-3)
  # several heads with weights
  two = _Decoder(logit_types={'logits': 1.0, 'aux_logits': 0.5})
  m6, ps6 = two.ComputeLoss(two.theta, NestedMap(logits=logits, aux_logits=logits * 0.0), tgt)
  uniform = float(np.log(12.0))
  assert float(m6['loss/aux_logits'][0]) == pytest.approx(uniform, rel=1e-5)
  assert float(m6['loss'][0]) == pytest.approx(float(m['loss'][0]) + 0.5 * uniform, rel=1e-4)
  del asr_decoder


def test_asr_decoder_scheduled_sampling_feeds_its_own_samples():
  from lingvo_b200.core import py_utils
  dec = _Decoder(min_ground_truth_prob=0.0, prob_decay_start_step=0, min_prob_step=10)
  enc, tgt = _DecInputs()
  py_utils.SetGlobalStep(0)
  try:
    assert float(dec.GroundTruthProbability()) == 1.0
    torch.manual_seed(0)
    gt = dec.ComputePredictions(dec.theta, enc, tgt)        # step plan, always ground truth
    ref = dec.ComputePredictionsDynamic(dec.theta, enc, tgt)
    torch.testing.assert_close(gt.logits, ref.logits)
    py_utils.SetGlobalStep(5)
    assert float(dec.GroundTruthProbability()) == pytest.approx(0.5)
    py_utils.SetGlobalStep(10)
    assert float(dec.GroundTruthProbability()) == 0.0
    torch.manual_seed(0)
    ss = dec.ComputePredictions(dec.theta, enc, tgt)        # always its own samples
    # step 0 sees <s> either way; later steps are conditioned on sampled tokens
    torch.testing.assert_close(ss.logits[:, 0], gt.logits[:, 0])
    assert not torch.allclose(ss.logits[:, 1:], gt.logits[:, 1:])
    m, _ = dec.ComputeLoss(dec.theta, ss, tgt)
    m['loss'][0].backward()
    assert dec.vars.Flatten()[0].grad is not None
  finally:
    py_utils.SetGlobalStep(0)
  with pytest.raises(AssertionError):
    _Decoder(min_ground_truth_prob=0.5, use_while_loop_based_unrolling=False)


def test_asr_decoder_shallow_fusion_and_adapters():
  from lingvo_b200.models.asr import fusion
  from lingvo_b200.models.lm import layers as lm_layers
  enc, tgt = _DecInputs()
  plain = _Decoder()
  fused = _Decoder(fusion=fusion.ShallowFusion.Params().Set(
      lm=lm_layers.NullLm.Params().Set(vocab_size=12), lm_weight=0.5))
  a = plain.ComputePredictionsDynamic(plain.theta, enc, tgt)
  b = fused.ComputePredictions(fused.theta, enc, tgt)       # fusion forces the step plan
  # a uniform LM shifts every log-prob by the same constant: log_softmax(am) + λ·log(1/V)
  want = torch.log_softmax(a.logits, -1) + 0.5 * float(np.log(1 / 12.0))
  torch.testing.assert_close(b.logits, want, atol=1e-5, rtol=1e-5)
  # beam search runs through the same step function and the fusion state
  enc2, _ = _DecInputs()
  out = fused.BeamSearchDecodeWithTheta(fused.theta, enc2)
  assert out.topk_ids.shape[0] == 3 * fused.params.beam_search.num_hyps_per_beam
  # adapters: per-utterance task ids select per-task residual adapters after every layer
  from lingvo_b200.core import layers
  ad = _Decoder(adapter_task_id_field='domain_ids',
                adapter_layer_tpl=layers.MultitaskAdapterLayer.Params().Set(
                    num_tasks=2, bottleneck_dim=3))
  enc3, _ = _DecInputs()
  enc3.domain_ids = torch.tensor([0, 1, 0])
  assert len(ad.adapters) == 2
  with torch.no_grad():
    for v in ad.vars.Flatten():
      if 'adapter' in v.var_name and 'up' in v.var_name:
        v.copy_(torch.randn_like(v) * 0.5)
  pa = ad.ComputePredictions(ad.theta, enc3, tgt)
  enc3.domain_ids = torch.tensor([1, 1, 0])
  pb = ad.ComputePredictions(ad.theta, enc3, tgt)
  assert not torch.allclose(pa.logits[0], pb.logits[0])      # utterance 0 changed task
  torch.testing.assert_close(pa.logits[1:], pb.logits[1:])   # the others did not


def test_asr_decoder_functional_unrolling_and_init_helpers():
  dec = _Decoder()
  enc, tgt = _DecInputs()
  seq = dec.ComputePredictions(dec.theta, enc, tgt)
  fun = dec.ComputePredictionsFunctional(dec.theta, enc, tgt)
  torch.testing.assert_close(fun.softmax_input, seq.softmax_input, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(fun.attention.probs, seq.attention.probs, atol=1e-5, rtol=1e-5)
  # gradients flow through the scanned step function
  m, _ = dec.ComputeLoss(dec.theta, fun, tgt)
  m['loss'][0].backward()
  assert all(v.grad is not None for v in dec.vars.Flatten() if v.requires_grad)
  # flat init tuple = fields of the step zero state
  rnn, ctx, probs, atten, fusion, misc, packed = dec.InitDecoder(dec.theta, enc, 3)
  state, _ = dec.DecoderStepZeroState(dec.theta, enc,
                                      torch.full((3, 1), dec.params.target_sos_id), 3)
  torch.testing.assert_close(ctx, state.atten_context)
  assert len(rnn) == 2 and probs.shape == (3, 7) and packed is not None
  assert isinstance(misc, NestedMap) and fusion is not None and atten is not None
  base = dec.BaseZeroState(dec.theta, enc, 3, misc)
  torch.testing.assert_close(base[1], ctx)
  assert dec.CreateTargetInfoMisc(tgt) == NestedMap()
  tgt.fst_bias_probs = torch.zeros(3, 5)
  assert 'fst_bias_probs' in dec.CreateTargetInfoMisc(tgt)
  dec.AddAdditionalDecoderSummaries(enc, tgt, None, seq.softmax_input)


def _Encoder(**kw):
  from lingvo_b200.models.asr import encoder as asr_encoder
  p = asr_encoder.AsrEncoder.Params().Set(
      name='enc', input_shape=[None, None, 12, 1], conv_filter_shapes=[(3, 3, 1, 4), (3, 3, 4, 4)],
      lstm_cell_size=8, num_lstm_layers=3, pad_steps=2, random_seed=11, **kw)
  enc = p.Instantiate()
  enc.InstantiateVariables()
  return enc


def _EncBatch(b=2, t=20):
  g = torch.Generator().manual_seed(0)
  pad = torch.zeros(b, t)
  pad[1, 14:] = 1
  return NestedMap(src_inputs=torch.randn(b, t, 12, 1, generator=g), paddings=pad)


def test_asr_encoder_layout_options():
  base = _Encoder()
  out = base.FProp(base.theta, _EncBatch())
  # 22 frames (20 + pad_steps) → stride 4 → 6; frequency 12 → 3, ·4 channels = 12 → padded 16
  assert out.encoded.shape == (6, 2, 16) and out.padding.shape == (6, 2)
  assert base._first_lstm_input_dim == 16 and base._first_lstm_input_dim_pad == 4
  assert base.FirstLstmLayerInputDimAndPadding([None, None, 4, 8]) == (32, 0)
  assert out.state == NestedMap() and not base.supports_streaming
  assert float(out.encoded[out.padding > 0].abs().max()) == 0.0       # padded frames zeroed
  assert len(base.proj) == 2 and base.output_dim == 16
  # per-layer outputs
  ex = _Encoder(extra_per_layer_outputs=True)
  o = ex.FProp(ex.theta, _EncBatch())
  assert o.conv_0.encoded.shape == (11, 2, 6, 4) and o.conv_1.encoded.shape == (6, 2, 3, 4)
  assert [o['rnn_%d' % i].encoded.shape for i in range(3)] == [(6, 2, 16)] * 3
  torch.testing.assert_close(o.rnn_2.encoded, o.encoded)
  # projection after the last layer too
  assert len(_Encoder(project_after_last_lstm=True).proj) == 3


def test_asr_encoder_residuals_highway_and_stacking():
  res = _Encoder(residual_start=2, residual_stride=1)
  out = res.FProp(res.theta, _EncBatch())
  assert out.encoded.shape == (6, 2, 16)
  hw = _Encoder(residual_start=2, highway_skip=True)
  assert len(hw.highway_skip) == 2
  assert hw.FProp(hw.theta, _EncBatch()).encoded.shape == (6, 2, 16)
  # the residual changes the function (same seed → same weights otherwise)
  plain = _Encoder()
  assert (plain.FProp(plain.theta, _EncBatch()).encoded - out.encoded).abs().max() > 1e-4
  from lingvo_b200.core import layers
  st = _Encoder(layer_index_before_stacking=1,
                stacking_layer_tpl=layers.StackingOverTime.Params().Set(
                    left_context=1, right_context=0, stride=2))
  o = st.FProp(st.theta, _EncBatch())
  assert o.encoded.shape == (3, 2, 16) and o.padding.shape == (3, 2)   # time halved after L1
  assert st.rnn[2].params.fwd.num_input_nodes == 32                     # 2 stacked frames


def test_asr_encoder_conv_lstm_blocks_and_final_proj():
  from lingvo_b200.core import layers
  enc = _Encoder(num_conv_lstm_layers=1, extra_per_layer_outputs=True,
                 final_proj=layers.ProjectionLayer.Params().Set(
                     name='final', output_dim=5, activation='NONE'))
  assert len(enc.conv_lstm_rnn) == 1
  assert tuple(enc.conv_lstm_cnn[0].params.filter_shape) == (3, 3, 8, 4)
  out = enc.FProp(enc.theta, _EncBatch())
  assert out.conv_lstm_0.encoded.shape == (6, 2, 3, 4)
  assert out.encoded.shape == (6, 2, 5) and enc.output_dim == 5
  out.encoded.sum().backward()
  assert all(v.grad is not None for v in enc.vars.Flatten() if v.requires_grad)


def test_asr_model_inference_from_wav_and_hooks(asr_records):
  import io
  import wave
  inp = input_generator.AsrInput.Params().Set(
      name='inp', file_pattern=asr_records, frame_size=80, bucket_upper_bound=[40],
      bucket_batch_limit=[4], file_buffer_size=8, file_parallelism=1, num_batcher_threads=1,
      target_max_length=16)
  inp.tokenizer = tokenizers.AsciiTokenizer.Params()
  p = asr_model.AsrModel.Params().Set(name='asr', input=inp)
  ep = p.encoder
  ep.Set(input_shape=[None, None, 80, 1], conv_filter_shapes=[(3, 3, 1, 2)],
         conv_filter_strides=[(2, 2)], num_cnn_layers=1, lstm_cell_size=8, num_lstm_layers=1,
         pad_steps=0)
  dp = p.decoder
  dp.Set(source_dim=16, emb_dim=4, rnn_cell_dim=8, rnn_layers=2, target_seq_len=5)
  dp.emb.vocab_size = 76
  dp.emb.max_num_shards = 1
  dp.attention.hidden_dim = 8
  dp.softmax.num_classes = 76
  dp.beam_search.num_hyps_per_beam = 2
  task = p.Instantiate()
  assert type(task.decoder_metrics).__name__ == 'DecoderMetrics'
  assert task.decoder_metrics.params.include_auxiliary_metrics
  # 0.3 s of a 440 Hz tone as a 16-bit mono WAV
  t = np.arange(int(0.3 * 16000)) / 16000.0
  pcm = (0.3 * 32767 * np.sin(2 * np.pi * 440 * t)).astype(np.int16)
  buf = io.BytesIO()
  with wave.open(buf, 'wb') as w:
    w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
    w.writeframes(pcm.tobytes())
  out = task.Inference()['default'](buf.getvalue())
  assert len(out.hypotheses) == 1 and len(out.hypotheses[0]) == 2
  assert all(isinstance(h, str) for h in out.hypotheses[0])
  assert out.scores.reshape(-1).shape[0] == 2
  assert out.src_frames.shape[2:] == (80, 1)
  assert out.encoder_frames.shape[1:] == (1, 16)
  # hooks
  batch = NestedMap(tgt=NestedMap(ids=torch.zeros(1, 2)), src=None)
  assert task._GetDecoderTargets(batch) is batch.tgt
  dt = task._MakeDecoderTheta(task.theta, batch)
  dt.extra = 1
  assert 'extra' not in task.theta.decoder
  ps = task.ProgramSchedule()
  assert ps.train_executions_per_eval == 0 and ps.train_program.steps_per_loop == 1000
  with pytest.raises(ValueError):
    asr_model.AsrModel.Params().Set(name='').Instantiate()
