"""MTDecoderV1 / MTBaseDecoder features added in round 2 (ref tasks/mt/decoder_test.py)."""
import pytest
import torch

from lingvo_b200.core import layers
from lingvo_b200.core import quant_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.mt import decoder as mt_decoder

V = 12


def _Decoder(**kw):
  sm = kw.pop('softmax', None)
  p = mt_decoder.MTDecoderV1.Params().Set(
      name='dec', source_dim=6, rnn_cell_dim=8, rnn_layers=3, residual_start=2,
      target_seq_len=6, random_seed=17, **kw)
  p.emb.Set(vocab_size=V, embedding_dim=8)
  if 'max_num_shards' in p.emb:
    p.emb.max_num_shards = 1
  p.attention.hidden_dim = 5
  p.softmax.num_classes = V
  p.softmax.num_shards = 1
  if sm is not None:
    p.softmax = sm
  p.beam_search.num_hyps_per_beam = 2
  p.beam_search.length_normalization = 0.0
  p.beam_search.coverage_penalty = 0.0
  dec = p.Instantiate()
  dec.InstantiateVariables()
  return dec


def _Inputs(b=3, t=5, s=4, seed=2):
  g = torch.Generator().manual_seed(seed)
  enc = NestedMap(encoded=torch.randn(s, b, 6, generator=g), padding=torch.zeros(s, b))
  ids = torch.randint(3, V, (b, t), generator=g)
  pad = torch.zeros(b, t)
  pad[0, 3:] = 1
  tgt = NestedMap(ids=ids, labels=torch.roll(ids, -1, 1), paddings=pad, weights=1 - pad)
  return enc, tgt


def test_predictions_loss_and_per_example_tensors():
  dec = _Decoder(per_example_tensors=True)
  enc, tgt = _Inputs()
  pred = dec.ComputePredictions(dec.theta, enc, tgt)
  assert pred.softmax_input.shape == (5, 3, 8) and pred.attention.probs.shape == (3, 5, 4)
  assert pred.source_enc_len.tolist() == [4, 4, 4]
  m, per = dec.ComputeLoss(dec.theta, pred, tgt)
  # per-sentence average: Σ_t w·xent per sequence, mean over the batch
  torch.testing.assert_close(m.loss[0], per.per_sequence_loss.mean())
  assert float(m.loss[1]) == 3.0 and per.per_example_loss.shape == (5, 3)
  assert per.logits.shape == (5, 3, V)
  word = _Decoder(per_word_avg_loss=True)
  mw, _ = word.ComputeLoss(word.theta, word.ComputePredictions(word.theta, enc, tgt), tgt)
  torch.testing.assert_close(mw.loss[0], mw.log_pplx[0])
  assert float(mw.loss[1]) == float(tgt.weights.sum())
  cut = dec._TruncateTargetSequence(NestedMap(
      ids=tgt.ids, labels=tgt.labels, weights=tgt.weights,
      paddings=torch.cat([tgt.paddings[:, :3], torch.ones(3, 2)], 1)))
  assert cut.ids.shape == (3, 3)
  assert dec._ExpandToNumHyps(torch.tensor([3, 2, 1]), 2).tolist() == [3, 2, 1, 3, 2, 1]


def test_context_to_softmax_prev_ctx_and_zero_first_step():
  enc, tgt = _Inputs()
  feed = _Decoder(feed_attention_context_vec_to_softmax=True)
  assert feed.ComputePredictions(feed.theta, enc, tgt).softmax_input.shape == (5, 3, 14)
  base = _Decoder()
  prev = _Decoder(use_prev_atten_ctx=True, use_zero_atten_state=True)
  a = base.ComputePredictions(base.theta, enc, tgt).softmax_input
  b = prev.ComputePredictions(prev.theta, enc, tgt).softmax_input
  assert a.shape == b.shape and (a - b).abs().max() > 1e-4
  zero = _Decoder(zero_token_embs_first_time_step=True)
  tgt2 = NestedMap(tgt)
  tgt2.ids = tgt.ids.clone(); tgt2.ids[:, 0] = 5           # the first id must not matter
  z1 = zero.ComputePredictions(zero.theta, enc, tgt).softmax_input
  z2 = zero.ComputePredictions(zero.theta, enc, tgt2).softmax_input
  torch.testing.assert_close(z1, z2)
  e = torch.ones(4, 2, 3)
  assert zero._ZeroOutFirstTimeStep(e)[0].abs().sum() == 0 and zero._ZeroOutFirstTimeStep(e)[1:].sum() == 18


def test_beam_search_step_plan_matches_training_logits():
  dec = _Decoder()
  enc, tgt = _Inputs()
  tgt.paddings = torch.zeros_like(tgt.paddings)
  pred = dec.ComputePredictions(dec.theta, enc, tgt)
  want = torch.log_softmax(dec.softmax.Logits(dec.theta.softmax, pred.softmax_input), -1)
  init, states = dec._InitBeamSearchStateCallback(dec.theta, enc, 1)
  for t in range(5):
    res, states = dec._PreBeamSearchStepCallback(dec.theta, enc, tgt.ids[:, t:t + 1], states, 1, t)
    torch.testing.assert_close(res.log_probs, want[t], atol=1e-4, rtol=1e-4)
  assert int(states.time_step) == 5
  out = dec.BeamSearchDecode(enc)
  assert out.topk_hyps.ids.shape[:2] == (3, 2)


def test_force_alignment_and_single_token_fast_decode():
  with pytest.raises(ValueError):
    _Decoder(force_alignment=True)
  dec = _Decoder(force_alignment=True, sentence_boundary_token_id=7)
  lp = torch.log_softmax(torch.randn(4, V), -1)
  src = torch.tensor([2, 2, 1, 1])
  hyp = torch.tensor([1, 2, 1, 2])
  out = dec._ForceAlignment(lp, src, hyp)
  assert float(out[0, 2]) < -1e30 and float(out[0, 7]) == float(lp[0, 7])   # needs a boundary
  assert float(out[1, 7]) < -1e30 and float(out[1, 2]) == float(lp[1, 2])   # may finish
  assert float(out[3, 7]) < -1e30
  keep = [i for i in range(V) if i not in (2, 7)]
  torch.testing.assert_close(out[:, keep], lp[:, keep])
  enc, _ = _Inputs(b=2)
  enc.num_sentences = torch.tensor([2, 1])
  init, states = dec._InitBeamSearchStateCallback(dec.theta, enc, 2)
  assert states.num_sentences.tolist() == [1, 1, 1, 1]
  step_ids = torch.full((4, 1), 1)
  res, states = dec._PreBeamSearchStepCallback(dec.theta, enc, step_ids, states, 2, 0)
  # source 0 has two sentences: EOS is blocked, the boundary allowed; source 1 the reverse
  assert (res.log_probs[[0, 2], 2] < -1e30).all() and (res.log_probs[[0, 2], 7] > -1e30).all()
  assert (res.log_probs[[1, 3], 7] < -1e30).all() and (res.log_probs[[1, 3], 2] > -1e30).all()
  states = dec._PostBeamSearchStepCallback(dec.theta, enc, torch.tensor([[7], [3], [4], [7]]),
                                           states)
  assert states.num_sentences.tolist() == [2, 1, 1, 2]
  res, _ = dec._PreBeamSearchStepCallback(dec.theta, enc, step_ids, states, 2, 1)
  assert float(res.log_probs[0, 2]) > -1e30 and float(res.log_probs[0, 7]) < -1e30
  no_key, _ = _Inputs(b=2)
  with pytest.raises(ValueError):
    dec.BeamSearchDecode(no_key)
  fast = _Decoder(single_token_fast_decode=True)
  enc1, _ = _Inputs(b=2)
  enc1.padding[1:, 0] = 1.0                               # source 0 has one token
  r = fast.BeamSearchDecode(enc1)
  assert int(r.topk_hyps.lens[0, 0]) == 1 and int(r.topk_hyps.ids[0, 0, 0]) == 2
  upd = fast._UpdateLogitsForSingleTokenFastDecode(lp, torch.tensor([True, False]), 2)
  assert float(upd[0, 2]) == 0.0 and float(upd[2, 2]) == 0.0 and torch.equal(upd[1], lp[1])


def test_init_step_ids_and_sigmoid_scores():
  dec = _Decoder(init_step_ids=True, use_sigmoid_activation=True)
  enc, tgt = _Inputs()
  enc = dec.AddExtraDecodingInfo(enc, tgt)
  assert torch.equal(enc.init_step_ids, tgt.ids[:, 0])
  init, states = dec._InitBeamSearchStateCallback(dec.theta, enc, 2)
  assert init.step_ids.reshape(-1).tolist() == tgt.ids[:, 0].tolist() * 2
  res, _ = dec._PreBeamSearchStepCallback(dec.theta, enc, init.step_ids, states, 2, 0)
  assert float(res.log_probs.exp().sum(-1).max()) > 1.5      # sigmoids do not sum to one
  assert dec.BeamSearchDecode(enc).topk_hyps.ids.shape[0] == 3


def test_shared_softmax_embedding_projection_and_clipping_cap():
  shared = layers.SharedSoftmaxLayer.Params().Set(name='softmax', vocab_size=V, embedding_dim=4)
  p = mt_decoder.MTDecoderV1.Params().Set(
      name='dec', source_dim=6, rnn_cell_dim=8, rnn_layers=2, target_seq_len=4, random_seed=3,
      emb_projection_tpl=layers.ProjectionLayer.Params().Set(batch_norm=False,
                                                             activation='NONE'))
  p.attention.hidden_dim = 5
  p.softmax = shared
  dec = p.Instantiate()
  dec.InstantiateVariables()
  assert 'emb' not in dec.children and {'emb_proj', 'out_proj'} <= set(dec.children)
  enc, tgt = _Inputs()
  pred = dec.ComputePredictions(dec.theta, enc, tgt)
  assert pred.softmax_input.shape == (5, 3, 4)              # projected down to the emb dim
  m, _ = dec.ComputeLoss(dec.theta, pred, tgt)
  m.loss[0].backward()
  assert dec.vars.softmax.Flatten()[0].grad is not None
  cc = _Decoder(cc_schedule=quant_utils.LinearClippingCapSchedule.Params().Set(
      start_step=0, end_step=1, start_cap=0.05, end_cap=0.05))
  pc = cc.ComputePredictions(cc.theta, enc, tgt)
  assert float(pc.softmax_input.abs().max()) <= 0.05 + 1e-6
  assert float(cc.ApplyClipping(cc.theta, torch.tensor([3.0]))) == pytest.approx(0.05)
  assert float(_Decoder().ApplyClipping(NestedMap(), torch.tensor([3.0]))) == 3.0


def test_packed_inputs_normalise_by_sentence_count():
  dec = _Decoder(packed_input=True)
  enc, tgt = _Inputs()
  enc.segment_id = torch.ones(4, 3)
  tgt.paddings = torch.zeros(3, 5); tgt.weights = torch.ones(3, 5)
  tgt.segment_ids = torch.tensor([[1, 1, 1, 2, 2], [1, 1, 1, 1, 1], [1, 1, 2, 2, 2]]).float()
  tgt.segment_pos = torch.tensor([[0, 1, 2, 0, 1], [0, 1, 2, 3, 4], [0, 1, 0, 1, 2]])
  pred = dec.ComputePredictions(dec.theta, enc, tgt)
  m, per = dec.ComputeLoss(dec.theta, pred, tgt)
  torch.testing.assert_close(m.loss[0], per.per_sequence_xent.sum() / 5.0)   # 2 + 1 + 2 sentences
  tgt.pop('segment_ids')
  with pytest.raises((AssertionError, AttributeError, KeyError)):
    dec.ComputeLoss(dec.theta, pred, tgt)


def _TransformerDecoder(**kw):
  p = mt_decoder.TransformerDecoder.Params().Set(
      name='tdec', source_dim=8, model_dim=8, num_trans_layers=2, hidden_dim=16,
      num_atten_heads=2, target_seq_len=6, random_seed=5, **kw)
  p.token_emb.vocab_size = V
  if 'max_num_shards' in p.token_emb:
    p.token_emb.max_num_shards = 1
  p.softmax.num_classes = V
  p.softmax.num_shards = 1
  p.beam_search.num_hyps_per_beam = 2
  dec = p.Instantiate()
  dec.InstantiateVariables()
  return dec


def _TInputs(b=2, t=5, s=4, layers_dim=None):
  g = torch.Generator().manual_seed(4)
  shape = (s, b, 8) if layers_dim is None else (s, b, 8, layers_dim)
  enc = NestedMap(encoded=torch.randn(*shape, generator=g), padding=torch.zeros(s, b))
  ids = torch.randint(3, V, (b, t), generator=g)
  tgt = NestedMap(ids=ids, labels=torch.roll(ids, -1, 1), paddings=torch.zeros(b, t),
                  weights=torch.ones(b, t))
  return enc, tgt


def test_transformer_decoder_extend_step_matches_full_pass():
  dec = _TransformerDecoder()
  enc, tgt = _TInputs()
  full = dec.ComputePredictions(dec.theta, enc, tgt).softmax_input           # [T, B, D]
  states = dec.InitPrefixStates(dec.theta, 2, 5)
  for t in range(5):
    step, states = dec.ExtendStep(dec.theta, enc, tgt.ids[:, t], t, states)
    torch.testing.assert_close(step, full[t], atol=2e-4, rtol=2e-4)
  assert dec.BeamSearchDecode(enc).topk_hyps.ids.shape[:2] == (2, 2)


def test_transformer_decoder_transparent_task_emb_and_zero_first_step():
  tr = _TransformerDecoder(is_transparent=True)
  enc, tgt = _TInputs(layers_dim=2)
  out = tr.ComputePredictions(tr.theta, enc, tgt).softmax_input
  # layer 1 reads slice 1 only: perturbing slice 0 must still change the output (layer 0),
  # and both slices matter
  e0 = NestedMap(enc); e0.encoded = enc.encoded.clone(); e0.encoded[..., 0] += 1.0
  e1 = NestedMap(enc); e1.encoded = enc.encoded.clone(); e1.encoded[..., 1] += 1.0
  assert (tr.ComputePredictions(tr.theta, e0, tgt).softmax_input - out).abs().max() > 1e-4
  assert (tr.ComputePredictions(tr.theta, e1, tgt).softmax_input - out).abs().max() > 1e-4
  assert tr.BeamSearchDecode(enc).topk_hyps.ids.shape[0] == 2
  task = _TransformerDecoder(
      task_emb=layers.SimpleEmbeddingLayer.Params().Set(vocab_size=3, embedding_dim=8),
      init_step_ids=True, zero_token_embs_first_time_step=True,
      ln_input=layers.LayerNorm.Params())
  enc, tgt = _TInputs()
  tgt.task_ids = torch.tensor([[1] * 5, [2] * 5])
  a = task.ComputePredictions(task.theta, enc, tgt).softmax_input
  tgt2 = NestedMap(tgt); tgt2.task_ids = torch.zeros(2, 5, dtype=torch.long)
  assert (task.ComputePredictions(task.theta, enc, tgt2).softmax_input - a).abs().max() > 1e-4
  tgt3 = NestedMap(tgt); tgt3.ids = tgt.ids.clone(); tgt3.ids[:, 0] = 9
  torch.testing.assert_close(task.ComputePredictions(task.theta, enc, tgt3).softmax_input, a)
  enc = task.AddExtraDecodingInfo(enc, tgt)
  assert enc.target_task_ids.tolist() == [1, 2] and torch.equal(enc.init_step_ids, tgt.ids[:, 0])
  init, _ = task._InitBeamSearchStateCallback(task.theta, enc, 2)
  assert init.step_ids.reshape(-1).tolist() == tgt.ids[:, 0].tolist() * 2
  assert task.BeamSearchDecode(enc).topk_hyps.ids.shape[0] == 2
