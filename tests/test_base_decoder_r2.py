"""Sampling, biased and stochastic beam search of BaseBeamSearchDecoder
(ref lingvo/core/base_decoder_test.py)."""
import numpy as np
import pytest
import torch

from lingvo_b200.core import base_decoder
from lingvo_b200.core.nested_map import NestedMap

VOCAB = 8


class _ToyDecoder(base_decoder.BaseBeamSearchDecoder):
  """A fixed bigram model: log p(next | prev) = table[prev]; no encoder dependence."""

  def __init__(self, params):
    super().__init__(params)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(VOCAB, VOCAB, generator=g) * 2.0
    logits[:, 0] = -20.0                      # never emit the pad id
    logits[:, 1] = -20.0                      # nor sos
    self.table = torch.log_softmax(logits, -1)

  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    n = encoder_outputs.padding.shape[1] * num_hyps_per_beam
    return (NestedMap(log_probs=torch.zeros(n, VOCAB), atten_probs=torch.zeros(n, 3)),
            NestedMap(time_step=torch.zeros((), dtype=torch.int64)))

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    n = step_ids.shape[0]
    return (NestedMap(log_probs=self.table[step_ids.squeeze(1)].clone(),
                      atten_probs=torch.full((n, 3), 1 / 3.0)),
            NestedMap(time_step=states.time_step + 1))


def _Decoder(**kw):
  p = _ToyDecoder.Params().Set(name='toy', target_seq_len=6, **kw)
  p.beam_search.num_hyps_per_beam = 3
  p.beam_search.length_normalization = 0.0
  p.beam_search.coverage_penalty = 0.0
  p.target_sequence_sampler.num_hyps_per_beam = 2
  return p.Instantiate()


def _Enc(b=2):
  return NestedMap(encoded=torch.zeros(3, b, 4), padding=torch.zeros(3, b))


def test_top_p_scatter_lookup_helpers():
  lp = torch.log(torch.tensor([[0.5, 0.3, 0.15, 0.05], [0.9, 0.05, 0.03, 0.02]]))
  kept = base_decoder._KeepTopP(lp, torch.tensor([0.7, 0.0]))
  assert (kept[0] > -1e8).tolist() == [True, True, False, False]      # cum before 3rd = .8 ≥ .7
  assert (kept[1] > -1e8).tolist() == [True, False, False, False]     # first always kept
  out = base_decoder._BatchScatter(torch.zeros(2, 5), torch.tensor([[1, 3], [0, 4]]),
                                   torch.tensor([[1., 2.], [3., 4.]]))
  assert out.tolist() == [[0, 1, 0, 2, 0], [3, 0, 0, 0, 4]]
  got = base_decoder._BatchLookup(torch.tensor([[3], [0]]), torch.tensor([[1, 3], [0, 4]]),
                                  torch.tensor([[1., 2.], [3., 4.]]))
  assert got.tolist() == [[2.0], [3.0]]


def test_gumbel_with_max_hits_the_target_and_is_reproducible():
  phi = torch.randn(6, 5)
  tmax = torch.randn(6, 1) - 3.0
  seed = torch.tensor([11, 12, 13])
  src = torch.randint(2, 9, (3, 4)); pad = torch.zeros(3, 4)
  a = base_decoder._SampleGumbelWithMax(phi, tmax, seed, 2, src, pad)
  b = base_decoder._SampleGumbelWithMax(phi, tmax, seed, 2, src, pad)
  torch.testing.assert_close(a, b)
  torch.testing.assert_close(a.max(1, keepdim=True).values, tmax, atol=1e-4, rtol=1e-4)
  c = base_decoder._SampleGumbelWithMax(phi, tmax, seed, 3, src, pad)
  assert (a - c).abs().max() > 1e-3                                   # time step matters
  # the noise of a sentence does not depend on which batch it sits in
  n1 = base_decoder._BatchSampleGumbel(seed, 1, src, pad, [2, 5], torch.float32)
  n2 = base_decoder._BatchSampleGumbel(seed[1:], 1, src[1:], pad[1:], [2, 5], torch.float32)
  torch.testing.assert_close(n1[1:], n2)


def test_sampling_decode_output_layout_and_scores():
  dec = _Decoder(random_seed=5)
  enc = _Enc()
  sample = dec.SampleTargetSequences(dec.theta, enc, 123)
  assert sample.ids.shape == (4, 6) and sample.logits.shape == (4, 6, VOCAB)
  out = dec._PostprocessSample(NestedMap(sample))
  # hyp-major [n·b] → source-major [b·n]
  assert out.topk_ids[1].tolist() == sample.ids[2].tolist()
  assert out.topk_ids[2].tolist() == sample.ids[1].tolist()
  w = 1.0 - sample.paddings
  lp = torch.log_softmax(sample.logits, -1).gather(-1, sample.ids.unsqueeze(-1)).squeeze(-1)
  torch.testing.assert_close(out.topk_scores[1], (lp * w).sum(1)[2])
  assert out.topk_lens.tolist() == [int(w[i].sum()) for i in (0, 2, 1, 3)]
  again = dec.SampleSequenceDecode(enc)
  again2 = dec.SampleSequenceDecode(enc)
  assert torch.equal(again.topk_ids, again2.topk_ids)                 # p.random_seed fixed
  greedy = dec.GreedySearchDecodeWithTheta(dec.theta, enc)
  assert greedy is not None


def test_biased_beam_search_forces_the_target_prefix():
  dec = _Decoder()
  enc = _Enc()
  plain = dec.BeamSearchDecode(enc)
  target = torch.tensor([[5, 6, 7, 2], [3, 3, 4, 2]])
  enc.targets = NestedMap(labels=target.clone(), paddings=torch.zeros(2, 4),
                          weights=torch.ones(2, 4))
  forced = dec.BeamSearchDecodeBiased(enc)
  ids = forced.topk_hyps.ids                       # [b, k, T]
  assert ids[0, 0, :4].tolist() == [5, 6, 7, 2] and ids[1, 0, :4].tolist() == [3, 3, 4, 2]
  assert int(forced.topk_hyps.lens[0, 0]) == 4
  assert not torch.equal(plain.topk_hyps.ids[:, 0, :4], ids[:, 0, :4])
  # zero weights: identical to the unbiased search
  enc.targets = NestedMap(labels=target.clone(), paddings=torch.zeros(2, 4),
                          weights=torch.zeros(2, 4))
  free = dec.BeamSearchDecodeBiased(enc)
  assert torch.equal(free.topk_hyps.ids, plain.topk_hyps.ids)
  # partial weights pull the first token only
  enc.targets = NestedMap(labels=target.clone(), paddings=torch.zeros(2, 4),
                          weights=torch.tensor([[1.0, 0, 0, 0], [1.0, 0, 0, 0]]))
  part = dec.BeamSearchDecodeBiased(enc)
  assert part.topk_hyps.ids[:, 0, 0].tolist() == [5, 3]


def test_stochastic_beam_search_samples_without_replacement():
  dec = _Decoder()

  def Run(seed, top_p=1.0):
    enc = _Enc()
    enc.stochastic_beam_search = NestedMap(
        top_p_threshold=torch.full((2,), top_p), seed=torch.tensor([seed, seed + 1]),
        src_ids=torch.tensor([[4, 5, 6], [7, 7, 2]]), src_paddings=torch.zeros(2, 3))
    return dec.StochasticBeamSearchDecodeBiased(enc, biased=False, stochastic=True)

  a, b, c = Run(1), Run(1), Run(2)
  assert torch.equal(a.topk_hyps.ids, b.topk_hyps.ids)               # same seed, same samples
  assert not torch.equal(a.topk_hyps.ids, c.topk_hyps.ids)
  for bi in range(2):                                                 # hyps of a beam are distinct
    rows = {tuple(r.tolist()) for r in a.topk_hyps.ids[bi]}
    assert len(rows) == 3
  off = Run(1, top_p=0.0)                                             # disabled → plain beam search
  plain = dec.BeamSearchDecode(_Enc())
  assert torch.equal(off.topk_hyps.ids, plain.topk_hyps.ids)
  assert dec.InferenceAdditionalEncoder(None) == (NestedMap(), NestedMap())
