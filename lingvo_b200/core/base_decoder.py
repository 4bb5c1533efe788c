"""Decoder base classes (ref `lingvo/core/base_decoder.py`).

`BaseDecoder` (ref :30): `FProp(theta, encoder_outputs, targets)` =
`ComputePredictions` + `ComputeLoss` → `(metrics, per_sequence)`.
`BaseBeamSearchDecoder` (ref :85) owns `beam_search` / `greedy_search` helpers and
routes their callbacks to `_InitBeamSearchStateCallback`,
`_PreBeamSearchStepCallback`, `_PostBeamSearchStepCallback`.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import beam_search_helper
from lingvo_b200.core import target_sequence_sampler
from lingvo_b200.core.nested_map import NestedMap


class BaseDecoder(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    raise NotImplementedError

  def FProp(self, theta, encoder_outputs, targets):
    predictions = self.ComputePredictions(theta, encoder_outputs, targets)
    return self.ComputeLoss(theta, predictions, targets)

  def ComputePredictions(self, theta, encoder_outputs, targets):
    raise NotImplementedError

  def ComputeLoss(self, theta, predictions, targets):
    raise NotImplementedError


LARGE_NEGATIVE_NUMBER = -1e9


def _KeepTopP(sorted_log_probs, p):
  """`[batch, k]` log-probs sorted descending → the same with everything outside the first
  `p[batch]` probability mass set to LARGE_NEGATIVE_NUMBER (the first entry always stays)
  (ref :92)."""
  cum = torch.cumsum(sorted_log_probs.exp(), -1) - sorted_log_probs.exp()    # exclusive
  mask = cum < p.unsqueeze(1)
  mask[:, 0] = True
  return torch.where(mask, sorted_log_probs,
                     torch.full_like(sorted_log_probs, LARGE_NEGATIVE_NUMBER))


def _BatchScatter(default_tensor, indices, values):
  """out[i, indices[i, j]] = values[i, j]; the rest of `default_tensor [batch, vocab]`."""
  return default_tensor.scatter(1, indices.long(), values.to(default_tensor.dtype))


def _BatchLookup(keys, table_keys, table_values):
  """`keys [batch, 1]` looked up in per-row tables → `[batch, 1]` values (first match)."""
  match = (keys == table_keys).int().argmax(1, keepdim=True)
  return table_values.gather(1, match)


def _BatchSampleGumbel(batch_seed, time_step, src_ids, src_paddings, shape, dtype):
  """Standard Gumbel noise `[batch] + shape`; row i is a pure function of (batch_seed[i] +
  Σ non-padded src_ids[i], time_step): the same sentence with the same seed draws the same
  noise whatever batch it is decoded in (ref :178)."""
  ids_sum = (src_ids * (1.0 - src_paddings).to(src_ids.dtype)).sum(1)
  seeds = (batch_seed.to(torch.int64) + ids_sum.to(torch.int64)).tolist()
  rows = []
  for seed in seeds:
    gen = torch.Generator().manual_seed((int(seed) * 1000003 + int(time_step)) % (2**63 - 1))
    u = torch.rand(list(shape), generator=gen, dtype=torch.float64).clamp(1e-12, 1 - 1e-12)
    rows.append(-torch.log(-torch.log(u)))
  return torch.stack(rows).to(dtype=dtype, device=batch_seed.device)


def _SampleGumbelWithMax(phi, target_max, batch_seed, time_step, src_ids, src_paddings):
  """Gumbel perturbations of the location parameters `phi [tgt_batch, k]` conditioned on
  their row maximum being `target_max [tgt_batch, 1]` (stochastic beam search, Kool et al.
  2019, appendix B.3 — the numerically stable form)."""
  tgt_batch, k = phi.shape
  src_batch = batch_seed.shape[0]
  n = tgt_batch // src_batch
  noise = _BatchSampleGumbel(batch_seed, time_step, src_ids, src_paddings, [n, k], phi.dtype)
  noise = noise.transpose(0, 1).reshape(tgt_batch, k).to(phi.device)   # hyp-major layout
  g_phi = phi + noise
  z = g_phi.max(1, keepdim=True).values
  v = target_max - g_phi + torch.log1p(torch.clamp(-torch.exp(g_phi - z), min=-1.0))
  return target_max - torch.relu(v) - torch.log1p(torch.exp(-v.abs()))


class BaseBeamSearchDecoder(BaseDecoder):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('target_sos_id', 1, 'SOS id.')
    p.Define('target_eos_id', 2, 'EOS id.')
    p.Define('target_seq_len', 0, 'Max target length when decoding.')
    p.Define('beam_search', beam_search_helper.BeamSearchHelper.Params(), 'Beam search.')
    p.Define('greedy_search', beam_search_helper.GreedySearchHelper.Params(), 'Greedy.')
    p.Define('target_sequence_sampler', target_sequence_sampler.TargetSequenceSampler.Params(),
             'TargetSequenceSampler params.')
    p.Define('bias_only_if_consistent', True,
             'Biased beam search stops pulling a hypothesis towards the targets once it has '
             'diverged from them.')
    p.Define('stochastic_beam_search_top_k', 8,
             'Stochastic beam search keeps (and perturbs) only the top k tokens per hyp.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    for sub in (p.beam_search, p.greedy_search):
      sub.target_seq_len = p.target_seq_len
      sub.target_sos_id = p.target_sos_id
      sub.target_eos_id = p.target_eos_id
    self.CreateChild('beam_search', p.beam_search)
    self.CreateChild('greedy_search', p.greedy_search)
    if p.target_sequence_sampler is not None:
      sp = p.target_sequence_sampler
      sp.target_seq_len = p.target_seq_len
      sp.target_sos_id = p.target_sos_id
      sp.target_eos_id = p.target_eos_id
      self.CreateChild('target_sequence_sampler', sp)

  def AddExtraDecodingInfo(self, encoder_outputs, targets):
    return encoder_outputs

  def BeamSearchDecode(self, encoder_outputs, num_hyps_per_beam_override=0):
    return self.BeamSearchDecodeWithTheta(self.theta, encoder_outputs,
                                          num_hyps_per_beam_override)

  def BeamSearchDecodeWithTheta(self, theta, encoder_outputs,
                                num_hyps_per_beam_override=0):
    return self.beam_search.BeamSearchDecode(
        theta, encoder_outputs, num_hyps_per_beam_override,
        self._InitBeamSearchStateCallback, self._PreBeamSearchStepCallback,
        self._PostBeamSearchStepCallback)

  def GreedySearchDecode(self, encoder_outputs):
    return self.GreedySearchDecodeWithTheta(self.theta, encoder_outputs)

  def GreedySearchDecodeWithTheta(self, theta, encoder_outputs):
    return self.greedy_search.GreedySearchDecode(
        theta, encoder_outputs, self._InitBeamSearchStateCallback,
        self._PreBeamSearchStepCallback, self._PostBeamSearchStepCallback)

  # -- sampling ------------------------------------------------------------------------------
  def SampleTargetSequences(self, theta, encoder_outputs, random_seed):
    """→ NestedMap(ids `[batch, T]`, paddings, logits) drawn token by token from the model
    (top-k / nucleus / temperature per the sampler params) (ref :438)."""
    return self.target_sequence_sampler.Sample(
        theta, encoder_outputs, random_seed, self._InitBeamSearchStateCallback,
        self._PreBeamSearchStepCallback, self._PostBeamSearchStepCallback)

  def _PostprocessSample(self, sample, is_tpu=False):
    """Adds the `BeamSearchDecodeOutput` fields to a sample (ref :387): `topk_ids / lens /
    scores` regrouped from hyp-major `[n · batch]` to source-major `[batch · n]` order, the
    score being the sample's total log-probability."""
    del is_tpu
    p = self.params
    n = p.target_sequence_sampler.num_hyps_per_beam
    bs, max_len = sample.ids.shape
    weights = 1.0 - sample.paddings.float()
    logp = torch.log_softmax(sample.logits.float(), -1)
    tok = logp.gather(-1, sample.ids.long().unsqueeze(-1)).squeeze(-1)
    scores = (tok * weights).sum(1)
    regroup = lambda t: t.reshape(n, bs // n, *t.shape[1:]).transpose(0, 1).reshape(t.shape)
    sample.topk_hyps = None
    sample.topk_ids = regroup(sample.ids)
    sample.topk_lens = regroup(weights.sum(1).to(torch.int32))
    sample.topk_scores = regroup(scores)
    return sample

  def SampleSequenceDecode(self, encoder_outputs, random_seed=None):
    """Decode by sampling; same output fields as `BeamSearchDecode`."""
    import random as _random  # pylint: disable=g-import-not-at-top
    seed = self.params.random_seed if random_seed is None else random_seed
    if seed is None:
      seed = _random.randrange(2**31 - 1)
    return self._PostprocessSample(
        self.SampleTargetSequences(self.theta, encoder_outputs, seed))

  # -- biased / stochastic beam search ------------------------------------------------------------
  def BeamSearchDecodeBiased(self, encoder_outputs, num_hyps_per_beam_override=0):
    """Beam search pulled towards `encoder_outputs.targets` (labels, paddings, weights
    `[batch, seq]`; weight 1 = forced decoding) (ref :460)."""
    return self.StochasticBeamSearchDecodeBiased(
        encoder_outputs, biased=True, stochastic=False,
        num_hyps_per_beam_override=num_hyps_per_beam_override)

  def StochasticBeamSearchDecodeBiased(self, encoder_outputs, biased, stochastic,
                                       num_hyps_per_beam_override=0):
    """Beam search with target biasing and / or stochastic beam search (sampling without
    replacement by Gumbel-perturbed scores + top-p filtering) (ref :481). `stochastic` reads
    `encoder_outputs.stochastic_beam_search` = NestedMap(top_p_threshold [batch], seed
    [batch], src_ids, src_paddings)."""
    p = self.params
    if biased:
      targets = encoder_outputs.targets
      targets.weights = targets.weights * (1.0 - targets.paddings)
      pad = max(0, p.beam_search.target_seq_len - targets.labels.shape[1])
      targets.labels = torch.nn.functional.pad(targets.labels, (0, pad))
      targets.weights = torch.nn.functional.pad(targets.weights, (0, pad))
    if stochastic:
      sbs = encoder_outputs.stochastic_beam_search
      sbs.enable = bool((sbs.top_p_threshold > 0).any())
    return self.beam_search.BeamSearchDecode(
        self.theta, encoder_outputs, num_hyps_per_beam_override,
        self._WrapInitBeamSearchStateCallback(biased, stochastic),
        self._WrapPreBeamSearchStepCallback(biased, stochastic),
        self._WrapPostBeamSearchStepCallback(stochastic))

  def _WrapInitBeamSearchStateCallback(self, biased, stochastic):
    k = self.params.stochastic_beam_search_top_k

    def Callback(theta, encoder_outputs, num_hyps_per_beam):
      results, states = self._InitBeamSearchStateCallback(theta, encoder_outputs,
                                                          num_hyps_per_beam)
      n = results.log_probs.shape[0]
      dev = results.log_probs.device
      if 'time_step' not in states:
        states.time_step = torch.zeros((), dtype=torch.int64, device=dev)
      if biased:
        states.consistent = torch.ones(n, dtype=torch.bool, device=dev)
      if stochastic:
        states.cumulative_log_probs = torch.zeros(n, 1, device=dev)
        states.perturbed_cumulative_log_probs = torch.zeros(n, 1, device=dev)
        states.tmp_states = NestedMap(
            top_k_log_probs=torch.zeros(n, k, device=dev),
            top_k_ids=torch.zeros(n, k, dtype=torch.int64, device=dev),
            new_perturbed_cumulative_log_probs=torch.zeros(n, k, device=dev))
      return results, states

    return Callback

  def _WrapPreBeamSearchStepCallback(self, biased, stochastic):
    k = self.params.stochastic_beam_search_top_k

    def Callback(theta, encoder_outputs, step_ids, states, num_hyps_per_beam, cur_step,
                 *args, **kwargs):
      p = self.params
      carried = {key: states.get(key) for key in (
          'consistent', 'cumulative_log_probs', 'perturbed_cumulative_log_probs',
          'tmp_states')}
      results, out_states = self._PreBeamSearchStepCallback(
          theta, encoder_outputs, step_ids, states, num_hyps_per_beam, cur_step, *args,
          **kwargs)
      t = int(cur_step)
      tgt_batch = step_ids.shape[0]

      def Tile(x):                      # [src_batch] → [n · src_batch], hyp-major
        return x.reshape(1, -1).repeat(num_hyps_per_beam, 1).reshape(tgt_batch)

      if biased:
        labels, weights = encoder_outputs.targets.labels, encoder_outputs.targets.weights
        consistent = carried['consistent']
        if bool((weights != 0).any()) and t < labels.shape[1]:
          prev = Tile(labels[:, max(t - 1, 0)])
          local = torch.ones_like(consistent) if t == 0 else prev == step_ids.squeeze(1)
          consistent = consistent & local
          label, weight = Tile(labels[:, t]), Tile(weights[:, t]).float()
          if p.bias_only_if_consistent:
            weight = weight * consistent.float()
          assert bool((weight <= 1.0).all()) and bool((weight >= 0.0).all())
          vocab = results.log_probs.shape[1]
          label_probs = torch.nn.functional.one_hot(label.long(), vocab).float()
          w = weight.unsqueeze(1)
          probs = (1.0 - w) * results.log_probs.float().exp() + w * label_probs
          results.log_probs = probs.clamp_min(1e-12).log()
        out_states.consistent = consistent
      if stochastic:
        sbs = encoder_outputs.stochastic_beam_search
        out_states.tmp_states = carried['tmp_states']
        if sbs.enable:
          top = torch.topk(results.log_probs.float(), k, dim=-1, sorted=True)
          thr = Tile(sbs.top_p_threshold.clamp(0.0, 1.0))
          filtered = _KeepTopP(top.values, thr)
          last_pert = carried['perturbed_cumulative_log_probs']
          cum = carried['cumulative_log_probs'] + filtered
          new_pert = _SampleGumbelWithMax(cum, last_pert, sbs.seed, t, sbs.src_ids,
                                          sbs.src_paddings)
          updated = torch.full_like(results.log_probs.float(), LARGE_NEGATIVE_NUMBER)
          results.log_probs = _BatchScatter(updated, top.indices, new_pert - last_pert)
          out_states.tmp_states = NestedMap(
              new_perturbed_cumulative_log_probs=new_pert, top_k_log_probs=top.values,
              top_k_ids=top.indices)
        out_states.cumulative_log_probs = carried['cumulative_log_probs']
        out_states.perturbed_cumulative_log_probs = carried['perturbed_cumulative_log_probs']
      if 'time_step' not in out_states:
        out_states.time_step = torch.as_tensor(t + 1)
      return results, out_states

    return Callback

  def _WrapPostBeamSearchStepCallback(self, stochastic):

    def Callback(theta, encoder_outputs, new_step_ids, other_states):
      final = self._PostBeamSearchStepCallback(theta, encoder_outputs, new_step_ids,
                                               other_states)
      if stochastic and encoder_outputs.stochastic_beam_search.enable:
        tmp = other_states.tmp_states
        ids = new_step_ids.reshape(-1, 1)
        final.perturbed_cumulative_log_probs = _BatchLookup(
            ids, tmp.top_k_ids, tmp.new_perturbed_cumulative_log_probs)
        final.cumulative_log_probs = other_states.cumulative_log_probs + _BatchLookup(
            ids, tmp.top_k_ids, tmp.top_k_log_probs)
      return final

    return Callback

  def InferenceAdditionalEncoder(self, feeds):
    """Hook: (fetches, feeds) of an additional encoder in the inference graph."""
    del feeds
    return NestedMap(), NestedMap()

  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    raise NotImplementedError

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    raise NotImplementedError

  def _PostBeamSearchStepCallback(self, theta, encoder_outputs, new_step_ids, states):
    return states
