"""Decoder base classes (ref `lingvo/core/base_decoder.py`).

`BaseDecoder` (ref :30): `FProp(theta, encoder_outputs, targets)` =
`ComputePredictions` + `ComputeLoss` → `(metrics, per_sequence)`.
`BaseBeamSearchDecoder` (ref :85) owns `beam_search` / `greedy_search` helpers and
routes their callbacks to `_InitBeamSearchStateCallback`,
`_PreBeamSearchStepCallback`, `_PostBeamSearchStepCallback`.
"""

from __future__ import annotations

from lingvo_b200.core import base_layer
from lingvo_b200.core import beam_search_helper
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


class BaseBeamSearchDecoder(BaseDecoder):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('target_sos_id', 1, 'SOS id.')
    p.Define('target_eos_id', 2, 'EOS id.')
    p.Define('target_seq_len', 0, 'Max target length when decoding.')
    p.Define('beam_search', beam_search_helper.BeamSearchHelper.Params(), 'Beam search.')
    p.Define('greedy_search', beam_search_helper.GreedySearchHelper.Params(), 'Greedy.')
    p.Define('target_sequence_sampler', None, 'Sampler params.')
    p.Define('bias_only_if_consistent', True, 'Kept for parity.')
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
    return self.greedy_search.GreedySearchDecode(
        self.theta, encoder_outputs, self._InitBeamSearchStateCallback,
        self._PreBeamSearchStepCallback, self._PostBeamSearchStepCallback)

  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    raise NotImplementedError

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    raise NotImplementedError

  def _PostBeamSearchStepCallback(self, theta, encoder_outputs, new_step_ids, states):
    return states
