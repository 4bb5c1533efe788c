"""Language-model fusion for the speech decoder (ref `lingvo/tasks/asr/fusion.py`).

A fusion layer owns an LM (`p.lm`), steps it alongside the acoustic decoder and combines
the two score streams. `FusionBase` implements the shared mechanics (LM state handling,
optional log-softmax of LM logits, stop-gradient into the LM); subclasses define
`FProp` (how AM outputs meet LM outputs) and `ComputeLogitsWithLM`. `NullFusion` passes
the acoustic stream through.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.lm import layers as lm_layers


class FusionBase(base_layer.BaseLayer):
  """Shared LM plumbing (ref :23)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('lm', lm_layers.NullLm.Params(), 'Language model params.')
    p.Define('base_model_logits_dim', None, 'Dim of the acoustic logits.')
    p.Define('lm_logits_dim', None, 'Dim of the LM logits (defaults to lm.vocab_size).')
    p.Define('apply_log_softmax_to_lm', True, 'Fuse LM log-probs rather than raw logits.')
    p.Define('train_lm', False, 'Back-propagate into the LM.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('lm', p.lm)

  def zero_state(self, theta, batch_size):
    """LM recurrent state + the last LM output."""
    st = NestedMap(lm_states=self.lm.zero_state(theta.lm, batch_size))
    dim = self.params.lm_logits_dim or self.params.lm.vocab_size
    dev = self.Device()
    st.lm_output = torch.zeros(batch_size, dim, device=dev)
    return st

  def _FPropLm(self, theta, state0, ids, paddings, misc=None):
    """Steps the LM over `ids [B,T]` → (new state, lm logits/log-probs [B,T,V])."""
    p = self.params
    ids_tm, pad_tm = ids.t(), paddings.t().float()
    ctx = torch.enable_grad() if (p.train_lm and torch.is_grad_enabled()) else torch.no_grad()
    with ctx:
      out, lm_state1 = self.lm.FProp(theta.lm, ids_tm, pad_tm, state0.lm_states)
      logits = out.logits if 'logits' in out else out.log_probs
      if p.apply_log_softmax_to_lm:
        logits = torch.log_softmax(logits.float(), -1)
    logits = logits.transpose(0, 1)
    if not p.train_lm:
      logits = logits.detach()
    state1 = NestedMap(lm_states=lm_state1, lm_output=logits[:, -1])
    return state1, logits

  def FProp(self, theta, state0, am_output, ids, paddings, misc=None):
    """→ (fused decoder output, new state)."""
    raise NotImplementedError()

  def ComputeLogitsWithLM(self, state, logits, is_eval=False):
    """Final fused logits given acoustic `logits` and the fusion state."""
    raise NotImplementedError()

  def AddAdditionalDecoderSummaries(self, source_encs, source_paddings, targets, seq_logits,
                                    name_suffix=''):
    return {}


class NullFusion(FusionBase):
  """Acoustic scores only (ref :173)."""

  def FProp(self, theta, state0, am_output, ids, paddings, misc=None):
    return am_output, state0

  def ComputeLogitsWithLM(self, state, logits, is_eval=False):
    return logits


class ShallowFusion(FusionBase):
  """log p_am + λ · log p_lm at every step (the classic decode-time fusion)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('lm_weight', 0.3, 'Interpolation weight λ.')
    return p

  def FProp(self, theta, state0, am_output, ids, paddings, misc=None):
    state1, _ = self._FPropLm(theta, state0, ids, paddings, misc)
    return am_output, state1

  def ComputeLogitsWithLM(self, state, logits, is_eval=False):
    return torch.log_softmax(logits.float(), -1) + self.params.lm_weight * state.lm_output
