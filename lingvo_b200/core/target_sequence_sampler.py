"""Sampling decoder (ref `lingvo/core/target_sequence_sampler.py:35`): draws target
sequences token by token through the beam-search callbacks with temperature, top-k,
nucleus (top-p) and ε filtering. All filtering is batched tensor code on the device."""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap


def _ComputePaddings(ids, eos_id):
  """1 after the first EOS (the EOS itself is not padding) (ref :27)."""
  is_eos = (ids == eos_id).to(torch.int32)
  after = torch.cumsum(is_eos, 1) - is_eos
  return (after > 0).float()


class TargetSequenceSampler(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('target_sos_id', 1, 'SOS id.')
    p.Define('target_eos_id', 2, 'EOS id.')
    p.Define('target_eoc_id', -1, 'End-of-chunk id.')
    p.Define('target_seq_len', 0, 'Max target length.')
    p.Define('top_k', 0, 'Top-k sampling if > 0.')
    p.Define('top_k_renormalize', True, 'Renormalise top-k probabilities.')
    p.Define('nucleus_p', 1.0, 'Nucleus sampling if < 1.')
    p.Define('epsilon', 0.0, 'Mask tokens with probability < epsilon.')
    p.Define('eps_fail_safe', True, 'Fall back to top-1 when ε exceeds the max probability.')
    p.Define('temperature', 1.0, 'Softmax temperature.')
    p.Define('use_stop_fn', False, 'Stop when every sample has emitted EOS.')
    p.Define('num_hyps_per_beam', 1, 'Samples per source.')
    p.Define('use_recurrent', False, 'Kept for parity.')
    p.name = 'target_sequence_sampler'
    return p

  def _Filter(self, logits):
    p = self.params
    logits = logits.float() / p.temperature
    if p.top_k > 0:
      kth = torch.topk(logits, min(p.top_k, logits.shape[-1]), -1).values[..., -1:]
      logits = logits.masked_fill(logits < kth, -1e30)
    if p.nucleus_p < 1.0:
      srt, idx = torch.sort(logits, -1, descending=True)
      cum = torch.softmax(srt, -1).cumsum(-1)
      drop = cum - torch.softmax(srt, -1) >= p.nucleus_p     # keep the token crossing p
      drop_orig = torch.zeros_like(drop).scatter(-1, idx, drop)
      logits = logits.masked_fill(drop_orig, -1e30)
    if p.epsilon > 0:
      probs = torch.softmax(logits, -1)
      low = probs < p.epsilon
      if p.eps_fail_safe:
        top1 = probs >= probs.max(-1, keepdim=True).values
        low = low & ~top1
      logits = logits.masked_fill(low, -1e30)
    return logits

  def Sample(self, decoder_theta, encoder_outputs, random_seed, init_state_callback,
             pre_step_callback, post_step_callback, init_step_ids=None):
    """→ NestedMap(logits [B,T,V], ids [B,T], paddings [B,T])."""
    p = self.params
    assert p.temperature > 0 and p.target_seq_len > 0
    res, state = init_state_callback(decoder_theta, encoder_outputs, p.num_hyps_per_beam)
    b = res.log_probs.shape[0]
    dev = res.log_probs.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(random_seed))
    ids = torch.full((b, 1), p.target_sos_id, dtype=torch.int64, device=dev) \
        if init_step_ids is None else init_step_ids.reshape(b, 1).long()
    all_ids, all_logits = [], []
    done = torch.zeros(b, dtype=torch.bool, device=dev)
    for t in range(p.target_seq_len):
      res, state = pre_step_callback(decoder_theta, encoder_outputs, ids, state,
                                     p.num_hyps_per_beam, t)
      logits = self._Filter(res.log_probs)
      nxt = torch.multinomial(torch.softmax(logits, -1), 1, generator=gen)
      nxt = torch.where(done.unsqueeze(1), torch.full_like(nxt, p.target_eos_id), nxt)
      all_ids.append(nxt)
      all_logits.append(logits)
      done = done | (nxt.squeeze(1) == p.target_eos_id)
      ids = nxt
      if post_step_callback is not None:
        state = post_step_callback(decoder_theta, encoder_outputs, ids, state)
      if p.use_stop_fn and (t + 1) % 8 == 0 and bool(done.all()):
        break
    out_ids = torch.cat(all_ids, 1)
    return NestedMap(logits=torch.stack(all_logits, 1), ids=out_ids,
                     paddings=_ComputePaddings(out_ids, p.target_eos_id))
