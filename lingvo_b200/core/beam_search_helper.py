"""Beam / greedy search drivers (ref `lingvo/core/beam_search_helper.py`).

`BeamSearchHelper.BeamSearchDecode` (ref :200-650) drives a decoder through
three callbacks —

  InitBeamSearchState(theta, encoder_outputs, num_hyps_per_beam)
      → (initial_results {log_probs, atten_probs}, other_states)
  PreBeamSearchStepCallback(theta, encoder_outputs, step_ids [n,1], states,
                            num_hyps_per_beam, cur_step)
      → (results {log_probs [n,V], atten_probs [n,S]}, new_states)
  PostBeamSearchStepCallback(theta, encoder_outputs, new_step_ids, states)
      → final_states

— calling the device beam-search step (`ops.beam_search`) between Pre and Post,
re-ordering every tensor of `other_states` by the surviving parent indices.
The loop checks `all_done` on the host only every `sync_every` steps, so the
GPU runs ahead of Python.
"""

from __future__ import annotations

import collections

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.ops import beam_search as bs_ops

BeamSearchDecodeOutput = collections.namedtuple(
    'BeamSearchDecodeOutput',
    ['topk_hyps', 'topk_ids', 'topk_lens', 'topk_scores', 'topk_decoded',
     'other_states'])
BeamSearchDecodeOutput.__new__.__defaults__ = (None,) * 6


class BeamSearchSharedParams(base_layer.BaseLayer):
  """Params shared by all search helpers (ref :89)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_hyps_per_beam', 8, 'Hyps kept per beam.')
    p.Define('target_seq_length_ratio', 1.0, 'Avg target/source length ratio.')
    p.Define('length_normalization', 0.0, 'Exponent α on (len+5)/5.')
    p.Define('coverage_penalty', 0.0, 'Coverage penalty β.')
    p.Define('valid_eos_max_logit_delta', 5.0, 'EOS must be within this of the best token.')
    p.Define('local_eos_threshold', -100.0, 'EOS local score threshold.')
    p.Define('beam_size', 3.0, 'Max score gap best-terminated vs active.')
    p.Define('target_sos_id', 1, 'SOS id.')
    p.Define('target_eos_id', 2, 'EOS id.')
    p.Define('target_eoc_id', -1, 'End-of-chunk id (NT only).')
    p.Define('target_seq_len', 0, 'Max decode steps.')
    p.Define('merge_paths', False, 'Merge hyps equal up to epsilons (RNN-T / NT; needs '
             'target_eoc_id).')
    p.Define('force_eos_in_top_k', False, 'EOS is always a candidate of its hyp.')
    p.Define('force_last_chunk_eoc_in_top_k', False, 'Kept for parity.')
    p.Define('batch_major_state', True, 'States are [hyp, …].')
    p.Define('batch_major_compute', False, 'Kept for parity.')
    p.Define('short_seq_limit', 0, 'Kept for parity.')
    p.Define('terminate_beams_independently', False,
             'Finished beams become no-ops (always the case here).')
    return p


class BeamSearchHelper(BeamSearchSharedParams):
  """Beam search over a step-wise decoder (ref :200)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('allow_empty_terminated_hyp', True, 'Allow </s> as the first token.')
    p.Define('ensure_full_beam', False, 'Stop only with K terminated hyps.')
    p.Define('force_eos_in_last_step', False, 'Force-terminate at the last step.')
    p.Define('atten_vecs_in_hypothesis_protos', True, 'Keep attention for coverage.')
    p.Define('merged_topk_buffer_size_factor', 2, 'Kept for parity.')
    p.Define('reorder_tarzan_states', True, 'Kept for parity.')
    p.Define('sync_every', 8, 'Host-side all_done check period (steps).')
    p.name = 'beam_search'
    return p

  def _ReOrder(self, states, parent):
    """Gathers dim-0 of every state tensor by the surviving parent rows."""
    def _One(x):
      if not isinstance(x, torch.Tensor) or x.dim() == 0 or x.shape[0] != parent.shape[0]:
        return x
      return x.index_select(0, parent)
    return states.Transform(_One)

  def BeamSearchDecode(self, theta, encoder_outputs, num_hyps_per_beam_override=0,
                       init_beam_search_state=None,
                       pre_beam_search_step_callback=None,
                       post_beam_search_step_callback=None, max_steps=None):
    p = self.params
    k = num_hyps_per_beam_override or p.num_hyps_per_beam
    max_steps = max_steps or p.target_seq_len
    assert max_steps > 0
    init_results, other_states = init_beam_search_state(theta, encoder_outputs, k)
    n = init_results.log_probs.shape[0]
    b = n // k
    dev = init_results.log_probs.device
    src_len = init_results.atten_probs.shape[-1] if init_results.get(
        'atten_probs') is not None else 1
    state = bs_ops.init_state(b, k, max_steps, src_len, dev)
    step_ids = torch.full((n, 1), p.target_sos_id, dtype=torch.int64, device=dev)
    if init_results.get('step_ids') is not None:     # decoder-provided first input ids
      step_ids = init_results.step_ids.reshape(n, 1).to(torch.int64)
    steps_run = 0
    path_ids = torch.zeros(n, dtype=torch.int64, device=dev) if p.merge_paths else None
    for t in range(max_steps):
      results, other_states = pre_beam_search_step_callback(
          theta, encoder_outputs, step_ids, other_states, k, t)
      step_out = bs_ops.beam_search_step(
          results.log_probs, results.get('atten_probs'), state, t,
          eos_id=p.target_eos_id, beam_size=p.beam_size, num_hyps_per_beam=k,
          valid_eos_max_logit_delta=p.valid_eos_max_logit_delta,
          local_eos_threshold=p.local_eos_threshold,
          ensure_full_beam=p.ensure_full_beam,
          force_eos_in_last_step=p.force_eos_in_last_step,
          is_last_step=(t == max_steps - 1),
          allow_empty_terminated_hyp=p.allow_empty_terminated_hyp,
          force_eos_in_top_k=p.force_eos_in_top_k,
          beam_independence=True, merge_paths=p.merge_paths, eoc_id=p.target_eoc_id,
          path_ids=path_ids)
      if p.merge_paths:
        state, all_done, path_ids = step_out
      else:
        state, all_done = step_out
      steps_run = t + 1
      parent = state.prev_hyps[t]
      step_ids = state.hyps[t].reshape(n, 1)
      other_states = self._ReOrder(other_states, parent)
      if post_beam_search_step_callback is not None:
        other_states = post_beam_search_step_callback(
            theta, encoder_outputs, step_ids, other_states)
      if (t + 1) % max(p.sync_every, 1) == 0 and bool(all_done):
        break
    src_lens = None
    if p.coverage_penalty > 0:
      pad = encoder_outputs.get('padding')
      if pad is not None:                      # [S, B] time-major paddings
        src_lens = (1.0 - pad.float()).sum(0)
      else:
        src_lens = torch.full((b,), float(src_len), device=dev)
    ids, lens, scores = bs_ops.top_k_terminated_hyps(
        state, src_lens, k, steps_run, p.length_normalization, p.coverage_penalty,
        p.target_seq_length_ratio, p.target_eos_id)
    t_used = ids.shape[-1]
    return BeamSearchDecodeOutput(
        topk_hyps=NestedMap(ids=ids, lens=lens, scores=scores),
        topk_ids=ids.reshape(b * k, t_used),
        topk_lens=lens.reshape(-1),
        topk_scores=scores,
        topk_decoded=None,
        other_states=other_states)


def MergeBeamSearchOutputs(max_hyps_per_beam, beam_search_outputs):
  """Merges several decode outputs for the same sources, keeping the best
  `max_hyps_per_beam` per beam (ref :681)."""
  ids = torch.cat([o.topk_hyps.ids for o in beam_search_outputs], 1)
  lens = torch.cat([o.topk_hyps.lens for o in beam_search_outputs], 1)
  scores = torch.cat([o.topk_hyps.scores for o in beam_search_outputs], 1)
  order = torch.argsort(scores, dim=1, descending=True, stable=True)[:, :max_hyps_per_beam]
  ids = ids.gather(1, order.unsqueeze(-1).expand(-1, -1, ids.shape[-1]))
  lens, scores = lens.gather(1, order), scores.gather(1, order)
  b = ids.shape[0]
  return BeamSearchDecodeOutput(
      topk_hyps=NestedMap(ids=ids, lens=lens, scores=scores),
      topk_ids=ids.reshape(b * max_hyps_per_beam, -1), topk_lens=lens.reshape(-1),
      topk_scores=scores, topk_decoded=None, other_states=None)


class GreedySearchHelper(base_layer.BaseLayer):
  """Arg-max decoding with the same callbacks (ref :752)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('target_sos_id', 1, 'SOS id.')
    p.Define('target_eos_id', 2, 'EOS id.')
    p.Define('target_seq_len', 0, 'Max decode steps.')
    p.name = 'greedy_search'
    return p

  def GreedySearchDecode(self, theta, encoder_outputs, init_beam_search_state=None,
                         pre_beam_search_step_callback=None,
                         post_beam_search_step_callback=None, max_steps=None):
    """Returns (hyp_ids [B,T], hyp_lens [B], done_hyps [B])."""
    p = self.params
    max_steps = max_steps or p.target_seq_len
    init_results, states = init_beam_search_state(theta, encoder_outputs, 1)
    b = init_results.log_probs.shape[0]
    dev = init_results.log_probs.device
    step_ids = torch.full((b, 1), p.target_sos_id, dtype=torch.int64, device=dev)
    ids = torch.full((b, max_steps), p.target_eos_id, dtype=torch.int64, device=dev)
    lens = torch.zeros(b, dtype=torch.int64, device=dev)
    done = torch.zeros(b, dtype=torch.bool, device=dev)
    for t in range(max_steps):
      results, states = pre_beam_search_step_callback(
          theta, encoder_outputs, step_ids, states, 1, t)
      nxt = results.log_probs.argmax(-1)
      ids[:, t] = torch.where(done, ids[:, t], nxt)
      lens = lens + (~done).to(torch.int64)
      done = done | (nxt == p.target_eos_id)
      step_ids = nxt.reshape(b, 1)
      if post_beam_search_step_callback is not None:
        states = post_beam_search_step_callback(theta, encoder_outputs, step_ids, states)
      if (t + 1) % 8 == 0 and bool(done.all()):
        break
    return ids, lens, done
