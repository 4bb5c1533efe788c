"""Beam / greedy search against brute force on a toy Markov decoder (CPU)."""

import itertools

import torch

from lingvo_b200.core import beam_search_helper
from lingvo_b200.core.nested_map import NestedMap

V, EOS, SOS = 5, 2, 1


def _Table(seed=0):
  g = torch.Generator().manual_seed(seed)
  return torch.log_softmax(torch.randn(V, V, generator=g) * 1.5, -1)


def _Callbacks(table, num_beams):
  def init(theta, enc, k):
    n = num_beams * k
    return (NestedMap(log_probs=torch.zeros(n, V), atten_probs=torch.zeros(n, 3)),
            NestedMap(steps=torch.zeros(n, 1)))

  def pre(theta, enc, step_ids, states, k, t):
    lp = table[step_ids.squeeze(1)]
    att = torch.full((step_ids.shape[0], 3), 1.0 / 3)
    return NestedMap(log_probs=lp, atten_probs=att), NestedMap(steps=states.steps + 1)

  def post(theta, enc, step_ids, states):
    return states
  return init, pre, post


def _BruteForce(table, max_len):
  """All sequences ending in EOS with ≤ max_len tokens; returns sorted (score, seq)."""
  out = []
  toks = [t for t in range(V) if t != EOS]
  for n in range(0, max_len):
    for seq in itertools.product(toks, repeat=n):
      full = list(seq) + [EOS]
      prev, s = SOS, 0.0
      for t in full:
        s += float(table[prev, t])
        prev = t
      out.append((s, full))
  return sorted(out, key=lambda x: -x[0])


def test_beam_search_finds_global_best_with_wide_beam():
  table = _Table()
  p = beam_search_helper.BeamSearchHelper.Params().Set(
      num_hyps_per_beam=16, beam_size=1e9, target_seq_len=4, target_sos_id=SOS,
      target_eos_id=EOS, valid_eos_max_logit_delta=1e9, local_eos_threshold=-1e9,
      sync_every=1)
  h = p.Instantiate()
  init, pre, post = _Callbacks(table, num_beams=2)
  out = h.BeamSearchDecode(None, NestedMap(), init_beam_search_state=init,
                           pre_beam_search_step_callback=pre,
                           post_beam_search_step_callback=post)
  want = _BruteForce(table, 4)
  for beam in range(2):
    got_scores = out.topk_hyps.scores[beam]
    got_ids = out.topk_hyps.ids[beam]
    lens = out.topk_hyps.lens[beam]
    for r in range(5):
      assert abs(float(got_scores[r]) - want[r][0]) < 1e-5, (r, got_scores[:6], want[:6])
      assert got_ids[r, :int(lens[r])].tolist() == want[r][1]
  # states were carried and re-ordered without shape damage
  assert out.other_states.steps.shape == (32, 1)


def test_beam_size_pruning_terminates_early():
  table = _Table(1)
  p = beam_search_helper.BeamSearchHelper.Params().Set(
      num_hyps_per_beam=4, beam_size=0.5, target_seq_len=12, target_sos_id=SOS,
      target_eos_id=EOS, sync_every=1)
  h = p.Instantiate()
  init, pre, post = _Callbacks(table, num_beams=1)
  out = h.BeamSearchDecode(None, NestedMap(), init_beam_search_state=init,
                           pre_beam_search_step_callback=pre,
                           post_beam_search_step_callback=post)
  assert int(out.topk_hyps.lens[0, 0]) >= 1
  assert out.topk_ids.shape[1] <= 12
  assert float(out.topk_hyps.scores[0, 0]) > -1e29


def test_length_normalization_prefers_longer():
  table = _Table(2)
  def run(alpha):
    p = beam_search_helper.BeamSearchHelper.Params().Set(
        num_hyps_per_beam=8, beam_size=1e9, target_seq_len=5, target_sos_id=SOS,
        target_eos_id=EOS, valid_eos_max_logit_delta=1e9, local_eos_threshold=-1e9,
        length_normalization=alpha, sync_every=1)
    init, pre, post = _Callbacks(table, 1)
    o = p.Instantiate().BeamSearchDecode(None, NestedMap(), init_beam_search_state=init,
                                         pre_beam_search_step_callback=pre,
                                         post_beam_search_step_callback=post)
    return float(o.topk_hyps.lens[0].float().mean())
  assert run(2.0) >= run(0.0)


def test_greedy_search():
  table = _Table(3)
  p = beam_search_helper.GreedySearchHelper.Params().Set(
      target_seq_len=6, target_sos_id=SOS, target_eos_id=EOS)
  init, pre, post = _Callbacks(table, 3)
  ids, lens, done = p.Instantiate().GreedySearchDecode(
      None, NestedMap(), init_beam_search_state=init,
      pre_beam_search_step_callback=pre, post_beam_search_step_callback=post)
  prev = SOS
  for t in range(int(lens[0])):
    assert int(ids[0, t]) == int(table[prev].argmax())
    prev = int(ids[0, t])


def test_merge_outputs():
  table = _Table()
  p = beam_search_helper.BeamSearchHelper.Params().Set(
      num_hyps_per_beam=4, beam_size=1e9, target_seq_len=3, target_sos_id=SOS,
      target_eos_id=EOS, sync_every=1)
  init, pre, post = _Callbacks(table, 1)
  o = p.Instantiate().BeamSearchDecode(None, NestedMap(), init_beam_search_state=init,
                                       pre_beam_search_step_callback=pre,
                                       post_beam_search_step_callback=post)
  m = beam_search_helper.MergeBeamSearchOutputs(3, [o, o])
  assert m.topk_hyps.ids.shape[:2] == (1, 3)


# ---- round 2: option surface of the reference op (force_eos_in_top_k, merge_paths) ----
def _OneStep(log_probs, k, **kw):
  from lingvo_b200.ops import beam_search as bs
  n = log_probs.shape[0]
  st = bs.init_state(n // k, k, 4, 1, log_probs.device)
  return bs.beam_search_step(log_probs, None, st, 0, eos_id=2, beam_size=3.0,
                             num_hyps_per_beam=k, **kw)


def test_force_eos_in_top_k_controls_whether_low_ranked_eos_may_terminate():
  import torch
  from lingvo_b200.ops import beam_search as bs
  k, v = 2, 8
  lp = torch.full((k, v), -9.0)
  lp[:, 5], lp[:, 6], lp[:, 2] = -0.5, -1.0, -2.0          # EOS (id 2) is only third best
  lp = torch.log_softmax(lp, -1)
  st, _ = _OneStep(lp, k, valid_eos_max_logit_delta=10.0)
  assert float(st.done_scores[0].max()) < bs.NEG / 2        # not in the hyp's top-2 ⇒ no EOS
  st2, _ = _OneStep(lp, k, valid_eos_max_logit_delta=10.0, force_eos_in_top_k=True)
  assert float(st2.done_scores[0].max()) > bs.NEG / 2       # forced into the candidate set
  assert st2.hyps[0].tolist() == st.hyps[0].tolist() == [5, 6]


def test_merge_paths_combines_candidates_that_differ_only_by_epsilons():
  import torch
  from lingvo_b200.ops import beam_search as bs
  k, v, eoc = 3, 6, 1
  n = k                                     # one beam
  st = bs.init_state(1, k, 4, 1, 'cpu')
  # after step 0 force three live hyps with distinct histories: h0 = [4], h1 = [eps], h2 = [5]
  st = st._replace(cumulative_scores=torch.tensor([-1.0, -1.2, -3.0]))
  path_ids = torch.tensor([(0 * 1000003 + 4 + 1), 0, (0 * 1000003 + 5 + 1)], dtype=torch.int64)
  lp = torch.full((n, v), -20.0)
  lp[0, eoc] = -0.1          # h0 emits epsilon   → path [4]
  lp[1, 4] = -0.2            # h1 emits 4         → path [4]  (same label sequence)
  lp[2, 3] = -0.3            # h2 emits 3         → path [5, 3]
  new, _, new_ids = bs.beam_search_step(lp, None, st, 1, eos_id=2, beam_size=100.0,
                                        num_hyps_per_beam=k, merge_paths=True, eoc_id=eoc,
                                        path_ids=path_ids)
  cum = new.cumulative_scores
  live = cum > bs.NEG / 2
  assert int(live.sum()) == 2                                  # the duplicate path is gone
  want = torch.logsumexp(torch.tensor([-1.0 - 0.1, -1.2 - 0.2]), 0)
  assert abs(float(cum[live].max()) - float(want)) < 1e-5     # mass of both alignments
  assert len(set(new_ids[live].tolist())) == 2
