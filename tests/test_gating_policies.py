"""Gating policies added in round 2 (ref lingvo/core/gshard_layers_test.py)."""
import numpy as np
import torch

from lingvo_b200.core import gshard_layers as gl
from lingvo_b200.core.nested_map import NestedMap


def test_token_shuffle_v2_skips_padding_and_respects_capacity():
  torch.manual_seed(0)
  g, s, e, c = 2, 16, 4, 6
  logits = torch.randn(g, s, e)
  pad = torch.zeros(g, s)
  pad[:, 12:] = 1
  aux, comb, disp = gl.TokenShufflingOnlogitsV2(logits, pad, 1, e, c, torch.float32)
  assert float(aux) == 0 and comb.shape == (g, s, e, c)
  assert comb[:, 12:].abs().sum() == 0                  # padded tokens go nowhere
  per_slot = disp.sum(1)                                # GEC: ≤ 1 token per slot
  assert per_slot.max() <= 1
  assert torch.all(disp.sum((1, 3)) == c)               # each expert is full (12 ≥ 6 tokens)
  gates = torch.softmax(logits, -1)
  sel = disp.sum(-1)                                    # GSE
  np.testing.assert_allclose(comb.sum(-1).numpy(), (gates * sel).numpy(), rtol=1e-6)
  # each expert picked its c best non-padded tokens
  for gi in range(g):
    for ei in range(e):
      want = set(gates[gi, :12, ei].topk(c).indices.tolist())
      got = set(torch.nonzero(sel[gi, :, ei]).flatten().tolist())
      assert want == got
  # slots are in sequence order
  slot = (disp * torch.arange(c)).sum(-1)
  for gi in range(g):
    for ei in range(e):
      toks = torch.nonzero(sel[gi, :, ei]).flatten()
      assert slot[gi, toks, ei].tolist() == list(range(c))


def test_optimal_transport_gating_is_balanced():
  torch.manual_seed(1)
  g, s, e = 2, 32, 4
  logits = torch.randn(g, s, e) + torch.tensor([3.0, 0, 0, 0])   # everyone prefers expert 0
  aux, comb, disp = gl.OptimalTransportOnlogits(logits, e, fprop_dtype=torch.float32)
  cap = s * 2 // e
  assert comb.shape == (g, s, e, cap) and float(aux) == 0
  assert torch.all(disp.sum((1, 3)) == cap)             # every expert takes exactly C tokens
  per_token = disp.sum((2, 3))
  assert per_token.max() <= e and per_token.float().mean() == 2.0
  # the plan spreads load: most tokens get ≤ 3 experts even with the biased logits
  assert (per_token <= 3).float().mean() > 0.9
  out = gl.OptimalTransportGating(torch.randn(8, e), torch.randn(g, s, 8), None, 1, e, cap,
                                  True, torch.float32)
  assert out.combine_tensor.shape == (g, s, e, cap)


def test_sentence_embeddings_and_gating():
  x = torch.tensor([[0.4, 0.6], [0.6, 0.4], [0.8, 0.8], [0.5, 0.9], [0.6, 0.3],
                    [0.4, 0.8], [0.6, 0.6], [0.0, 1.0], [0.2, 0.4], [0.1, 0.5]]).reshape(1, 10, 2)
  seg = torch.tensor([[1, 1, 2, 0, 0, 3, 3, 3, 3, 0]])
  emb = gl.GetSentenceEmbeddings(x, seg)
  np.testing.assert_allclose(emb[0, 0].numpy(), [0.5, 0.5], rtol=1e-6)
  np.testing.assert_allclose(emb[0, 1].numpy(), [0.5, 0.5], rtol=1e-6)
  np.testing.assert_allclose(emb[0, 2].numpy(), [0.8, 0.8], rtol=1e-6)
  assert emb[0, 3].abs().sum() == 0 and emb[0, 9].abs().sum() == 0
  np.testing.assert_allclose(emb[0, 5].numpy(), [0.3, 0.7], rtol=1e-6)
  torch.manual_seed(0)
  w = torch.randn(2, 4)
  pad = (seg == 0).float()
  out = gl.SentenceTop2Gating(w, x, pad, seg, 1, 4, 8, True, torch.float32)
  sel = out.dispatch_tensor.sum(-1)                     # GSE
  # every token of a sentence goes to the same experts
  for a, b in [(0, 1), (5, 6), (6, 7), (7, 8)]:
    assert torch.equal(sel[0, a], sel[0, b])
  assert sel[0, 3].sum() == 0
  # task gating: routing depends on task embeddings only
  task = torch.randn(1, 1, 2).expand(1, 10, 2)
  out2 = gl.TaskTop2Gating(w, x, pad, task, 1, 4, 10, True, torch.float32)
  sel2 = out2.dispatch_tensor.sum(-1)
  live = torch.nonzero(pad[0] == 0).flatten()
  for i in live[1:]:
    assert torch.equal(sel2[0, live[0]], sel2[0, i])


def test_gating_wrappers_dispatch_to_policies():
  torch.manual_seed(0)
  g, s, m, e, c = 2, 8, 4, 4, 4
  w, x = torch.randn(m, e), torch.randn(g, s, m)
  for fn in (gl.Top2Gating, gl.TokenShuffleGating, gl.TokenShuffleGatingV2):
    out = fn(w, x, None, 1, e, c, True, torch.float32)
    assert out.combine_tensor.shape[:3] == (g, s, e)
    assert out.dispatch_tensor.shape == out.combine_tensor.shape
  ids = torch.randint(0, e, (g, s))
  out = gl.HashGating(w, x, None, 1, e, c, True, torch.float32, expert_id=ids)
  assert torch.equal(out.dispatch_tensor.sum(-1).argmax(-1)[out.dispatch_tensor.sum((-1, -2)) > 0],
                     ids[out.dispatch_tensor.sum((-1, -2)) > 0])
  # non-local dispatch reshapes back to the input's leading dims
  out = gl.TokenShuffleGatingV2(w, x, None, 1, e, c, False, torch.float32)
  assert out.combine_tensor.shape[:2] == (g, s)


def test_gather_k_matches_reference_example():
  sel = torch.tensor([[0, 0, 1, 1], [0, 1, 1, 0], [0, 0, 0, 0], [1, 1, 1, 0], [1, 1, 1, 1]])
  v = torch.tensor([[1, 3, 5, 7], [9, 11, 13, 15], [17, 19, 21, 23], [25, 27, 29, 31],
                    [33, 35, 37, 39]], dtype=torch.float32)
  (out, out3), pad = gl.GatherK(sel, [v, v.unsqueeze(-1).repeat(1, 1, 2)], 3)
  assert pad.tolist() == [[1, 0, 0], [1, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0]]
  live = (1 - pad)
  assert (out * live).tolist() == [[0, 5, 7], [0, 11, 13], [0, 0, 0], [25, 27, 29],
                                   [35, 37, 39]]
  assert out3.shape == (5, 3, 2) and torch.equal(out3[..., 0], out)


def test_conv1d_state_layer_matches_full_causal_window():
  p = gl.Conv1DStateLayer.Params().Set(name='cs', shape=[None, None, 3], kernel_size=3)
  layer = p.Instantiate()
  b, beam, t = 2, 2, 5
  x = torch.randn(b, beam, t, 3)
  state = layer.InitState(b, beam, dtype=torch.float32)
  for i in range(t):
    win, state = layer.Step(state, x[:, :, i])
    assert win.shape == (b * beam, 3, 3)
    want = torch.zeros(b, beam, 3, 3)
    lo = max(0, i - 2)
    want[:, :, 3 - (i - lo + 1):] = x[:, :, lo:i + 1]
    np.testing.assert_allclose(win.reshape(b, beam, 3, 3).numpy(), want.numpy())
  # prefix: window = last k prefix inputs for every beam
  pre = torch.randn(b, 4, 3)
  st = layer.LoadPrefix(layer.InitState(b, beam, dtype=torch.float32), pre)
  assert torch.equal(st[:, 0], pre[:, -3:]) and torch.equal(st[:, 1], pre[:, -3:])
  # zero inputs skipped
  layer2 = p.Copy().Set(skip_store_zero_state=True).Instantiate()
  st2 = st.clone()
  _, st3 = layer2.Step(st2, torch.zeros(b, beam, 3))
  assert torch.equal(st3, st)
  # beam reorder
  parent = torch.tensor([[1, 1], [0, 0]])
  stx = torch.arange(b * beam * 3 * 3, dtype=torch.float32).reshape(b, beam, 3, 3)
  re = gl.Conv1DStateLayer.Reorder(stx, None, parent)
  assert torch.equal(re[0, 0], stx[0, 1]) and torch.equal(re[1, 1], stx[1, 0])
  assert gl.ShardedWeightParams([4, 8], tensor_split_dims_mapping=[-1, 0]
                                ).tensor_split_dims_mapping == [-1, 0]


def test_compute_gating_accepts_the_reference_positional_signature():
  torch.manual_seed(0)
  g, s, m, e = 2, 8, 4, 4
  w, x = torch.randn(m, e), torch.randn(g, s, m)
  pad = torch.zeros(g, s)
  pos = gl.ComputeGating(w, x, pad, 1, e, 0, True, torch.float32, 'top_2', False, 'all', 0.0,
                         True, 2.0, None, torch.float32, torch.float32)
  kw = gl.ComputeGating(w, x, pad, 1, e, 0, True, torch.float32, gating_func='top_2',
                        use_xla_sharding=False, capacity_factor=2.0, mask_dtype=torch.float32,
                        gating_logits_dtype=torch.float32)
  torch.testing.assert_close(pos.combine_tensor, kw.combine_tensor)
