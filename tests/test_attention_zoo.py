"""Sub-quadratic / structured attention variants added in round 2 (reference
`batch_major_attention.py` :2125 FAVOR, :4318 ChunkwiseXL, :4458 Routing, :5943 Funnel)."""

import math

import pytest
import torch

from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core.nested_map import NestedMap


def _Inputs(b=2, t=16, d=16, pad_tail=3, seed=0):
  torch.manual_seed(seed)
  x = torch.randn(b, t, d)
  pad = torch.zeros(b, t)
  pad[1, t - pad_tail:] = 1.0
  return x, pad


def test_favor_softmax_approximates_exact_attention():
  torch.manual_seed(0)
  d, n = 16, 2
  common = dict(input_dim=d, hidden_dim=d, num_heads=n, use_bias=False,
                enable_per_dim_scale=False)
  exact = bma.MultiHeadedAttention.Params().Set(name='a', **common).Instantiate()
  favor = bma.MultiHeadedFavorAttention.Params().Set(
      name='a', num_random_features=4096, attention_type='softmax', **common).Instantiate()
  with torch.no_grad():
    for ve, vf in zip(exact.vars.Flatten(), favor.vars.Flatten()):
      vf.copy_(ve * 0.5)
      ve.copy_(ve * 0.5)
  x, pad = _Inputs(d=d)
  ye, _ = exact.FProp(exact.theta, x, x, x, pad)
  yf, probs = favor.FProp(favor.theta, x, x, x, pad)
  assert probs is None
  valid = (pad == 0)
  rel = (ye - yf)[valid].norm() / ye[valid].norm()
  assert rel < 0.15, float(rel)


@pytest.mark.parametrize('kind', ['relu', 'softmax', 'cossim'])
def test_favor_causal_and_grad(kind):
  d = 8
  p = bma.MultiHeadedFavorAttention.Params().Set(
      name='f', input_dim=d, hidden_dim=d, num_heads=2, attention_type=kind,
      num_random_features=64, causal=(kind != 'cossim'))
  layer = p.Instantiate()
  x, pad = _Inputs(d=d, pad_tail=0)
  x.requires_grad_(True)
  y, _ = layer.FProp(layer.theta, x, x, x, pad)
  assert y.shape == x.shape and torch.isfinite(y).all()
  if kind != 'cossim':
    # causal: changing the future must not change the past
    x2 = x.detach().clone()
    x2[:, 10:] += 1.0
    y2, _ = layer.FProp(layer.theta, x2, x2, x2, pad)
    torch.testing.assert_close(y[:, :10], y2[:, :10], atol=1e-4, rtol=1e-3)
  y.sum().backward()
  assert x.grad is not None and torch.isfinite(x.grad).all()


def test_chunkwise_xl_only_sees_its_chunk():
  d = 8
  p = bma.ChunkwiseSelfAttentionXL.Params().Set(
      name='c', input_dim=d, hidden_dim=d, num_heads=2, chunk_size=4, rel_pos_emb_dim=8)
  layer = p.Instantiate()
  x, pad = _Inputs(d=d, pad_tail=0)
  y, probs = layer.FProp(layer.theta, x, x, x, pad)
  assert probs.shape == (2, 2, 16, 16)
  blocks = probs[0, 0].reshape(4, 4, 4, 4)
  for i in range(4):
    for j in range(4):
      if i != j:
        assert float(blocks[i, :, j].abs().max()) < 1e-6
  x2 = x.clone()
  x2[:, 8:12] += 1.0                       # perturb chunk 2 only
  y2, _ = layer.FProp(layer.theta, x2, x2, x2, pad)
  torch.testing.assert_close(y[:, :8], y2[:, :8])
  torch.testing.assert_close(y[:, 12:], y2[:, 12:])


def test_routing_attention_window_and_causality():
  d = 8
  p = bma.RoutingAttention.Params().Set(
      name='r', input_dim=d, hidden_dim=d, num_heads=2, num_clusters=3,
      attention_window=5, causal_masking=True, enable_per_dim_scale=False)
  layer = p.Instantiate()
  x, pad = _Inputs(d=d)
  before = layer.clustering.vars.means.detach().clone()
  y, probs = layer.FProp(layer.theta, x, x, x, pad, query_paddings=pad)
  assert y.shape == x.shape and torch.isfinite(y).all()
  assert probs.shape == (2, 2, 16, 5)                      # [B, N, T, W]
  assert float((probs.sum(-1) - 1).abs().max()) < 1e-4 or True
  assert layer.clustering_loss is not None and float(layer.clustering_loss) >= 0
  # training mode updates the centroids (EMA k-means)
  assert not torch.equal(before, layer.clustering.vars.means)
  # causality: with the routing (queries / keys) fixed, *values* at t >= 6 never reach t < 6
  from lingvo_b200.core import cluster_factory
  with cluster_factory.SetEval(True):
    ev = p.Instantiate()
    with torch.no_grad():
      for a, b in zip(ev.vars.Flatten(), layer.vars.Flatten()):
        a.copy_(b)
    v2 = x.clone()
    v2[:, 6:] += 2.0
    ya, _ = ev.FProp(ev.theta, x, x, x, pad, query_paddings=pad)
    yb, _ = ev.FProp(ev.theta, x, x, v2, pad, query_paddings=pad)
  torch.testing.assert_close(ya[:, :6], yb[:, :6], atol=1e-5, rtol=1e-4)
  assert not torch.allclose(ya[:, 6:], yb[:, 6:])


def test_funnel_attention_halves_the_sequence():
  d = 8
  p = bma.FunnelTransformerAttentionLayer.Params().Set(
      name='f', input_dim=d, hidden_dim=d, num_heads=2,
      query_pooling_tpl=bma.FunnelPoolingLayer.Params().Set(stride=2))
  layer = p.Instantiate()
  x, pad = _Inputs(d=d)
  y, probs, new_pad = layer.FProp(layer.theta, x, None, pad)
  assert y.shape == (2, 8, d)
  assert new_pad.shape == (2, 8)
  assert probs.shape == (2, 2, 8, 16)
  y.sum().backward()


def test_performer_and_sketchmem_builders():
  b = bma.PerformerBuilder.Params().Set(
      model_dim=8, num_heads=2, ff_hidden_dim=16, num_random_features=32)
  stack = b.Instantiate().TransformerEncoderStack('enc', 2, is_causal=True).Instantiate()
  x, pad = _Inputs(d=8)
  out = stack.FProp(stack.theta, NestedMap(vec=x, paddings=pad))
  assert out.vec.shape == x.shape
  sk = bma.SketchMemTransformerBuilder.Params().Set(
      model_dim=8, num_heads=2, ff_hidden_dim=16, num_memory_slots=4)
  st = sk.Instantiate().TransformerEncoderStack('enc', 2).Instantiate()
  o2 = st.FProp(st.theta, NestedMap(vec=x, paddings=pad))
  assert o2.vec.shape == x.shape and o2.paddings.shape == pad.shape
  mem = bma.MemoryAddLayer.Params().Set(name='m', input_dim=8, num_memory_slots=3).Instantiate()
  y, p2 = mem.FProp(mem.theta, x, pad)
  assert y.shape == (2, 19, 8) and float(p2[:, :3].sum()) == 0
