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
    # causal: changing the future must not change the past (up to the key stabiliser's
    # epsilon term, whose weight depends on the max over ALL keys, as in the reference)
    x2 = x.detach().clone()
    x2[:, 10:] += 1.0
    y2, _ = layer.FProp(layer.theta, x2, x2, x2, pad)
    torch.testing.assert_close(y[:, :10], y2[:, :10], atol=3e-3, rtol=0)
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


# ------------------------------------------------------------- self_attention_layer.py --
def test_block_sparse_attention_equals_dense_with_block_diagonal_mask():
  from lingvo_b200.core import self_attention_layer as sal
  p = sal.BlockSparseAttention.Params().Set(
      name='bs', input_dim=16, hidden_dim=16, num_heads=4, src_block_size=4, tgt_block_size=4)
  layer = p.Instantiate()
  x = torch.randn(2, 12, 16, requires_grad=True)
  pad = torch.zeros(2, 12)
  pad[1, 10:] = 1
  y, probs = layer.FPropDefaultTheta(x, x, x, pad)
  assert y.shape == (2, 12, 16) and probs is None
  dense = bma.MultiHeadedAttention.Params().Set(name='bs', input_dim=16, hidden_dim=16,
                                                num_heads=4).Instantiate()
  blk = torch.arange(12) // 4
  psp = (blk[:, None] != blk[None, :]).float().unsqueeze(0).expand(2, 12, 12)
  want, _ = dense.FProp(layer.theta, x, x, x, pad, per_step_padding=psp)
  valid = (1 - pad).unsqueeze(-1)
  torch.testing.assert_close(y * valid, want * valid, atol=1e-5, rtol=1e-5)
  # block l never sees other blocks: perturbing block 2 leaves blocks 0-1 untouched
  x2 = x.detach().clone()
  x2[:, 8:] += 1.0
  y2, _ = layer.FPropDefaultTheta(x2, x2, x2, pad)
  torch.testing.assert_close(y2[:, :8], y.detach()[:, :8])
  y.sum().backward()
  assert x.grad is not None
  for bad in (dict(src_block_size=None), dict(src_block_size=4, packed_input=True)):
    with pytest.raises(AssertionError):
      sal.BlockSparseAttention.Params().Set(name='b', input_dim=8, hidden_dim=8,
                                            **bad).Instantiate()
  with pytest.raises(AssertionError):
    layer.FPropDefaultTheta(x[:, :10], x[:, :10], x[:, :10], pad[:, :10])   # 10 % 4 != 0


def test_self_attention_builder_per_layer_templates_and_strided_final_layer():
  from lingvo_b200.core import self_attention_layer as sal
  dense = bma.MultiHeadedAttention.Params()
  sparse = sal.BlockSparseAttention.Params().Set(src_block_size=4, tgt_block_size=4)
  b = sal.Builder.Params().Set(model_dim=16, num_heads=2, ff_hidden_dim=32,
                               atten_tpl=[dense, sparse]).Instantiate()
  stack = b.TransformerStack('stack', 2).Instantiate()
  x = NestedMap(vec=torch.randn(2, 8, 16), paddings=torch.zeros(2, 8))
  out = stack.FPropDefaultTheta(x)
  assert out.vec.shape == (2, 8, 16)
  def Walk(layer):
    yield layer
    for c in layer.children.Flatten():
      yield from Walk(c)
  kinds = [type(l).__name__ for l in Walk(stack) if isinstance(l, bma.MultiHeadedAttention)]
  assert kinds == ['MultiHeadedAttention', 'BlockSparseAttention']
  with pytest.raises(AssertionError):
    b.TransformerStack('bad', 3)                                  # list length ≠ layers

  b1 = sal.Builder.Params().Set(model_dim=16, num_heads=2, ff_hidden_dim=32).Instantiate()
  strided = b1.TransformerStackV2('s', 2, final_layer_stride=2).Instantiate()
  pad = torch.zeros(2, 8)
  pad[1, 6:] = 1
  o = strided.FPropDefaultTheta(NestedMap(vec=torch.randn(2, 8, 16), paddings=pad))
  assert o.vec.shape == (2, 4, 16) and torch.equal(o.paddings, pad[:, ::2])
  firstn = b1.TransformerStackV2('f', 2, final_layer_first_n=3).Instantiate()
  o = firstn.FPropDefaultTheta(NestedMap(vec=torch.randn(2, 8, 16), paddings=pad))
  assert o.vec.shape == (2, 3, 16) and o.paddings.shape == (2, 3)
  plain = b1.TransformerStackV2('p', 2).Instantiate()
  assert plain.FPropDefaultTheta(NestedMap(vec=torch.randn(2, 8, 16),
                                           paddings=pad)).vec.shape == (2, 8, 16)


@pytest.mark.parametrize('parallel', [False, True])
def test_simplified_transformer_builder(parallel):
  from lingvo_b200.core import self_attention_layer as sal
  b = sal.SimplifiedTransformerBuilder.Params().Set(
      model_dim=16, num_heads=2, ff_hidden_dim=32, parallel_attention_mlp=parallel).Instantiate()
  assert b.params.atten_tpl.enable_shaped_attention
  stack = b.TransformerStack('s', 2).Instantiate()
  x = torch.randn(2, 6, 16, requires_grad=True)
  pad = torch.zeros(2, 6)
  pad[1, 4:] = 1
  out = stack.FPropDefaultTheta(NestedMap(vec=x, paddings=pad))
  assert out.vec.shape == (2, 6, 16) and torch.equal(out.paddings, pad)
  out.vec.sum().backward()
  assert x.grad is not None and torch.isfinite(x.grad).all()
  blk = stack.children['iter_000'].children['block']
  assert ('feedforward' in blk.children) == parallel and ('ff' in blk.children) != parallel
  # no skip connection around attention: with the attention output projection zeroed the
  # block output no longer depends on x in sequential mode (FF sees zeros) …
  v2 = b.TransformerStackV2('v2', 2, final_layer_stride=2).Instantiate()
  o2 = v2.FPropDefaultTheta(NestedMap(vec=x.detach(), paddings=pad))
  assert o2.vec.shape == (2, 3, 16) and torch.equal(o2.paddings, pad[:, ::2])
  if parallel:
    assert float(o2.vec[1, 2].abs().sum()) == 0.0          # padded position is zeroed


def test_favor_chunked_causal_custom_gradients_match_autograd():
  from lingvo_b200.core import favor_attention as fa
  torch.manual_seed(0)
  l, b, h, m, d = 150, 2, 2, 5, 3                   # not a multiple of the 64-step chunk
  qs = torch.rand(l, b, h, m, dtype=torch.float64, requires_grad=True)
  ks = torch.rand(l, b, h, m, dtype=torch.float64, requires_grad=True)
  vs = torch.randn(l, b, h, d, dtype=torch.float64, requires_grad=True)
  num = fa.chunked_causal_numerator(qs, ks, vs)
  den = fa.chunked_causal_denominator(qs, ks)
  ref_num = fa.causal_numerator(qs, ks, vs)         # plain autograd formulation
  ref_den = fa.causal_denominator(qs, ks)
  torch.testing.assert_close(num, ref_num)
  torch.testing.assert_close(den, ref_den)
  wn, wd = torch.randn_like(num), torch.randn_like(den)
  got = torch.autograd.grad((num * wn).sum() + (den * wd).sum(), [qs, ks, vs])
  want = torch.autograd.grad((ref_num * wn).sum() + (ref_den * wd).sum(), [qs, ks, vs])
  for g, w in zip(got, want):
    torch.testing.assert_close(g, w, rtol=1e-9, atol=1e-9)
  # the func / grad pairs are usable on their own (reference API)
  out, sums = fa.chunked_causal_numerator_func(qs.detach(), ks.detach(), vs.detach())
  dq, dk, dv = fa.chunked_causal_numerator_grad(qs.detach(), ks.detach(), vs.detach(), sums, wn)
  torch.testing.assert_close(out, ref_num.detach())
  assert dq.shape == qs.shape and dk.shape == ks.shape and dv.shape == vs.shape


def _BmaBuilder(**kw):
  return bma.Builder.Params().Set(model_dim=8, num_heads=2, ff_hidden_dim=16, **kw).Instantiate()


def _BInput(b=2, t=8, d=8):
  torch.manual_seed(0)
  pad = torch.zeros(b, t)
  pad[1, 6:] = 1
  return NestedMap(vec=torch.randn(b, t, d), paddings=pad)


def test_builder_gated_feedforward_and_glu():
  from lingvo_b200.core.nested_map import NestedMap as NM
  b = _BmaBuilder()
  ffn = b.GatedGeluFeedforward('gff').Instantiate()
  i = _BInput()
  o = ffn.FPropDefaultTheta(i)
  assert o.vec.shape == i.vec.shape and float(o.vec[1, 6:].abs().max()) == 0
  names = {v.var_name for v in ffn.vars.Flatten()}
  assert any(n.endswith('feedforward/wi0/w/var') for n in names)
  assert len(names) == 5                                        # LN scale+bias, wi0, wi1, wo: no linear biases
  ff = ffn.GetDescendant('feedforward') if hasattr(ffn, 'GetDescendant') else None
  x = i.vec
  ln = ff.ln.FPropDefaultTheta(x)
  h = torch.nn.functional.gelu(ln @ ff.wi0.vars.w, approximate='tanh') * (ln @ ff.wi1.vars.w)
  want = (x + h @ ff.wo.vars.w) * (1 - i.paddings).unsqueeze(-1)
  torch.testing.assert_close(o.vec, want, atol=1e-5, rtol=1e-5)
  glu = b._Glu('glu').Instantiate()
  v = torch.randn(2, 3, 8)
  torch.testing.assert_close(glu.FProp(NM(), v), v[..., 4:] * torch.sigmoid(v[..., :4]))
  tanh = _BmaBuilder(glu_with_tanh=True)._Glu('glu').Instantiate()
  torch.testing.assert_close(tanh.FProp(NM(), v), torch.tanh(v[..., 4:]) * torch.sigmoid(v[..., :4]))
  nores = _BmaBuilder(ff_apply_residual=False, ff_use_paddings=False).GatedFeedforward('g').Instantiate()
  assert nores.FPropDefaultTheta(i).vec.shape == i.vec.shape


def test_builder_lconv_stack_moe_and_funnel_layer():
  b = _BmaBuilder()
  stack = b.LConvStack('lconv', [3, 5], is_causal=True).Instantiate()
  i = _BInput()
  o = stack.FPropDefaultTheta(i)
  assert o.vec.shape == i.vec.shape
  i2 = NestedMap(vec=i.vec.clone(), paddings=i.paddings)
  i2.vec[0, 5:] += 1.0                                        # causal: the past is unaffected
  o2 = stack.FPropDefaultTheta(i2)
  torch.testing.assert_close(o.vec[0, :5], o2.vec[0, :5], atol=1e-5, rtol=1e-5)
  piped = _BmaBuilder(num_splits=2, num_micro_batches=2).LConvStack('lp', [3, 3]).Instantiate()
  assert type(piped).__name__ == 'PipeliningLayer' and piped.num_stages == 2
  moe = _BmaBuilder(num_experts=2, num_groups=1).MoE('moe').Instantiate()
  om = moe.FPropDefaultTheta(i)
  assert om.vec.shape == i.vec.shape and torch.isfinite(om.vec).all()
  fun = b.FunnelEncoderLayer('funnel', stride=2).Instantiate()
  of = fun.FPropDefaultTheta(i)
  assert of.vec.shape == (2, 4, 8) and of.paddings.shape == (2, 4)
  assert of.paddings[1].tolist() == [0, 0, 0, 1]
  same = b.FunnelEncoderLayer('f1', stride=1).Instantiate().FPropDefaultTheta(i)
  assert same.vec.shape == i.vec.shape
  assert b.Seq('s', b._Id('a'), b._Id('b')).cls.__name__ == 'SequentialLayer'


def test_positional_atten_logits_rel_position_bias_and_one_step():
  from lingvo_b200.core import attention_util
  torch.manual_seed(0)
  b, t, n, h = 2, 5, 3, 4
  lyr = attention_util.PositionalAttenLogits.Params().Set(name='pal').Instantiate()
  q, k = torch.randn(b, t, n, h), torch.randn(b, t, n, h)
  emb = torch.randn(2 * t - 1, n, h)
  u, v = torch.randn(n, h), torch.randn(n, h)
  want = torch.zeros(b, n, t, t)
  for i in range(t):
    for j in range(t):
      want[:, :, i, j] = ((q[:, i] + u) * k[:, j]).sum(-1) + \
          ((q[:, i] + v) * emb[i - j + t - 1]).sum(-1)
  got = lyr._AttenLogits(q, k, emb, u, v)
  torch.testing.assert_close(got, want, atol=1e-5, rtol=1e-5)
  content = torch.einsum('BTNH,BSNH->BNTS', q + u, k)
  torch.testing.assert_close(lyr.AttenLogitsXL(q, k, emb, u, v), want, atol=1e-5,
                             rtol=1e-5)
  # skip_term_b drops the q·R term.
  skip = lyr._AttenLogits(q, k, emb, u, v, skip_term_b=True)
  want_skip = content + lyr.RelPositionBias(v, emb, True).unsqueeze(0)
  torch.testing.assert_close(skip, want_skip)
  assert lyr.RelPositionBias(v, emb, True).shape == (n, t, t)
  rpe = lyr.AttenLogitsRPE(q, k, emb)
  torch.testing.assert_close(rpe, lyr._AttenLogits(q, k, emb, torch.zeros(n, h),
                                                   torch.zeros(n, h)))
  # One step at time i equals row i of the full logits when fed the right embedding rows.
  i = 3
  key_sb = k.transpose(0, 1)                                         # [S,B,N,H]
  step_emb = torch.stack([emb[i - j + t - 1] for j in range(t)])     # [S,N,H]
  one = lyr.AttenLogitsXLOneStep(q[:, i], key_sb, step_emb, u, v)
  torch.testing.assert_close(one, want[:, :, i, :].permute(2, 0, 1), atol=1e-5, rtol=1e-5)
  per_seq = lyr.AttenLogitsXLOneStep(q[:, i], key_sb, step_emb.unsqueeze(0).expand(b, -1, -1, -1),
                                     u, v)
  torch.testing.assert_close(per_seq, one)
  one_skip = lyr.AttenLogitsXLOneStep(q[:, i], key_sb, step_emb, u, v, skip_term_b=True)
  torch.testing.assert_close(one_skip, skip[:, :, i, :].permute(2, 0, 1), atol=1e-5, rtol=1e-5)
  rpe1 = lyr.AttenLogitsRPEOneStep(q[:, i], key_sb, step_emb.unsqueeze(1))
  want_rpe1 = torch.einsum('BNH,SBNH->SBN', q[:, i], key_sb + step_emb.unsqueeze(1))
  torch.testing.assert_close(rpe1, want_rpe1)
  import pytest
  with pytest.raises(ValueError):
    lyr._AttenLogits(q, k, emb, torch.zeros(n, h + 1), v)
