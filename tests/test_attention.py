"""Batch-major / time-major attention and Transformer layer tests (CPU).

Mirrors the reference's strategy (`batch_major_attention_test.py`,
`attention_test.py`): shape checks, FProp ≡ step-by-step ExtendStep, mask
semantics, and equivalence between layer variants.
"""

import numpy as np
import pytest
import torch

from lingvo_b200.core import attention
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import layers_with_attention as lwa
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def _Inputs(b=2, t=6, d=8, seed=0):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(b, t, d, generator=g)
  pad = torch.zeros(b, t)
  if b > 1:
    pad[1, t - 2:] = 1.0
  return x, pad


def _Mha(cls=bma.MultiHeadedAttention, **kw):
  p = cls.Params().Set(name='atten', input_dim=8, hidden_dim=8, num_heads=2, **kw)
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  return p.Instantiate()


def test_mha_matches_manual_softmax():
  l = _Mha(enable_per_dim_scale=False)
  x, pad = _Inputs()
  out, probs = l.FPropDefaultTheta(x, x, x, pad)
  assert out.shape == (2, 6, 8) and probs.shape == (2, 2, 6, 6)
  th = l.theta
  q = torch.einsum('BTD,DNH->BTNH', x, th.query.w) + th.query.b
  k = torch.einsum('BTD,DNH->BTNH', x, th.key.w) + th.key.b
  v = torch.einsum('BTD,DNH->BTNH', x, th.value.w) + th.value.b
  logits = torch.einsum('BTNH,BSNH->BNTS', q * 4 ** -0.5, k)
  logits = logits.masked_fill(pad.view(2, 1, 1, 6) > 0, -1e30)
  pr = torch.softmax(logits, -1)
  ctx = torch.einsum('BNTS,BSNH->BTNH', pr, v)
  ref = torch.einsum('BTNH,DNH->BTD', ctx, th.post.w) + th.post.b
  torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(probs, pr, atol=1e-6, rtol=1e-5)
  # padded keys get zero probability
  assert float(probs[1, :, :, 4:].abs().max()) < 1e-12


def test_fused_path_equals_probs_path():
  a = _Mha()
  b = _Mha(return_atten_probs=False)
  x, pad = _Inputs()
  o1, p1 = a.FPropDefaultTheta(x, x, x, pad)
  o2, p2 = b.FProp(a.theta, x, x, x, pad)
  assert p2 is None and p1 is not None
  torch.testing.assert_close(o1, o2, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('kw', [dict(), dict(num_kv_heads=1), dict(use_mqa=True)])
def test_extend_step_matches_fprop(kw):
  l = _Mha(**kw)
  x, _ = _Inputs()
  b, t, _ = x.shape
  pad = torch.zeros(b, t)
  causal = bma.CausalPadding(t).unsqueeze(0).expand(b, t, t)
  full, _ = l.FPropDefaultTheta(x, x, x, pad, per_step_padding=causal)
  st = l.InitStates(l.theta, b, t)
  outs = []
  for i in range(t):
    o, st = l.ExtendStep(l.theta, x[:, i:i + 1], st, None, time_step=i)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-5, rtol=1e-5)


def test_rope_extend_step():
  from lingvo_b200.core import layers
  l = _Mha(rope_tpl=layers.RotaryPositionalEmbeddingLayer.Params())
  x, _ = _Inputs()
  b, t, _ = x.shape
  causal = bma.CausalPadding(t).unsqueeze(0).expand(b, t, t)
  full, _ = l.FPropDefaultTheta(x, x, x, torch.zeros(b, t), per_step_padding=causal)
  st = l.InitStates(l.theta, b, t)
  outs = []
  for i in range(t):
    o, st = l.ExtendStep(l.theta, x[:, i:i + 1], st, None, time_step=i)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-5, rtol=1e-5)


def test_segment_mask_isolates_segments():
  l = _Mha(packed_input=True)
  x, _ = _Inputs(b=1, t=6)
  seg = torch.tensor([[1, 1, 1, 2, 2, 2]])
  mask = bma.SegmentMask(seg, seg)
  out, probs = l.FPropDefaultTheta(x, x, x, torch.zeros(1, 6), segment_mask=mask)
  assert float(probs[0, :, :3, 3:].abs().max()) < 1e-12
  # running the segments separately gives the same answer
  o1, _ = l.FProp(l.theta, x[:, :3], x[:, :3], x[:, :3], torch.zeros(1, 3),
                  segment_mask=torch.zeros(1, 1, 3, 3))
  torch.testing.assert_close(out[:, :3], o1, atol=1e-5, rtol=1e-5)


def test_causal_segment_mask():
  seg = torch.tensor([[1, 1, 2, 2]])
  m = bma.CausalSegmentMask(seg)
  allowed = (m[0, 0] == 0).int().tolist()
  assert allowed == [[1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 1, 1]]


def test_xl_attention_zero_bias_and_shift():
  l = _Mha(bma.MultiHeadedAttentionXL, rel_pos_emb_dim=8)
  x, pad = _Inputs()
  out, probs = l.FPropDefaultTheta(x, x, x, pad)
  assert out.shape == (2, 6, 8)
  # brute-force relative term
  th = l.theta
  q = (torch.einsum('BTD,DNH->BTNH', x, th.query.w) + th.query.b)
  q = l.per_dim_scale.FProp(th.per_dim_scale, q)
  k = torch.einsum('BTD,DNH->BTNH', x, th.key.w) + th.key.b
  t = 6
  logits = torch.einsum('BTNH,BSNH->BNTS', q + th.u, k)
  for i in range(t):
    for j in range(t):
      sin = bma._SinusoidTable(torch.tensor([i - j]), 8, x.device)
      r = torch.einsum('LD,DNH->LNH', sin, th.pos_proj.w)[0]
      logits[:, :, i, j] += torch.einsum('BNH,NH->BN', q[:, i] + th.v, r)
  logits = logits.masked_fill(pad.view(2, 1, 1, 6) > 0, -1e30)
  torch.testing.assert_close(probs, torch.softmax(logits, -1), atol=1e-5, rtol=1e-4)


def test_rpe_attention_runs_and_extend():
  l = _Mha(bma.MultiHeadedAttentionRPE, rel_pos_radius=3, skip_value_emb=True)
  x, _ = _Inputs()
  b, t, _ = x.shape
  causal = bma.CausalPadding(t).unsqueeze(0).expand(b, t, t)
  full, _ = l.FPropDefaultTheta(x, x, x, torch.zeros(b, t), per_step_padding=causal)
  st = l.InitStates(l.theta, b, t)
  outs = []
  for i in range(t):
    o, st = l.ExtendStep(l.theta, x[:, i:i + 1], st, None, time_step=i)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-5, rtol=1e-5)


def test_local_attention_band():
  l = _Mha(bma.LocalSelfAttention, left_context=2, right_context=1)
  x, pad = _Inputs(b=1, t=8)
  _, probs = l.FPropDefaultTheta(x, x, x, torch.zeros(1, 8))
  nz = (probs[0, 0] > 0).int()
  for i in range(8):
    for j in range(8):
      assert int(nz[i, j]) == int(i - 1 <= j <= i + 1)


def test_local_attention_stream_step_matches_fprop():
  l = _Mha(bma.LocalSelfAttention, left_context=3, right_context=0,
           return_atten_probs=False)
  x, _ = _Inputs(b=2, t=8)
  full, _ = l.FPropDefaultTheta(x, x, x, torch.zeros(2, 8))
  st = l.zero_state(2)
  outs = []
  for i in range(0, 8, 2):
    o, _, st = l.StreamStep(l.theta, x[:, i:i + 2], torch.zeros(2, 2), st)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-5, rtol=1e-5)


def test_chunkwise_attention():
  l = _Mha(bma.ChunkwiseSelfAttention, chunk_size=4)
  x, _ = _Inputs(b=1, t=8)
  _, probs = l.FPropDefaultTheta(x, x, x, torch.zeros(1, 8))
  assert float(probs[0, :, :4, 4:].abs().max()) == 0
  assert float(probs[0, :, 4:, :4].abs().max()) == 0


def _Stack(n=2, **kw):
  p = bma.StackedTransformerLayers.Params().Set(
      name='stack', num_layers=n, mdl_dim=8, hidden_dim=16, num_atten_heads=2, **kw)
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  return p.Instantiate()


def test_stacked_transformer_shapes_and_grads():
  l = _Stack(final_layer_norm=True)
  x, pad = _Inputs()
  out, pad_out = l.FPropDefaultTheta(x, pad)
  assert out.shape == x.shape and pad_out is pad
  out.sum().backward()
  for v in l.vars.Flatten():
    assert v.grad is not None, v.var_name
  names = sorted(v.var_name for v in l.vars.Flatten())
  assert any(n.endswith('layer_0/self_atten/atten/query/w/var') for n in names), names[:5]


def test_stacked_decoder_extend_step():
  l = _Stack(mask_self_atten=True, has_aux_atten=True)
  x, _ = _Inputs()
  aux, aux_pad = _Inputs(t=5, seed=3)
  b, t, _ = x.shape
  full, _ = l.FPropDefaultTheta(x, torch.zeros(b, t), aux, aux_pad)
  st = l.InitStates(l.theta, b, t)
  outs = []
  for i in range(t):
    o, st = l.ExtendStep(l.theta, x[:, i:i + 1], aux, aux_pad, st, i)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-4, rtol=1e-4)


def test_builder_stack():
  b = bma.Builder.Params().Set(model_dim=8, num_heads=2, ff_hidden_dim=16).Instantiate()
  p = b.TransformerEncoderStack('enc', 2)
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  l = p.Instantiate()
  x, pad = _Inputs()
  out = l.FPropDefaultTheta(NestedMap(vec=x, paddings=pad))
  assert out.vec.shape == x.shape
  assert float(out.vec[1, 4:].abs().max()) == 0      # padded frames zeroed


def test_funnel_pooling_and_upsample():
  pool = bma.FunnelPoolingLayer.Params().Set(name='pool', stride=2).Instantiate()
  x = torch.arange(12.).view(1, 6, 2)
  pad = torch.tensor([[0., 0, 0, 0, 1, 1]])
  y, py = pool.FPropDefaultTheta(x, pad)
  assert y.shape == (1, 3, 2)
  torch.testing.assert_close(y[0, 0], (x[0, 0] + x[0, 1]) / 2)
  assert py.tolist() == [[0, 0, 1]]
  up = bma.FunnelUpsampleLayer.Params().Set(name='up', upsample_rate=2).Instantiate()
  assert up.FPropDefaultTheta(y).shape == (1, 6, 2)


# ---------------------------------------------------------------- time-major ----
def _Src(t=5, b=2, d=8, seed=1):
  g = torch.Generator().manual_seed(seed)
  src = torch.randn(t, b, d, generator=g)
  pad = torch.zeros(t, b)
  pad[t - 1, 0] = 1.0
  return src, pad


@pytest.mark.parametrize('cls,kw', [
    (attention.AdditiveAttention, dict(hidden_dim=7)),
    (attention.DotProductAttention, dict(hidden_dim=8)),
    (attention.MultiHeadedAttention, dict(hidden_dim=8, context_dim=8,
                                          num_attention_heads=2)),
])
def test_time_major_attention_contract(cls, kw):
  p = cls.Params().Set(name='a', source_dim=8, query_dim=8, **kw)
  p.params_init = py_utils.WeightInit.Gaussian(0.3)
  l = p.Instantiate()
  src, pad = _Src()
  packed = l.InitForSourcePacked(l.theta, src, src, pad)
  q = torch.randn(4, 8)            # query batch = 2 × source batch (beam search)
  ctx, probs, _ = l.ComputeContextVector(l.theta, q)
  assert ctx.shape[0] == 4 and probs.shape == (4, 5)
  torch.testing.assert_close(probs.sum(-1), torch.ones(4), atol=1e-5, rtol=1e-5)
  assert float(probs[0, 4]) == 0 and float(probs[2, 4]) == 0   # source row 0 padded
  assert float(probs[1, 4]) > 0
  del packed


def test_location_sensitive_and_monotonic():
  p = attention.LocationSensitiveAttention.Params().Set(
      name='loc', source_dim=8, query_dim=8, hidden_dim=6, location_filter_size=3,
      location_num_filters=4)
  l = p.Instantiate()
  src, pad = _Src()
  l.InitForSourcePacked(l.theta, src, src, pad)
  st = l.ZeroAttentionState(5, 2)
  ctx, probs, st1 = l.ComputeContextVector(l.theta, torch.randn(2, 8), st)
  assert ctx.shape == (2, 8) and st1.shape == st.shape
  m = attention.MonotonicAttention.Params().Set(
      name='mono', source_dim=8, query_dim=8, hidden_dim=6).Instantiate()
  m.InitForSourcePacked(m.theta, src, src, pad)
  st = m.ZeroAttentionState(5, 2)
  ctx, probs, st = m.ComputeContextVector(m.theta, torch.randn(2, 8), st)
  assert probs.shape == (2, 5) and float(probs.sum(-1).max()) <= 1.0 + 1e-5


def test_monotonic_prob_modes_agree():
  g = torch.Generator().manual_seed(0)
  pc = torch.rand(3, 7, generator=g)
  prev = torch.softmax(torch.randn(3, 7, generator=g), -1)
  a = attention.MonotonicAttentionProb(pc, prev, 'parallel')
  b = attention.MonotonicAttentionProb(pc, prev, 'recursive')
  torch.testing.assert_close(a, b, atol=1e-5, rtol=1e-4)


def test_time_major_transformer_layer_extend_step():
  p = lwa.TransformerLayer.Params().Set(
      name='tr', source_dim=8, mask_self_atten=True, has_aux_atten=True)
  p.tr_atten_tpl.num_attention_heads = 2
  p.tr_fflayer_tpl.hidden_dim = 16
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  l = p.Instantiate()
  g = torch.Generator().manual_seed(0)
  x = torch.randn(6, 2, 8, generator=g)
  aux, aux_pad = _Src()
  full, probs = l.FPropDefaultTheta(x, torch.zeros(6, 2), aux, aux_pad)
  assert full.shape == x.shape and probs.shape == (6, 2, 5)
  st = NestedMap(key=torch.zeros(0, 2, 8), value=torch.zeros(0, 2, 8))
  outs = []
  for i in range(6):
    o, _, st = l.ExtendStep(l.theta, x[i], st, aux, aux_pad)
    outs.append(o)
  torch.testing.assert_close(torch.stack(outs), full, atol=1e-4, rtol=1e-4)


def test_feed_forward_layer_variants():
  for act in ('RELU', 'GATED_GELU'):
    p = lwa.TransformerFeedForwardLayer.Params().Set(
        name='ff', input_dim=8, hidden_dim=16, activation=act)
    l = p.Instantiate()
    x, pad = _Inputs()
    y = l.FPropDefaultTheta(x, pad)
    assert y.shape == x.shape
  p = lwa.TransformerFeedForwardLayer.Params().Set(
      name='ff2', input_dim=8, output_dim=12, hidden_dim=16)
  assert p.Instantiate().FPropDefaultTheta(*_Inputs()).shape == (2, 6, 12)


def test_self_attentive_layer():
  l = lwa.SelfAttentiveLayer.Params().Set(name='sa', input_dim=8, hidden_dim=5,
                                          num_heads=3).Instantiate()
  x, pad = _Inputs()
  out, pen = l.FPropDefaultTheta(x, pad)
  assert out.shape == (2, 3, 8) and pen.dim() == 0


def test_gshard_rel_table_matches_bias_path():
  """`_RelTable` (fast path input) reproduces the Toeplitz bias of `_Bias`."""
  from lingvo_b200.core import gshard_builder
  from lingvo_b200.ops import attention as A
  b = gshard_builder.DenseBuilder.Params().Set(
      model_dim=16, attention_num_heads=2, attention_key_value_dim=8,
      relative_attention_num_buckets=8, relative_attention_max_distance=16,
      relative_attention_use_universal_1d_position=True)
  lp = gshard_builder.SelfAttentionLayer.Params().Set(
      name='sa', b=b, relative_bias=True, decoder=True)
  layer = lp.Instantiate()
  l = 12
  seg = torch.ones(1, l, dtype=torch.int32)
  pos = torch.arange(l).unsqueeze(0)
  bias = layer._Bias(layer.theta, seg, pos)
  mask = layer._Mask(seg, pos, torch.float32)
  rel = layer._RelTable(layer.theta, l, torch.device('cpu'))
  rebuilt = A._RelToeplitz(rel, l).unsqueeze(0) + mask
  torch.testing.assert_close(bias, rebuilt)
