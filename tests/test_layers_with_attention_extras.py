"""Round-2 additions to `layers_with_attention.py`: TransformerShardedMoeLayer (reference
:832), EvolvedTransformer enc/dec (:1807-2290), StyleLayer, TransformerWithContextLayer,
CCT layers."""

import pytest
import torch

from lingvo_b200.core import gshard_layers
from lingvo_b200.core import layers_with_attention as lwa
from lingvo_b200.core import py_utils


def _TB(t=6, b=3, d=16, seed=0):
  torch.manual_seed(seed)
  x = torch.randn(t, b, d)
  pad = torch.zeros(t, b)
  pad[-2:, 1] = 1.0
  return x, pad


@pytest.mark.parametrize('use_glu', [False, True])
def test_sharded_moe_layer_matches_dense_gshard_oracle(use_glu):
  torch.manual_seed(1)
  b, t, d, h, e = 4, 8, 16, 32, 4
  p = lwa.TransformerShardedMoeLayer.Params().Set(
      name='moe', input_dim=d, hidden_dim=h, num_experts=e, num_groups=2,
      expert_capacity_factor=2.0, use_glu=use_glu, expert_weight_shards=2)
  layer = p.Instantiate()
  x = torch.randn(b, t, d, requires_grad=True)
  pad = torch.zeros(b, t)
  pad[1, 5:] = 1.0
  with py_utils.AuxLossContext() as ctx:
    y = layer.FPropDefaultTheta(x, pad)
    assert len(ctx.aux_losses) == 1 and float(ctx.aux_losses[0]) > 0
  assert y.shape == x.shape
  # oracle: dense one-hot einsums (the reference's GPU path) on the same gating weights
  th = layer.theta
  xn = layer.layer_norm.FProp(th.layer_norm, x)
  g, s = 2, b * t // 2
  xg = xn.reshape(g, s, d)
  pg = pad.reshape(g, s)
  logits = torch.matmul(xg, th.gate)
  gating = gshard_layers.Top2GatingOnLogits(
      None, pg, logits, 1, e, 0, torch.float32, capacity_factor=4.0,
      legacy_mtf_behavior=False)
  _, combine, dispatch = gating
  wi = torch.cat([th.wi_0, th.wi_1], -1)
  wo = torch.cat([th.wo_0, th.wo_1], 1)
  xe = torch.einsum('GSEC,GSM->EGCM', dispatch.float(), xg)
  hh = torch.relu(torch.einsum('EGCM,EMH->EGCH', xe, wi))
  if use_glu:
    hh = hh * torch.einsum('EGCM,EMH->EGCH', xe, torch.cat([th.wi_gate_0, th.wi_gate_1], -1))
  ye = torch.einsum('EGCH,EHM->EGCM', hh, wo)
  yo = torch.einsum('GSEC,EGCM->GSM', combine, ye).reshape(b, t, d)
  want = x + yo * (1 - pad).unsqueeze(-1)
  torch.testing.assert_close(y, want, atol=2e-5, rtol=1e-4)
  y.sum().backward()
  assert all(v.grad is not None for v in layer.vars.Flatten() if v.requires_grad)


def test_sharded_moe_layer_expert_choice_and_min_group_size():
  p = lwa.TransformerShardedMoeLayer.Params().Set(
      name='moe', input_dim=8, hidden_dim=16, num_experts=2, num_groups=8,
      min_group_size=16, gating_func='expert_choice', expert_capacity_factor=1.0)
  layer = p.Instantiate()
  assert layer._NumGroups(4 * 8, 4) == 2            # 32 tokens / 16 per group
  x = torch.randn(4, 8, 8)
  y = layer.FPropDefaultTheta(x, torch.zeros(4, 8))
  assert y.shape == x.shape and torch.isfinite(y).all()
  y4 = layer.FPropDefaultTheta(x.reshape(4, 8, 2, 4), torch.zeros(4, 8))
  assert y4.shape == (4, 8, 2, 4)


def test_evolved_transformer_encoder_and_decoder_layers():
  d = 16
  x, pad = _TB(d=d)
  enc = lwa.EvolvedTransformerEncoderLayer.Params().Set(name='enc', source_dim=d)
  enc.transformer_tpl.tr_atten_tpl.num_attention_heads = 2
  enc.transformer_tpl.tr_fflayer_tpl.hidden_dim = 32
  e = enc.Instantiate()
  y, probs = e.FPropDefaultTheta(x, pad)
  assert y.shape == x.shape and torch.isfinite(y).all()
  dec = lwa.EvolvedTransformerDecoderLayer.Params().Set(name='dec', source_dim=d)
  dec.tr_atten_tpl.num_attention_heads = 2
  dec.tr_double_heads_atten_tpl.num_attention_heads = 4
  dec.transformer_tpl.tr_atten_tpl.num_attention_heads = 2
  dec.transformer_tpl.tr_fflayer_tpl.hidden_dim = 32
  dl = dec.Instantiate()
  src, spad = _TB(t=5, d=d, seed=1)
  out, _ = dl.FPropDefaultTheta(x, pad, src, spad)
  assert out.shape == x.shape
  # decoder is causal in time: perturbing the last step leaves earlier outputs unchanged
  x2 = x.clone()
  x2[-1] += 1.0
  out2, _ = dl.FPropDefaultTheta(x2, pad, src, spad)
  torch.testing.assert_close(out[:-1], out2[:-1], atol=1e-5, rtol=1e-4)
  out.sum().backward()


def test_style_layer_mixture_and_lookup():
  p = lwa.StyleLayer.Params().Set(name='style', input_dim=12, output_dim=8, num_styles=5,
                                  num_heads=2)
  s = p.Instantiate()
  q = torch.randn(3, 12)
  emb = s.FPropDefaultTheta(q)
  assert emb.shape == (3, 8)
  one = s.EmbLookup(s.theta, torch.tensor([0, 4, 2]))
  assert one.shape == (3, 8) and float(one.abs().max()) <= 1.0
  probs = torch.softmax(torch.randn(3, 5), -1)
  mix = s.StyleEmbFromProbs(s.theta, probs)
  assert mix.shape == (3, 8)
  onehot = torch.eye(5)[[0, 4, 2]]
  torch.testing.assert_close(s.StyleEmbFromProbs(s.theta, onehot), one)


def test_transformer_with_context_layer_uses_all_three_sources():
  d = 16
  p = lwa.TransformerWithContextLayer.Params().Set(name='ctx', source_dim=d)
  p.tr_atten_tpl.num_attention_heads = 2
  p.tr_fflayer_tpl.hidden_dim = 32
  layer = p.Instantiate()
  x, pad = _TB(d=d)
  src, spad = _TB(t=5, d=d, seed=1)
  ctxv, cpad = _TB(t=4, d=d, seed=2)
  y, probs = layer.FPropDefaultTheta(x, pad, src, spad, ctxv, cpad)
  assert y.shape == x.shape and probs.shape[-1] == 5
  y2, _ = layer.FPropDefaultTheta(x, pad, src, spad, ctxv + 1.0, cpad)
  assert not torch.allclose(y, y2)


def test_cct_layers_gate_between_identity_and_transform():
  d = 16
  x, pad = _TB(d=d)
  ff = lwa.CCTFeedForwardLayer.Params().Set(name='ff', input_dim=d, hidden_dim=32,
                                            num_blocks=4).Instantiate()
  y = ff.FPropDefaultTheta(x, pad)
  assert y.shape == x.shape and 0.0 <= float(ff.last_gate_mean) <= 1.0
  at = lwa.CCTAttentionLayer.Params().Set(name='at', source_dim=d, num_attention_heads=2)
  a = at.Instantiate()
  z, probs = a.FPropDefaultTheta(x, pad)
  assert z.shape == x.shape and probs is not None
  (y.sum() + z.sum()).backward()
