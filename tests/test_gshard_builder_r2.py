"""GShard builder layers added in round 2 (ref lingvo/core/gshard_builder_test.py)."""
import numpy as np
import pytest
import torch

from lingvo_b200.core import gshard_builder as gb
from lingvo_b200.core.nested_map import NestedMap


def _Builder(**kw):
  p = gb.DenseBuilder.Params().Set(
      model_dim=16, attention_num_heads=2, attention_key_value_dim=8, ff_dim=32,
      relative_attention_type='bias', relative_attention_num_buckets=8,
      relative_attention_max_distance=16, fprop_dtype=torch.float32, dtype=torch.float32,
      label_smoothing=0.0, **kw)
  return p.Instantiate()


def _Inst(lp, seed=7):
  lp.random_seed = seed
  layer = lp.Instantiate()
  layer.InstantiateVariables()
  return layer


def _Inputs(b=2, l=6, m=16, seed=0):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(b, l, m, generator=g)
  seg = torch.tensor([[1, 1, 1, 2, 2, 0], [1, 1, 1, 1, 1, 1]])[:b, :l]
  pos = torch.tensor([[0, 1, 2, 0, 1, 0], [0, 1, 2, 3, 4, 5]])[:b, :l]
  return x, seg, pos


def test_depthwise_conv_autoregressive_is_causal_and_segment_local():
  b = _Builder()
  layer = _Inst(b.DepthwiseConvAutoregressive('dconv', 3))
  assert [float(layer.vars['w_%d' % k][0]) for k in range(3)] == pytest.approx(
      [0.5, 0.5 / 3, 0.5 / 3])
  x, seg, pos = _Inputs()
  with torch.no_grad():
    for k in range(3):
      layer.vars['w_%d' % k].copy_(torch.randn(16))
  y = layer.FProp(layer.theta, x, pos)
  w = [layer.vars['w_%d' % k] for k in range(3)]
  want = torch.zeros_like(x)
  for bi in range(2):
    for t in range(6):
      for k in range(3):
        if int(pos[bi, t]) >= k:                    # does not reach before the segment start
          want[bi, t] += w[k] * x[bi, t - k]
  torch.testing.assert_close(y, want)
  # per-head variant + incremental decoding equals the full pass
  ph = _Inst(b.DepthwiseConvAutoregressive('ph', 3, model_dims=[2, 8]))
  with torch.no_grad():
    for k in range(3):
      ph.vars['w_%d' % k].copy_(torch.randn(2, 8))
  xh = torch.randn(2, 5, 2, 8)
  full = ph.FProp(ph.theta, xh)
  st = ph.InitState(2, xh.device, xh.dtype)
  outs = []
  for t in range(5):
    o, st = ph.ExtendStep(ph.theta, xh[:, t:t + 1], st)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-6, rtol=1e-6)
  assert b.CausalDepthwiseConv('c', 3).cls is gb.DepthwiseConvAutoregressiveLayer


def test_dec_enc_attention_masks_other_segments_and_decodes_incrementally():
  b = _Builder()
  layer = _Inst(b.DecEncAttention('cross'))
  x, seg, pos = _Inputs()
  enc = torch.randn(2, 4, 16)
  enc_seg = torch.tensor([[1, 1, 2, 2], [1, 1, 1, 0]])
  out, aux = layer.FProp(layer.theta, x, seg, pos, encoder_output=enc,
                         encoder_segment_id=enc_seg)
  assert out.shape == x.shape and float(aux) == 0.0
  # decoder segment 1 of example 0 must not depend on encoder segment 2
  enc2 = enc.clone(); enc2[0, 2:] += 5.0
  out2, _ = layer.FProp(layer.theta, x, seg, pos, encoder_output=enc2,
                        encoder_segment_id=enc_seg)
  torch.testing.assert_close(out[0, :3], out2[0, :3])
  assert (out[0, 3:5] - out2[0, 3:5]).abs().max() > 1e-4
  # cached K/V path = full path
  kv = layer.ProjectEncoder(layer.theta, enc)
  out3, _ = layer.FProp(layer.theta, x, seg, pos, kv=kv, encoder_segment_id=enc_seg)
  torch.testing.assert_close(out3, out)
  # generic attention core with a BLM bias
  core = _Inst(b.Attention('core'))
  q, k, v = torch.randn(2, 3, 2, 8), torch.randn(2, 5, 2, 8), torch.randn(2, 5, 2, 8)
  bias = torch.zeros(2, 3, 5); bias[:, :, 4] = -1e9
  o = core.FProp(core.theta, q, k, v, bias)
  probs = torch.softmax(torch.einsum('BLHD,BMHD->BHLM', q, k)[..., :4], -1)
  torch.testing.assert_close(o, torch.einsum('BHLM,BMHD->BLHD', probs, v[:, :4]),
                             atol=1e-5, rtol=1e-5)


def test_encoder_decoder_stack_with_cross_attention_and_decode():
  b = _Builder()
  stack_p = b.DecoderLayerStack(
      'dec', [b.DecSelfAttentionRelativeBias('self'), b.DecEncAttention('cross'),
              b.DenseReluDense('ffn')], num=2)
  stack = _Inst(stack_p)
  bsz, l = 2, 5
  x = torch.randn(bsz, l, 16)
  seg = torch.ones(bsz, l, dtype=torch.long)
  pos = torch.arange(l).unsqueeze(0).expand(bsz, l)
  enc = torch.randn(bsz, 4, 16)
  enc_seg = torch.ones(bsz, 4, dtype=torch.long)
  out = stack.FProp(stack.theta, NestedMap(
      vec=x, segment_id=seg, segment_pos=pos, aux_loss=torch.zeros(()),
      encoder_output=enc, encoder_segment_id=enc_seg))
  assert out.vec.shape == (bsz, l, 16)
  state = stack.InitDecodeState(bsz, l, x.device, x.dtype)
  enc_state = stack.InitEncoderState(stack.theta, enc, enc_seg)
  steps = [stack.ExtendStep(stack.theta, x[:, t:t + 1], state, t, encoder_state=enc_state)
           for t in range(l)]
  torch.testing.assert_close(torch.cat(steps, 1), out.vec, atol=2e-4, rtol=2e-4)
  out.vec.sum().backward()
  assert all(v.grad is not None for v in stack.vars.Flatten())


def test_multi_dconv_head_attention_layer():
  b = _Builder(mdha_rope=True)
  mdha = _Inst(b.DecMultiDconvHeadAttentionRelativeBias('mdha'))
  plain = _Inst(b.DecSelfAttentionRelativeBias('plain'))
  names = {v.var_name.split('/', 1)[1] for v in mdha.vars.Flatten()}
  assert {'q_dconv/w_0/var', 'k_dconv/w_2/var', 'v_dconv/w_1/var'} <= names
  x, seg, pos = _Inputs()
  y, _ = mdha.FProp(mdha.theta, x, seg, pos)
  assert y.shape == x.shape and torch.isfinite(y).all()
  # causal: the future does not change the past; packing: other segments neither
  x2 = x.clone(); x2[1, 4:] += 1.0; x2[0, 3:5] += 1.0
  y2, _ = mdha.FProp(mdha.theta, x2, seg, pos)
  torch.testing.assert_close(y[1, :4], y2[1, :4], atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(y[0, :3], y2[0, :3], atol=1e-5, rtol=1e-5)
  assert mdha.vars.Flatten().__len__() == plain.vars.Flatten().__len__() + 9


def test_parallel_attention_ffn_block_and_encoder_residual_weight():
  b = _Builder()
  par = _Inst(b.ParallelDecSelfAttentionRelativeBiasFFN('par', 'gelu', gated=True))
  x, seg, pos = _Inputs()
  y, aux = par.FProp(par.theta, x, seg, pos)
  a, _ = par.atten.FProp(par.theta.atten, x, seg, pos)
  f, _ = par.ffn.FProp(par.theta.ffn, x, seg, pos)
  torch.testing.assert_close(y, a + f)
  blk_p = b.EncoderLayer('enc', b.DenseReluDenseGatedSILU('ffn'), residual_weight=0.5)
  blk = _Inst(blk_p)
  inp = NestedMap(vec=x, segment_id=seg, segment_pos=pos, aux_loss=torch.zeros(()))
  o = blk.FProp(blk.theta, inp)
  full = _Inst(b.EncoderLayer('enc', b.DenseReluDenseGatedSILU('ffn')))
  o1 = full.FProp(full.theta, inp)
  mask = (seg != 0).unsqueeze(-1).float()
  torch.testing.assert_close(o.vec - x * mask, 0.5 * (o1.vec - x * mask), atol=1e-5, rtol=1e-5)
  assert b.DenseReluDenseGatedGELU('g').activation == 'gelu'


def test_smoothed_softmax_repeat_and_misc_builders():
  b = _Builder()
  sm = _Inst(b.SmoothedSoftmax('sm', 11, label_smoothing=0.1))
  vec = torch.randn(2, 3, 16)
  ids = torch.randint(0, 11, (2, 3))
  out = sm.FProp(sm.theta, vec, ids, torch.tensor([[1., 1, 0], [1, 1, 1]]))
  logp = torch.log_softmax(vec @ sm.vars.w, -1)
  tgt = torch.full((2, 3, 11), 0.1 / 10)
  tgt.scatter_(-1, ids.unsqueeze(-1), 0.9)
  torch.testing.assert_close(out.per_token_loss, -(tgt * logp).sum(-1), atol=1e-5, rtol=1e-5)
  assert out.logits.shape == (2, 3, 11)
  w = _Inst(b.SoftmaxWeight('w', 11))
  assert w.Logits(w.theta, vec).shape == (2, 3, 11)
  masked = _Inst(b.Mask().Set(name='m')).FProp(NestedMap(), vec, torch.tensor([[1, 0, 1], [0, 0, 2]]))
  assert masked[0, 1].abs().sum() == 0 and masked[1, 2].abs().sum() > 0
  rep = _Inst(b.Repeat('rep', b._LN('ln'), repeat=3))
  assert rep.FProp(rep.theta, vec).shape == vec.shape
  assert len(rep.vars.Flatten()) == 3                      # per-layer variables
  p = gb.DenseBuilder.Params()
  gb.DenseBuilder.SetFPropDtype(p, torch.bfloat16)
  assert p.fprop_dtype == torch.bfloat16 and p.attention_logits_dtype == torch.float32
  assert b.LN('x').cls is gb.RmsNormLayer and b.PN('y').kind == 'pn'
  assert _Inst(b.Split('s')).FProp(NestedMap(), vec) is vec


def test_multi_dconv_head_attention_incremental_decode_matches_full():
  b = _Builder(mdha_rope=True)
  stack = _Inst(b.DecoderLayerStack(
      'dec', [b.DecMultiDconvHeadAttentionRelativeBias('mdha'), b.DenseReluDense('ffn')], num=2))
  bsz, l = 2, 6
  x = torch.randn(bsz, l, 16)
  seg = torch.ones(bsz, l, dtype=torch.long)
  pos = torch.arange(l).unsqueeze(0).expand(bsz, l)
  out = stack.FProp(stack.theta, NestedMap(vec=x, segment_id=seg, segment_pos=pos,
                                           aux_loss=torch.zeros(())))
  state = stack.InitDecodeState(bsz, l, x.device, x.dtype)
  steps = [stack.ExtendStep(stack.theta, x[:, t:t + 1], state, t) for t in range(l)]
  torch.testing.assert_close(torch.cat(steps, 1), out.vec, atol=2e-4, rtol=2e-4)
