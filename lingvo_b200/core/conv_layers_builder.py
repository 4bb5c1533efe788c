"""Builder for padding-aware conv stacks (ref `lingvo/core/conv_layers_builder.py`).

Every block consumes and produces `(activations [B,T,F,C], paddings [B,T])`;
`Conv2D`, `DepthwiseConv2D`, `SeparableConv2D`, `NormalizedDepthwiseConv2D`,
`CausalPooling` are composed with `_Seq`/`_Graph` from `builder.Base`."""

from __future__ import annotations

import torch

from lingvo_b200.core import activations
from lingvo_b200.core import base_layer
from lingvo_b200.core import bn_layers
from lingvo_b200.core import builder
from lingvo_b200.core import builder_layers
from lingvo_b200.core import conv_layers_with_time_padding as conv_lib
from lingvo_b200.core import layers


class BiasLayer(builder_layers.BiasLayer):
  """Bias that passes paddings through (ref :42)."""

  def FProp(self, theta, inputs, paddings=None):
    out = super().FProp(theta, inputs)
    return out if paddings is None else (out, paddings)


class CausalPoolingLayer(base_layer.BaseLayer):
  """Pooling over the `left_context` most recent frames (ref :49)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('pooling_type', 'AVG', 'AVG | MAX.')
    p.Define('left_context', None, 'Frames incl. the current one (-1: all the past).')
    return p

  def FProp(self, theta, inputs, paddings):
    p = self.params
    b, t = inputs.shape[:2]
    w = t if p.left_context in (None, -1) else p.left_context
    mask = (1.0 - paddings.float()).view(b, t, *([1] * (inputs.dim() - 2)))
    x = inputs * mask.to(inputs.dtype)
    if p.pooling_type == 'AVG':
      cs = torch.cumsum(x, 1)
      cn = torch.cumsum(mask, 1)
      if w < t:
        cs = cs - torch.nn.functional.pad(cs, (0, 0) * (x.dim() - 2) + (w, 0))[:, :t]
        cn = cn - torch.nn.functional.pad(cn, (0, 0) * (x.dim() - 2) + (w, 0))[:, :t]
      out = cs / cn.clamp_min(1.0).to(cs.dtype)
    else:
      neg = torch.finfo(x.dtype).min
      xm = x.masked_fill(mask == 0, neg)
      xp = torch.nn.functional.pad(xm, (0, 0) * (x.dim() - 2) + (w - 1, 0), value=neg)
      out = xp.unfold(1, w, 1).amax(-1)
      out = torch.where(out == neg, torch.zeros_like(out), out)
    return out * mask.to(out.dtype), paddings


class Builder(builder.Base):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('norm_layer_tpl', bn_layers.BatchNormLayer.Params(), 'Norm after conv (None: bias).')
    p.Define('weight_norm', False, 'Weight normalisation.')
    p.Define('v2_padding', False, 'Correct strided padding.')
    return p

  def _Conv(self, cls, name, filter_shape, stride, dilation, is_causal):
    p = self.params
    tpl = cls.Params().Set(name=name, filter_shape=tuple(filter_shape),
                           filter_stride=tuple(stride or (1, 1)),
                           dilation_rate=tuple(dilation or (1, 1)), weight_norm=p.weight_norm,
                           v2_padding=p.v2_padding)
    if is_causal:
      tpl.is_causal = True
    return tpl

  def _NormOrBias(self, name, dims):
    p = self.params
    if p.norm_layer_tpl is None:
      return BiasLayer.Params().Set(name=name, dims=dims)
    tpl = p.norm_layer_tpl.Copy().Set(name=name)
    if 'dim' in tpl:
      tpl.dim = dims
    return _PaddedNorm.Params().Set(name=name, norm=tpl)

  def _Act(self, name, activation):
    return _PaddedFn.Params().Set(name=name, activation=activation)

  def Conv2D(self, name, filter_shape, stride=None, dilation_rate=None, activation='RELU',
             conv_last=False, is_causal=False):
    conv = self._Conv(conv_lib.Conv2DLayerWithPadding, 'conv_2d', filter_shape, stride,
                      dilation_rate, is_causal)
    norm = self._NormOrBias('normbias', filter_shape[3] if not conv_last else filter_shape[2])
    act = self._Act('act', activation)
    seq = [norm, act, conv] if conv_last else [conv, norm, act]
    return _PaddedSeq.Params().Set(name=name, sub=seq)

  def DepthwiseConv2D(self, name, filter_shape, stride=None, dilation_rate=None,
                      activation='RELU', conv_last=False, is_causal=False):
    conv = self._Conv(conv_lib.DepthwiseConv2DLayer, 'conv_2d', filter_shape, stride,
                      dilation_rate, is_causal)
    out_c = filter_shape[2] * filter_shape[3]
    norm = self._NormOrBias('normbias', out_c if not conv_last else filter_shape[2])
    act = self._Act('act', activation)
    seq = [norm, act, conv] if conv_last else [conv, norm, act]
    return _PaddedSeq.Params().Set(name=name, sub=seq)

  def SeparableConv2D(self, name, filter_shape, depth_multiplier=1, stride=None,
                      dilation_rate=None, activation='RELU', conv_last=False, is_causal=False):
    h, w, cin, cout = filter_shape
    dw = self._Conv(conv_lib.DepthwiseConv2DLayer, 'conv_2d_dw', (h, w, cin, depth_multiplier),
                    stride, dilation_rate, is_causal)
    pw = self._Conv(conv_lib.Conv2DLayerWithPadding, 'conv_2d_pw',
                    (1, 1, cin * depth_multiplier, cout), None, None, False)
    norm = self._NormOrBias('normbias', cout if not conv_last else cin)
    act = self._Act('act', activation)
    seq = [norm, act, dw, pw] if conv_last else [dw, pw, norm, act]
    return _PaddedSeq.Params().Set(name=name, sub=seq)

  def NormalizedDepthwiseConv2D(self, name, kernel_size, num_heads, in_dim, dropconnect_prob=0,
                                deterministic_dropout=False, is_causal=False):
    cls = conv_lib.CausalNormalizedDepthwiseConv2DLayer if is_causal else \
        conv_lib.NormalizedDepthwiseConv2DLayer
    return cls.Params().Set(name=name, filter_shape=(kernel_size, 1, num_heads, 1),
                            weight_tiling_factor=in_dim // num_heads,
                            dropconnect_prob=dropconnect_prob,
                            deterministic_dropout=deterministic_dropout)

  def CausalPooling(self, name, pooling_type='AVG', left_context=None):
    return CausalPoolingLayer.Params().Set(name=name, pooling_type=pooling_type,
                                           left_context=left_context)


class _PaddedSeq(base_layer.BaseLayer):
  """Sequential over `(x, paddings)` pairs."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'Layer params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChildren('sub', list(self.params.sub))

  def FProp(self, theta, x, paddings):
    for i, l in enumerate(self.sub):
      x, paddings = l.FProp(theta.sub[i], x, paddings)
    return x, paddings


class _PaddedNorm(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('norm', None, 'Norm params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('norm', self.params.norm)

  def FProp(self, theta, x, paddings):
    pad = paddings.view(paddings.shape[0], -1, *([1] * (x.dim() - 2)))
    n = self.norm
    if isinstance(n, bn_layers.GroupNormLayer):
      y = n.FProp(theta.norm, x, paddings)
      return (y[0] if isinstance(y, tuple) else y), paddings
    if isinstance(n, bn_layers.BatchNormLayer):
      return n.FProp(theta.norm, x, pad), paddings
    return n.FProp(theta.norm, x), paddings


class _PaddedFn(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('activation', 'RELU', 'Activation.')
    return p

  def FProp(self, theta, x, paddings):
    return activations.GetFn(self.params.activation)(x), paddings
