"""Self-attention stacks built with the Builder DSL (ref `lingvo/core/self_attention_layer.py`):
`Builder` (:294) / `SimplifiedTransformerBuilder` (:428) and
`StackedTransformerEncoderLayers` (:696)."""

from __future__ import annotations

from lingvo_b200.core import base_layer
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core.nested_map import NestedMap


class Builder(bma.Builder):
  """Encoder-stack builder: pre-LN self-attention + FFN blocks."""

  def TransformerStack(self, name, num_layers=1, feed_forward_qdomain=None):
    del feed_forward_qdomain
    return self.TransformerEncoderStack(name, num_layers)

  def TransformerStackV2(self, name, num_layers=1, *, final_layer_first_n=None,
                         final_layer_stride=1, feed_forward_qdomain=None):
    del final_layer_first_n, final_layer_stride, feed_forward_qdomain
    return self.TransformerEncoderStack(name, num_layers)


class SimplifiedTransformerBuilder(Builder):
  """Parallel attention+FFN blocks without LayerNorm on the skip path (arXiv 2311.01906)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('parallel_attention_mlp', True, 'Attention and MLP share one residual.')
    p.atten_tpl = bma.MultiHeadedAttention.Params().Set(enable_shaped_attention=True)
    return p


class StackedTransformerEncoderLayers(base_layer.BaseLayer):
  """`Builder.TransformerStack` behind the `(vec, paddings) → (vec, paddings)` interface."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('builder', Builder.Params(), 'Builder params.')
    p.Define('num_layers', 1, 'Layers.')
    p.Define('mdl_dim', 0, 'Model dim.')
    p.Define('hidden_dim', 0, 'FFN hidden dim.')
    p.Define('num_atten_heads', 0, 'Heads.')
    p.Define('dropout_prob', 0.0, 'Dropout.')
    return p

  @classmethod
  def Cast(cls, params):
    """Converts `StackedTransformerLayers` params into this class's params (ref :700)."""
    p = cls.Params()
    for k in ('name', 'num_layers', 'mdl_dim', 'hidden_dim', 'num_atten_heads',
              'dropout_prob'):
      if k in params:
        p.Set(**{k: params.Get(k)})
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    b = p.builder.Copy().Set(model_dim=p.mdl_dim, ff_hidden_dim=p.hidden_dim,
                             num_heads=p.num_atten_heads,
                             residual_dropout_prob=p.dropout_prob).Instantiate()
    self.CreateChild('stack', b.TransformerStack('stack', p.num_layers))

  def FProp(self, theta, vec, paddings, segment_mask=None):
    i = NestedMap(vec=vec, paddings=paddings)
    if segment_mask is not None:
      i.segment_mask = segment_mask
    out = self.stack.FProp(theta.stack, i)
    return out.vec, out.paddings
