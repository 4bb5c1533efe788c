"""MT-specific layers (ref `lingvo/tasks/mt/layers.py`).

`TransformerStack` (ref :26): N time-major Transformer layers with optional final
layer norm, optional cross attention in every layer, and the *transparent* mode in
which the outputs of all layers (plus the embedding) are merged by learned softmax
weights — one merged tensor per decoder layer.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import layers
from lingvo_b200.core import layers_with_attention


class TransformerStack(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('model_dim', 1024, 'Model dim.')
    p.Define('num_transformer_layers', 6, 'Number of layers.')
    p.Define('transformer_tpl', layers_with_attention.TransformerLayer.Params(),
             'Layer template, or a list of templates applied round-robin.')
    p.Define('ln_tpl', layers.LayerNorm.Params(), 'Final layer-norm template.')
    p.Define('ln_output', False, 'Layer-normalise the stack output.')
    p.Define('is_transparent', False, 'Emit learned mergers of all layer outputs.')
    p.Define('num_transparent_outputs', 6, 'Number of merged outputs.')
    p.Define('transparent_merger_tpl',
             layers.WeightedSumLayer.Params().Set(add_weight_summaries=True), 'Merger tpl.')
    p.Define('packed_input', False, 'Packed inputs (segment ids).')
    p.Define('has_aux_attention', False, 'Every layer also attends to aux vectors.')
    p.Define('mask_self_atten', False, 'Causal self attention.')
    p.transformer_tpl.tr_atten_tpl.num_attention_heads = 8
    p.transformer_tpl.tr_fflayer_tpl.hidden_dim = 8192
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    tpls = p.transformer_tpl if isinstance(p.transformer_tpl, (list, tuple)) else [
        p.transformer_tpl]
    assert p.num_transformer_layers % len(tpls) == 0
    ps = []
    for i in range(p.num_transformer_layers):
      ps.append(tpls[i % len(tpls)].Copy().Set(
          name='trans_%d' % i, source_dim=p.model_dim, packed_input=p.packed_input,
          has_aux_atten=p.has_aux_attention, mask_self_atten=p.mask_self_atten))
    self.CreateChildren('trans', ps)
    if p.ln_output:
      self.CreateChild('layer_norm_out', p.ln_tpl.Copy().Set(
          name='enc_out_ln', input_dim=p.model_dim))
    if p.is_transparent:
      if not p.num_transparent_outputs:
        raise ValueError('num_transparent_outputs should be greater than 0.')
      self.CreateChildren('transparent_merger', [
          p.transparent_merger_tpl.Copy().Set(
              name='transparent_%d' % i, num_sources=1 + p.num_transformer_layers)
          for i in range(p.num_transparent_outputs)])

  def FProp(self, theta, transformer_input, paddings, src_segment_id=None, aux_vecs=None,
            aux_paddings=None, aux_segment_id=None, return_atten_probs=False):
    """[T,B,D] → (outputs, paddings, segment_id). Transparent mode returns a list of
    `[T,B,D]` tensors in training and one stacked `[T,B,D,K]` tensor in eval."""
    p = self.params
    if p.packed_input:
      assert src_segment_id is not None, 'packed_input needs src_segment_id'
    x = transformer_input
    outs, probs = [x], None
    for i, layer in enumerate(self.trans):
      x, probs = layer.FProp(theta.trans[i], x, paddings, aux_vecs=aux_vecs,
                             aux_paddings=aux_paddings, source_segment_id=src_segment_id,
                             aux_segment_id=aux_segment_id)
      outs.append(x)
    if p.ln_output:
      x = self.layer_norm_out.FProp(theta.layer_norm_out, x)
      outs[-1] = x
    if p.is_transparent:
      merged = [m.FProp(theta.transparent_merger[i], outs)
                for i, m in enumerate(self.transparent_merger)]
      x = torch.stack(merged, -1) if self.do_eval else merged
    seg = src_segment_id if p.packed_input else None
    if return_atten_probs:
      return x, paddings, seg, probs
    return x, paddings, seg
