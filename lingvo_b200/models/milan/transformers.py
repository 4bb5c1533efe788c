"""Text-side encoders (ref `lingvo/tasks/milan/transformers.py`).

`GetTransformerStackWithEmbeddingInput` builds the "BERT adapter": pre-computed token
embeddings `[B, T, input_dim]` + lengths `[B]` → projection → N Transformer layers →
the first position's output `[B, output_dim]`.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import layers
from lingvo_b200.models.milan import utils
from lingvo_b200.models.mt import layers as mt_layers


class EmbeddingSequenceEncoder(base_layer.BaseLayer):
  """(features [B,T,D_in], lengths [B]) → [B, D_out]."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_projection', layers.ProjectionLayer.Params(), 'Input projection.')
    p.Define('transformer_stack', mt_layers.TransformerStack.Params(), 'Stack.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('input_projection', p.input_projection)
    self.CreateChild('transformer_stack', p.transformer_stack)

  def FProp(self, theta, features, lengths):
    assert features.dim() == 3 and lengths.dim() == 1
    t = features.shape[1]
    paddings = (torch.arange(t, device=features.device).unsqueeze(0) >=
                lengths.reshape(-1, 1)).to(features.dtype)
    x = self.input_projection.FProp(theta.input_projection, features)
    out, _, _ = self.transformer_stack.FProp(
        theta.transformer_stack, utils.BatchMajorToTimeMajor(x),
        utils.BatchMajorToTimeMajor(paddings))
    return out[0]


def GetTransformerStackWithEmbeddingInput(*, input_dim, num_layers, hidden_dim,
                                          num_attention_heads, output_dim, name=''):
  """Params of the adapter encoder (ref :25)."""
  stack = mt_layers.TransformerStack.Params().Set(
      name='transformer_stack', model_dim=output_dim, num_transformer_layers=num_layers)
  stack.transformer_tpl.tr_fflayer_tpl.hidden_dim = hidden_dim
  stack.transformer_tpl.tr_atten_tpl.num_attention_heads = num_attention_heads
  return EmbeddingSequenceEncoder.Params().Set(
      name=name or 'embedding_sequence_encoder',
      input_projection=layers.ProjectionLayer.Params().Set(
          name='input_projection', has_bias=True, batch_norm=False, input_dim=input_dim,
          output_dim=output_dim, activation='NONE'),
      transformer_stack=stack)
