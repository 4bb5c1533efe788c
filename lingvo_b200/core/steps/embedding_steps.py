"""Embedding steps (ref `lingvo/core/steps/embedding_steps.py`)."""
import torch

from lingvo_b200.core import layers
from lingvo_b200.core import step
from lingvo_b200.core.nested_map import NestedMap


class EmbeddingStep(step.Step):
  """ids → embeddings (ref :25)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.EmbeddingLayer.Params().Set(max_num_shards=1), 'Embedding.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('emb', self.params.emb)

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    ids = step_inputs.inputs[0].long()
    return NestedMap(output=self.emb.EmbLookup(theta.emb, ids)), state0


class StatefulEmbeddingStep(step.Step):
  """Embedding + learned position embedding with the position kept in state (ref :70)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('target_vocab_size', 0, 'Vocab size.')
    p.Define('emb', layers.EmbeddingLayer.Params().Set(max_num_shards=1), 'Token embedding.')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('embedding_dim', 0, 'Embedding dim.')
    p.Define('num_prev_tokens', 0, 'Previous tokens kept in state (n-gram context).')
    p.Define('include_current_token', True, 'Embed the current token too.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('emb', p.emb.Copy().Set(vocab_size=p.target_vocab_size,
                                             embedding_dim=p.embedding_dim))
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.embedding_dim))

  def ZeroState(self, theta, prepared_inputs, batch_size):
    p = self.params
    return NestedMap(t=0, prev=torch.zeros(batch_size, max(p.num_prev_tokens, 0),
                                           dtype=torch.int64, device=self.Device()))

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    p = self.params
    ids = step_inputs.inputs[0].long().reshape(-1, 1)
    toks = torch.cat([state0.prev, ids], 1) if p.num_prev_tokens else ids
    if not p.include_current_token:
      toks = toks[:, :-1]
    emb = self.emb.EmbLookup(theta.emb, toks).sum(1)
    pos = self.position_emb.FProp(theta.position_emb, state0.t + 1)[state0.t]
    prev = toks[:, -p.num_prev_tokens:] if p.num_prev_tokens else state0.prev
    return NestedMap(output=emb + pos.to(emb.dtype)), NestedMap(t=state0.t + 1, prev=prev)
