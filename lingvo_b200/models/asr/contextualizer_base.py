"""Context injection API for the speech decoder (ref
`lingvo/tasks/asr/contextualizer_base.py`).

A contextualizer owns a set of biasing phrases (the *context map*) and exposes an
attention-like interface the decoder queries each step; its output is concatenated to the
audio attention context. `NullContextualizer` contributes nothing (dim 0).
"""

from __future__ import annotations

from lingvo_b200.core import base_layer


class ContextualizerBase(base_layer.BaseLayer):
  """Interface (ref :19)."""

  def SetContextMap(self, context_map, theta):
    """Encodes / stores the biasing phrases for the current batch."""
    raise NotImplementedError()

  def InitAttention(self, theta, packed_src, misc_states):
    """Pre-computes per-utterance attention sources."""
    raise NotImplementedError()

  def ZeroAttention(self, theta, dec_bs, misc_states, audio_context, packed_src):
    """Initial (context vector, state) before the first decode step."""
    raise NotImplementedError()

  def QueryAttention(self, theta, attn_query, misc_states, audio_context, packed_src):
    """One attention step → context vector [dec_bs, GetContextDim()]."""
    raise NotImplementedError()

  def GetContextDim(self):
    raise NotImplementedError()


class NullContextualizer(ContextualizerBase):
  """No biasing (ref :107)."""

  def SetContextMap(self, context_map, theta):
    return None

  def InitAttention(self, theta, packed_src, misc_states=None):
    return None

  def ZeroAttention(self, theta, dec_bs, misc_states, audio_context, packed_src):
    return audio_context

  def QueryAttention(self, theta, attn_query, misc_states, audio_context, packed_src):
    return audio_context

  def GetContextDim(self):
    return 0
