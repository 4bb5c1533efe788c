"""Embedding-table manager (ref `lingvo/core/tpu_embedding_manager.py`).

The reference's `TPUEmbeddingManager` owns the TF2 `TPUEmbeddingV2` object: it is configured
with every table of the model, enqueues the sparse ids of each step, hands the activations to
the embedding layers and applies the table gradients after the backward pass. Here it is the
single object a training loop talks to for sparse tables:

    mgr = tpu_embedding_manager.TpuEmbeddingManager(model)   # or Default() + v2 layers
    …forward / backward…
    mgr.ApplyGradients(global_step)                           # sparse updates, all tables

It also reports the exchange plan (`Describe()`): rows per rank, features per table — what the
reference prints when it builds the embedding configuration.
"""

from __future__ import annotations

from lingvo_b200.core import tpu_embedding_layers
from lingvo_b200.core.nested_map import NestedMap

_DEFAULT = None


class TpuEmbeddingManager:

  def __init__(self, model=None):
    self._layers = []
    self._steps_applied = 0
    if model is not None:
      self.Collect(model)

  def Collect(self, model):
    """Finds every embedding layer of `model` (in addition to registered ones)."""
    found = [l for _, l in model.Walk()
             if isinstance(l, tpu_embedding_layers.TPUEmbeddingLayer)]
    for l in found:
      self.Register(l)
    return list(self._layers)

  def Register(self, layer):
    if all(layer is not l for l in self._layers):
      self._layers.append(layer)

  def Reset(self):
    self._layers = []
    self._steps_applied = 0

  @property
  def enabled(self):
    return bool(self._layers)

  @property
  def layers(self):
    return list(self._layers)

  @property
  def steps_applied(self):
    return self._steps_applied

  def Describe(self):
    """Per-table plan: vocabulary, dim, rows held by this rank, served features."""
    out = []
    for layer in self._layers:
      for t in layer.tables:
        p = t.params
        out.append(NestedMap(layer=layer.params.name, table=p.name, vocab_size=p.vocab_size,
                             dim=p.embedding_dim, local_rows=t._local_rows,   # pylint: disable=protected-access
                             features=list(p.input_keys), combiner=p.combiner))
    return out

  def ApplyGradients(self, global_step=0):
    for l in self._layers:
      l.ApplyGradients(global_step)
    self._steps_applied += 1


TPUEmbeddingManager = TpuEmbeddingManager      # reference spelling


def Default() -> TpuEmbeddingManager:
  """The process-wide manager v2 layers register with."""
  global _DEFAULT
  if _DEFAULT is None:
    _DEFAULT = TpuEmbeddingManager()
  return _DEFAULT
