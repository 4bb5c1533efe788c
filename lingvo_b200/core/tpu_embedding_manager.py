"""Table registry (ref `lingvo/core/tpu_embedding_manager.py`): collects every
`TPUEmbeddingLayer` of a model so the trainer can apply the sparse table updates once
per step (the reference's send-gradients / load / retrieve ops)."""
from lingvo_b200.core import tpu_embedding_layers


class TpuEmbeddingManager:

  def __init__(self, model=None):
    self._layers = []
    if model is not None:
      self.Collect(model)

  def Collect(self, model):
    self._layers = [l for _, l in model.Walk()
                    if isinstance(l, tpu_embedding_layers.TPUEmbeddingLayer)]
    return self._layers

  @property
  def enabled(self):
    return bool(self._layers)

  def ApplyGradients(self, global_step=0):
    for l in self._layers:
      l.ApplyGradients(global_step)
