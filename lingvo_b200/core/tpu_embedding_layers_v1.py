"""The v1 embedding-table API (ref `lingvo/core/tpu_embedding_layers_v1.py`).

The reference's v1 layers talk to the TF1 `TPUEmbedding` mid-level API through a per-graph
singleton, `TpuEmbeddingCollection`, that the trainer programs query: which tables exist,
which features they serve, the activations of the current task scope, whether a task should
stop gradients into the tables (`SetTaskMode` — eval / decode programs sharing the tables),
the gradient-multiplier schedule, and `ApplyGradients` to push feature gradients to the
tables. Here the tables are the row-sharded sparse tables of `tpu_embedding_layers.py`; this
module provides that collection-centred surface on top of them.
"""

from __future__ import annotations

from typing import AbstractSet, Dict, List, Mapping, Sequence, Tuple

import torch

from lingvo_b200.core import tpu_embedding_layers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.core.tpu_embedding_layers import (  # noqa: F401  (v1 names)
    TPUEmbeddingAdagradOptimizer, TPUEmbeddingAdamOptimizer, TPUEmbeddingFTRLOptimizer,
    TPUEmbeddingSGDOptimizer)


class TpuEmbeddingCollection:
  """Process-wide registry of embedding tables and per-task activations (ref :45)."""

  _INSTANCE = None

  @classmethod
  def Get(cls) -> 'TpuEmbeddingCollection':
    if cls._INSTANCE is None:
      cls._INSTANCE = cls()
    return cls._INSTANCE

  @classmethod
  def Reset(cls):
    cls._INSTANCE = None

  def __init__(self):
    self._table_vars = NestedMap()
    self._bf16_inference_vars: List[str] = []
    self._layers: List['TPUEmbeddingLayer'] = []
    self._activations_by_task: Dict[str, Mapping[str, torch.Tensor]] = {}
    self._summary_tensors: List[Tuple[str, torch.Tensor, torch.Tensor]] = []
    self._feature_names: AbstractSet[str] = frozenset()
    self._gradient_multiplier_schedule = None
    self._task_mode: Dict[str, str] = {}

  # -- tables ---------------------------------------------------------------------
  def AddTableVariables(self, table_name, var_list, is_inference_with_bfloat16=False):
    if table_name in self._table_vars:
      raise ValueError('Variables for table %s already exist.' % table_name)
    self._table_vars[table_name] = list(var_list)
    if is_inference_with_bfloat16:
      self._bf16_inference_vars.append(table_name)

  @property
  def table_variables(self) -> NestedMap:
    return self._table_vars

  @property
  def inference_with_bfloat16_var_names(self):
    return list(self._bf16_inference_vars)

  def AddLayer(self, layer):
    self._layers.append(layer)
    self.feature_names = frozenset(layer.feature_names)

  @property
  def layers(self):
    return list(self._layers)

  # -- features / activations -------------------------------------------------------
  @property
  def feature_names(self) -> AbstractSet[str]:
    return self._feature_names

  @feature_names.setter
  def feature_names(self, names: AbstractSet[str]):
    names = frozenset(names)
    if self._feature_names and self._feature_names != names:
      raise ValueError('feature_names already set to %s, cannot change to %s' %
                       (sorted(self._feature_names), sorted(names)))
    self._feature_names = names

  def AddActivations(self, task_call_scope: str, activations: Mapping[str, torch.Tensor]):
    if task_call_scope in self._activations_by_task:
      raise ValueError('Activations for task %s already exist.' % task_call_scope)
    self._activations_by_task[task_call_scope] = activations
    return activations

  def GetActivations(self, task_call_scope: str):
    return self._activations_by_task.get(task_call_scope)

  def ClearActivations(self):
    self._activations_by_task = {}

  def AddSummaryTensor(self, name, value, weight=1.0):
    self._summary_tensors.append((name, value, torch.as_tensor(weight)))

  @property
  def summary_tensors(self) -> Sequence[Tuple[str, torch.Tensor, torch.Tensor]]:
    return list(self._summary_tensors)

  # -- gradients ----------------------------------------------------------------------
  def SetGradientMultiplierSchedule(self, multiplier_schedule):
    if self._gradient_multiplier_schedule is not None:
      raise ValueError('gradient_multiplier_schedule was set before.')
    self._gradient_multiplier_schedule = multiplier_schedule

  def SetTaskMode(self, task_call_scope, mode):
    """`mode` ∈ {'train', 'eval', 'decode', …}: only 'train' tasks update the tables."""
    self._task_mode[task_call_scope] = mode

  def ShouldStopGradient(self, task_call_scope):
    if task_call_scope not in self._task_mode:
      raise ValueError('Called ShouldStopGradient for unknown task %s.' % task_call_scope)
    return self._task_mode[task_call_scope] != 'train'

  def ApplyGradients(self, task_call_scope, feature_to_gradient_dict=None, global_step=0):
    """Applies the queued sparse table updates of a training task (no-op for tasks whose mode
    stops gradients). `feature_to_gradient_dict` is accepted for API parity: gradients reach
    the tables through autograd, not through an explicit send op."""
    del feature_to_gradient_dict
    if task_call_scope in self._task_mode and self.ShouldStopGradient(task_call_scope):
      for layer in self._layers:
        layer.DiscardGradients()
      return 0
    mult = None
    if self._gradient_multiplier_schedule is not None:
      mult = float(self._gradient_multiplier_schedule.Value(global_step))
    n = 0
    for layer in self._layers:
      layer.ApplyGradients(global_step, multiplier=mult)
      n += 1
    return n


class TPUEmbeddingTable(tpu_embedding_layers.TPUEmbeddingTable):
  """v1 table: registers its variables in the collection (ref :655)."""

  def _InstantiateSelfAndChildren(self):
    super()._InstantiateSelfAndChildren()
    coll = TpuEmbeddingCollection.Get()
    if self.params.name not in coll.table_variables:
      coll.AddTableVariables(self.params.name, [self.vars['var']],
                             self.params.inference_variable_dtype == torch.bfloat16)

  def DiscardGradients(self):
    self._pending = []
    self._pending_grads = []


class TPUEmbeddingLayer(tpu_embedding_layers.TPUEmbeddingLayer):
  """v1 layer: tables + routing + the collection protocol (ref :836)."""

  def __init__(self, params):
    p = params.Copy()
    p.tables = [tp.Copy().Set(cls=TPUEmbeddingTable) if tp.cls is
                tpu_embedding_layers.TPUEmbeddingTable else tp for tp in p.tables]
    super().__init__(p)
    coll = TpuEmbeddingCollection.Get()
    coll.AddLayer(self)
    if self.params.gradient_multiplier_schedule is not None:
      try:
        coll.SetGradientMultiplierSchedule(self.gradient_multiplier_schedule)
      except ValueError:
        pass

  @property
  def feature_names(self):
    return sorted(self._route)

  def EmbLookup(self, theta, ids_map, task_call_scope='default'):
    out = super().EmbLookup(theta, ids_map)
    coll = TpuEmbeddingCollection.Get()
    coll._activations_by_task[task_call_scope] = out   # pylint: disable=protected-access
    return out

  def DiscardGradients(self):
    for t in self.tables:
      if hasattr(t, 'DiscardGradients'):
        t.DiscardGradients()
      else:
        t._pending, t._pending_grads = [], []   # pylint: disable=protected-access

  def ApplyGradients(self, global_step=0, multiplier=None):
    if multiplier is None:
      return super().ApplyGradients(global_step)
    for t in self.tables:
      t.gradient_multiplier = multiplier
      t.ApplyGradients(global_step)
