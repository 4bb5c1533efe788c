"""The v2 embedding-table API (ref `lingvo/core/tpu_embedding_layers_v2.py`).

In the reference, v2 binds the tables to the TF2 `TPUEmbeddingV2` object held by
`tpu_embedding_manager.TPUEmbeddingManager` (per-table optimizer objects are handed to the
mid-level API instead of slot variables being created by lingvo, and lookups dequeue
activations that the manager enqueued). The equivalent split here:

  * `TPUEmbedding*Optimizer` — the same sparse optimizers, plus `CreateOptimizerFn()` that
    returns a plain callable `(table, rows, grads, lr)` the manager can apply, mirroring the
    "optimizer object handed to the embedding API" role (ref `_TPUEmbeddingOptimizerV2Mixin`);
  * `TPUEmbeddingTable.GetDeviceName / table_config` — static description of a table
    (vocabulary, dim, combiner, owner rank of a row) the manager uses to plan the exchange;
  * `TPUEmbeddingLayer` — registers itself with the process-wide
    `tpu_embedding_manager.TpuEmbeddingManager` on construction, so `manager.ApplyGradients()`
    after backward updates every table of every v2 layer (the trainer's only hook).
"""

from __future__ import annotations

from lingvo_b200.core import tpu_embedding_layers
from lingvo_b200.core.nested_map import NestedMap


class _TPUEmbeddingOptimizerV2Mixin:
  """Adds the "optimizer as an object" surface of the v2 API (ref :57)."""

  def CreateOptimizerFn(self):
    """→ callable(table, slots, rows, grads, lr) applying one sparse update."""
    def Apply(table, slots, rows, grads, lr):
      self.Apply(lr, table, slots, rows, grads)
    return Apply

  def CreateSlotVariablesAndOps(self, table_vars, tpu_embedding_table):
    """v1-only in the reference; v2 optimizers own their slots."""
    raise NotImplementedError('The v2 embedding API creates optimizer slots itself.')


class TPUEmbeddingSGDOptimizer(_TPUEmbeddingOptimizerV2Mixin,
                               tpu_embedding_layers.TPUEmbeddingSGDOptimizer):
  pass


class TPUEmbeddingAdagradOptimizer(_TPUEmbeddingOptimizerV2Mixin,
                                   tpu_embedding_layers.TPUEmbeddingAdagradOptimizer):
  pass


class TPUEmbeddingAdamOptimizer(_TPUEmbeddingOptimizerV2Mixin,
                                tpu_embedding_layers.TPUEmbeddingAdamOptimizer):
  pass


class TPUEmbeddingFTRLOptimizer(_TPUEmbeddingOptimizerV2Mixin,
                                tpu_embedding_layers.TPUEmbeddingFTRLOptimizer):
  pass


class TPUEmbeddingTable(tpu_embedding_layers.TPUEmbeddingTable):
  """v2 table with a static config the manager plans with (ref :193)."""

  @property
  def table_config(self):
    p = self.params
    return NestedMap(name=p.name, vocabulary_size=p.vocab_size, dim=p.embedding_dim,
                     combiner=p.combiner, max_sequence_length=p.max_sequence_length,
                     features=list(p.input_keys), num_shards=self._world,
                     local_rows=self._local_rows)

  def OwnerOf(self, row: int) -> int:
    """Rank that stores global row `row` (row-sharded: `row % world`)."""
    return int(row) % self._world


class TPUEmbeddingLayer(tpu_embedding_layers.TPUEmbeddingLayer):
  """v2 layer: tables + routing, self-registered with the manager (ref :257)."""

  def __init__(self, params):
    p = params.Copy()
    p.tables = [tp.Copy().Set(cls=TPUEmbeddingTable) if tp.cls is
                tpu_embedding_layers.TPUEmbeddingTable else tp for tp in p.tables]
    super().__init__(p)
    from lingvo_b200.core import tpu_embedding_manager   # pylint: disable=g-import-not-at-top
    tpu_embedding_manager.Default().Register(self)

  @property
  def table_configs(self):
    return [t.table_config for t in self.tables]
