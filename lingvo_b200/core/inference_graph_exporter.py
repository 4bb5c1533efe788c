"""Inference export (ref `lingvo/core/inference_graph_exporter.py`).

The reference freezes a TF `GraphDef` plus `subgraphs{name: feeds, fetches}`. The
B200 equivalent of "a frozen graph" is a self-contained **inference bundle**:

  <dir>/inference_graph.json   model name, task, subgraph names + their feed/fetch
                               keys, dtype policy, bundle format version
  <dir>/params.txt             the full model Params (`ToText`) – rebuilds the layers
  <dir>/weights.pt             {var_name: tensor} (optionally bf16, EMA applied)

`Predictor` (predictor.py) rebuilds the model from `params.txt`, loads the weights
and serves the subgraphs returned by `task.Inference()`; static-shape subgraphs are
captured into CUDA graphs on first use.
"""

from __future__ import annotations

import collections
import inspect
import json
import os
from typing import Dict, Optional

import torch

from lingvo_b200.core import hyperparams
from lingvo_b200.core import py_utils

BUNDLE_VERSION = 1
InferenceDeviceOptions = collections.namedtuple(
    'InferenceDeviceOptions', ['device', 'retain_device_placement', 'var_options',
                               'gen_init_op', 'dtype_override', 'fprop_dtype_override'])
InferenceDeviceOptions.__new__.__defaults__ = ('', False, None, True, None, None)


class InferenceGraph:
  """In-memory description of an exported bundle (the `InferenceGraph` proto's role)."""

  def __init__(self, model_name='', task_name='', subgraphs=None, asset_dir=''):
    self.model_name = model_name
    self.task_name = task_name
    self.subgraphs = subgraphs or {}      # name → {'feeds': [...], 'fetches': [...]}
    self.asset_dir = asset_dir

  def ToJson(self):
    return json.dumps({'version': BUNDLE_VERSION, 'model_name': self.model_name,
                       'task_name': self.task_name, 'subgraphs': self.subgraphs}, indent=1)

  @classmethod
  def FromJson(cls, text, asset_dir=''):
    d = json.loads(text)
    return cls(d.get('model_name', ''), d.get('task_name', ''), d.get('subgraphs', {}),
               asset_dir)


def _SubgraphSpec(fn):
  """feeds = the callable's argument names; fetches are discovered at first run."""
  try:
    feeds = [p.name for p in inspect.signature(fn).parameters.values()
             if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]
  except (TypeError, ValueError):
    feeds = []
  return {'feeds': feeds, 'fetches': []}


class InferenceGraphExporter:

  @classmethod
  def Export(cls, model_cfg, model_task_name=None, device_options=InferenceDeviceOptions(),
             freeze_checkpoint=None, freeze_defaults=False, export_path=None,
             subgraph_filter=None, random_seed=None, disable_packed_input=True,
             prune_graph=True, export_graph_collections=False, model=None) -> InferenceGraph:
    """Builds the model (or uses `model`), optionally loads `freeze_checkpoint`, and
    writes the bundle to `export_path` (a directory). Returns the InferenceGraph."""
    del freeze_defaults, prune_graph, export_graph_collections
    if random_seed is not None:
      torch.manual_seed(random_seed)
    cfg = model_cfg.Copy()
    if disable_packed_input:
      def _Off(p):
        if isinstance(p, hyperparams.Params) and 'packed_input' in p:
          p.packed_input = False
      _Off(cfg.task) if 'task' in cfg and cfg.task is not None else None
    cfg.is_inference = True
    if model is None:
      model = cfg.Instantiate()
    task = model.GetTask(model_task_name) if model_task_name else model.tasks[0]
    if freeze_checkpoint:
      from lingvo_b200.core import checkpointer  # pylint: disable=g-import-not-at-top
      checkpointer.Checkpointer(os.path.dirname(freeze_checkpoint), model).RestoreFromPath(
          checkpoint_path=freeze_checkpoint)
    subgraphs = task.Inference()
    if subgraph_filter:
      keep = set(subgraph_filter if isinstance(subgraph_filter, (list, tuple))
                 else [subgraph_filter])
      subgraphs = {k: v for k, v in subgraphs.items() if k in keep}
    graph = InferenceGraph(
        model_name=str(model_cfg.get('model', '') if hasattr(model_cfg, 'get') else ''),
        task_name=model_task_name or '', subgraphs={k: _SubgraphSpec(v)
                                                    for k, v in subgraphs.items()})
    if export_path:
      os.makedirs(export_path, exist_ok=True)
      dtype = device_options.dtype_override
      weights = {}
      for v in model.vars.Flatten():
        t = v.detach().cpu()
        if dtype is not None and t.is_floating_point():
          t = t.to(dtype)
        weights[v.var_name] = t
      torch.save(weights, os.path.join(export_path, 'weights.pt'))
      with open(os.path.join(export_path, 'params.txt'), 'w') as f:
        f.write(model_cfg.ToText())
      with open(os.path.join(export_path, 'inference_graph.json'), 'w') as f:
        f.write(graph.ToJson())
      graph.asset_dir = export_path
    return graph


def LoadInferenceGraph(path) -> InferenceGraph:
  d = path if os.path.isdir(path) else os.path.dirname(path)
  with open(os.path.join(d, 'inference_graph.json')) as f:
    return InferenceGraph.FromJson(f.read(), d)
