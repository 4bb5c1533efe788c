"""Predictor: serve the subgraphs of an exported inference bundle
(ref `lingvo/core/predictor.py:40-395`).

    pred = Predictor(inference_graph=<bundle dir or InferenceGraph>, device='cuda:0')
    out = pred.Run(['log_pplx_per_token'], subgraph_name='default', ids=..., paddings=...)

Feeds are keyword arguments named after the subgraph callable's parameters; fetches
select keys of the returned NestedMap (`None` → everything). With `use_cuda_graph`
a subgraph is captured per distinct input-shape signature and replayed (static
buffers are refilled with `copy_`), which removes per-call launch overhead.
"""

from __future__ import annotations

import os
import threading
from typing import Dict, List, Optional

import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import hyperparams
from lingvo_b200.core import inference_graph_exporter
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def LoadInferenceGraph(path):
  return inference_graph_exporter.LoadInferenceGraph(path)


class Predictor:

  def __init__(self, inference_graph, subgraph_name=None, checkpoint=None,
               device_type='gpu', tf_master='', session_config=None, clear_device_placement=False,
               device=None, use_cuda_graph=False, model_cfg=None):
    del tf_master, session_config, clear_device_placement
    if isinstance(inference_graph, str):
      inference_graph = LoadInferenceGraph(inference_graph)
    self._graph = inference_graph
    self._default_subgraph = subgraph_name or 'default'
    if device is None:
      device = 'cuda' if (device_type == 'gpu' and torch.cuda.is_available()) else 'cpu'
    self._device = torch.device(device)
    self._lock = threading.Lock()
    self._use_cuda_graph = use_cuda_graph and self._device.type == 'cuda'
    self._captured = {}
    d = inference_graph.asset_dir
    if model_cfg is None:
      with open(os.path.join(d, 'params.txt')) as f:
        text = f.read()
      model_cfg = self._ParamsFromText(text)
    model_cfg.is_inference = True
    from lingvo_b200.core import cluster_factory  # pylint: disable=g-import-not-at-top
    with cluster_factory.SetEval(True):
      self._model = model_cfg.Instantiate()
    weights = torch.load(os.path.join(d, 'weights.pt'), map_location='cpu') \
        if os.path.exists(os.path.join(d, 'weights.pt')) else {}
    with torch.no_grad():
      for v in self._model.vars.Flatten():
        if v.var_name in weights:
          v.data.copy_(weights[v.var_name].to(v.dtype))
    if checkpoint:
      self.Load(checkpoint)
    self._model.to(self._device)
    self._task = self._model.GetTask(inference_graph.task_name) if inference_graph.task_name \
        else self._model.tasks[0]
    from lingvo_b200.core import cluster_factory as cf  # pylint: disable=g-import-not-at-top
    with cf.SetEval(True):
      self._subgraphs = self._task.Inference()

  @staticmethod
  def _ParamsFromText(text):
    """Rebuilds the model Params from `params.txt` (the `cls` line names the class)."""
    cls_line = [l for l in text.splitlines() if l.startswith('cls :')]
    import importlib  # pylint: disable=g-import-not-at-top
    spec = cls_line[0].split(':', 1)[1].strip()
    # "type/module/Class"
    parts = spec.split('/')
    module, name = parts[-2], parts[-1]
    p = getattr(importlib.import_module(module), name).Params()
    p.FromText(text)
    return p

  @property
  def subgraphs(self) -> List[str]:
    return sorted(self._subgraphs)

  # -- signature introspection (ref :157-222) ----------------------------------------------------
  def _get_subgraph(self, subgraph_name=None):   # pylint: disable=invalid-name
    name = subgraph_name or self._default_subgraph
    if name not in self._subgraphs:
      raise KeyError('Subgraph %s not defined. Valid subgraphs: %s' % (name, self.subgraphs))
    return name, self._subgraphs[name], (self._graph.subgraphs or {}).get(name, {})

  def _get_subgraph_feeds(self, subgraph_name=None):   # pylint: disable=invalid-name
    """{feed key: shape or None}: keys from the bundle's signature, else from the callable's
    arguments; shapes from the bundle's `feeds_meta` when recorded."""
    import inspect  # pylint: disable=g-import-not-at-top
    _, fn, spec = self._get_subgraph(subgraph_name)
    keys = list(spec.get('feeds') or [])
    if not keys:
      try:
        keys = [q.name for q in inspect.signature(fn).parameters.values()
                if q.kind in (q.POSITIONAL_OR_KEYWORD, q.KEYWORD_ONLY)]
      except (TypeError, ValueError):
        keys = []
    meta = spec.get('feeds_meta') or {}
    return {k: (meta.get(k) or {}).get('shape') for k in keys}

  def _get_subgraph_fetches(self, subgraph_name=None):   # pylint: disable=invalid-name
    """{fetch key: shape or None}: from the bundle's signature, completed by what the
    subgraph returned on earlier `Run`s."""
    name, _, spec = self._get_subgraph(subgraph_name)
    meta = spec.get('fetches_meta') or {}
    out = {k: (meta.get(k) or {}).get('shape') for k in (spec.get('fetches') or [])}
    for k, shape in self.__dict__.setdefault('_seen_fetches', {}).get(name, {}).items():
      out.setdefault(k, shape)
      if out[k] is None:
        out[k] = shape
    return out

  @property
  def fetch_keys(self):
    return sorted(self._get_subgraph_fetches())

  @property
  def feed_keys(self):
    return sorted(self._get_subgraph_feeds())

  @property
  def fetch_shapes(self):
    return NestedMap(self._get_subgraph_fetches())

  @property
  def feed_shapes(self):
    return NestedMap(self._get_subgraph_feeds())

  def subgraph_fetch_keys(self, subgraph_name):   # pylint: disable=invalid-name
    return sorted(self._get_subgraph_fetches(subgraph_name))

  def subgraph_feed_keys(self, subgraph_name):   # pylint: disable=invalid-name
    return sorted(self._get_subgraph_feeds(subgraph_name))

  def subgraph_fetch_shapes(self, subgraph_name):   # pylint: disable=invalid-name
    return NestedMap(self._get_subgraph_fetches(subgraph_name))

  def subgraph_feed_shapes(self, subgraph_name):   # pylint: disable=invalid-name
    return NestedMap(self._get_subgraph_feeds(subgraph_name))

  def _LoadCheckpoint(self, checkpoint):
    self.Load(checkpoint)

  def Load(self, checkpoint):
    from lingvo_b200.core import checkpointer  # pylint: disable=g-import-not-at-top
    checkpointer.Checkpointer(os.path.dirname(checkpoint), self._model).RestoreFromPath(
        checkpoint_path=checkpoint)

  def _ToDevice(self, x):
    if isinstance(x, torch.Tensor):
      return x.to(self._device, non_blocking=True)
    if isinstance(x, (list, tuple)) and x and not isinstance(x[0], str):
      return torch.as_tensor(x, device=self._device)
    if hasattr(x, '__array__'):
      return torch.as_tensor(x, device=self._device)
    return x

  def Run(self, fetch_keys=None, validate_fetches=True, subgraph_name=None, **kwargs):
    """Runs one subgraph. Returns a list (if `fetch_keys` is a list) or a NestedMap."""
    name = subgraph_name or self._default_subgraph
    fn = self._subgraphs[name]
    feeds = {k: self._ToDevice(v) for k, v in kwargs.items()}
    from lingvo_b200.core import cluster_factory  # pylint: disable=g-import-not-at-top
    with self._lock, torch.no_grad(), cluster_factory.SetEval(True):
      out = self._RunGraphed(name, fn, feeds) if self._use_cuda_graph else fn(**feeds)
    if not isinstance(out, NestedMap):
      out = NestedMap(out) if isinstance(out, dict) else NestedMap(output=out)
    seen = self.__dict__.setdefault('_seen_fetches', {}).setdefault(name, {})
    for k, v in out.items():
      seen[k] = list(v.shape) if isinstance(v, torch.Tensor) else None
    if fetch_keys is None:
      return out
    if isinstance(fetch_keys, str):
      return out.get(fetch_keys)
    missing = [k for k in fetch_keys if k not in out]
    if missing and validate_fetches:
      raise KeyError('%s is not in the list of available fetches: %s' % (missing, sorted(out)))
    return [out.get(k) for k in fetch_keys]

  def _RunGraphed(self, name, fn, feeds):
    sig = (name,) + tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(feeds.items())
                          if isinstance(v, torch.Tensor))
    ent = self._captured.get(sig)
    if ent is None:
      static = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in feeds.items()}
      s = torch.cuda.Stream()
      s.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(s):
        for _ in range(2):                      # warm-up outside capture
          fn(**static)
      torch.cuda.current_stream().wait_stream(s)
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        out = fn(**static)
      ent = (g, static, out)
      self._captured[sig] = ent
    g, static, out = ent
    for k, v in feeds.items():
      if isinstance(v, torch.Tensor):
        static[k].copy_(v)
    g.replay()
    return out.Transform(lambda x: x.clone()) if isinstance(out, NestedMap) else out
