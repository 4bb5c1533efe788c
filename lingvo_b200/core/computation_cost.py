"""Per-layer cost accounting (ref `lingvo/core/computation_cost.py`)."""
from lingvo_b200.core import bn_layers

COST_METRICS = {'flops': 'int64'}


def Prepare(layer):
  """Registers a `flops` accumulator on `layer` and all descendants."""

  def _Walk(l):
    if isinstance(l, (list, tuple)):
      for x in l:
        _Walk(x)
      return
    for name in COST_METRICS:
      if name not in l.accumulators:
        l.RegisterAccumulator(name, bn_layers.AddingAccumulator([], 'float32'))
    for _, child in sorted(l.children.items()):
      _Walk(child)
  _Walk(layer)


def Add(layer, cost_metric_name, cost):
  if cost_metric_name in layer.accumulators:
    layer.accumulators[cost_metric_name].Update(cost)


def Get(layer, cost_metric_name):
  if cost_metric_name not in layer.accumulators:
    raise ValueError('Prepare was not called for %s' % layer.path)
  return layer.accumulators[cost_metric_name].GetValue()
