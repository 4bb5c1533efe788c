"""Composition layers used by builders.

Reference `lingvo/core/builder_layers.py` (1515 LoC): `FirstNLayer`,
`ArgIndexLayer`, `CreateNestedMapLayer`, `RepeatLayer` (scan over stacked
per-layer vars :117-266), `SoftCondLayer`, `ParallelRepeatLayer`,
`SequentialLayer`, `UnarySequentialLayer`, `GraphLayer` with string
signatures `"a,b->c"` (:597-1003), `ParallelLayer`, `MapLayer`,
`LinearLayer`/`BiasLayer`, `BranchLayer`, `BatchParallelLayer`, `FnLayer`,
`RematerializationLayer`, `PrintShapeLayer`, `ReshapeLayer`, `ConcatLayer`,
`SliceLayer`.

The graph-signature parser here is a small recursive-descent parser written
from the grammar described in the reference docstring (:651-686):
  signature := inputs '->' outputs ; item := path | '[' items ']' | '(' k=item,… ')'
"""

from __future__ import annotations

import re
from typing import Any, Callable, Dict, List

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core import quant_utils
from lingvo_b200.core.nested_map import NestedMap

WeightParams = py_utils.WeightParams
WeightInit = py_utils.WeightInit


def _ToTuple(x):
  return x if isinstance(x, tuple) else (x,)


def _MaybeStackExtraTheta(theta, all_vars, repeat):
  return theta


class FirstNLayer(base_layer.BaseLayer):
  """Returns the first n args."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('n', None, 'The number of args to return.')
    return p

  def FProp(self, theta, *args):
    p = self.params
    assert len(args) >= p.n
    return args[0] if p.n == 1 else tuple(args[:p.n])

  @classmethod
  def FPropMeta(cls, p, *args):
    return NestedMap(flops=0, out_shapes=args[:p.n])


class ArgIndexLayer(base_layer.BaseLayer):
  """Select args with a list of indices."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('idx', [], 'Indices of the args to return.')
    return p

  def FProp(self, theta, *args):
    r = tuple(args[i] for i in self.params.idx)
    return r[0] if len(r) == 1 else r

  @classmethod
  def FPropMeta(cls, p, *args):
    return NestedMap(flops=0, out_shapes=tuple(args[i] for i in p.idx))


class CreateNestedMapLayer(base_layer.BaseLayer):
  """Packs args into a NestedMap with the given keys."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('keys', [], 'Keys of the NestedMap.')
    return p

  def FProp(self, theta, *args):
    assert len(args) == len(self.params.keys)
    out = NestedMap()
    for k, v in zip(self.params.keys, args):
      out.Set(k, v)
    return out


class RepeatLayer(base_layer.BaseLayer):
  """Applies `body` `repeat` times; variables are *stacked* `[repeat, …]`.

  Reference :117-266 runs a `Recurrent` scan over the stacked variables. Here
  the body layer is built once under a `VariableShapePrefixContext(repeat)`
  (so every var gets a leading `[repeat]` dim with per-slice fan-in/out), and
  FProp loops over `theta[i]` — optionally rematerialising each iteration.
  Per-step dropout seeds advance through the step-seed machinery.
  """

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('body', None, 'The param for the main network layer.')
    p.Define('repeat', 1, 'Repeat layers specified in `body` this many times.')
    p.Define('per_layer_vars', False, 'Use separate variables for each layer.')
    p.Define('unroll', 'never', 'Kept for parity: always|never|eval_only.')
    p.Define('remat', False, 'Rematerialise each iteration in backward.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.repeat > 0
    if p.per_layer_vars:
      self.CreateChildren('body_iter', [
          p.body.Copy().Set(name='%s_%d' % (p.body.name or 'body', i))
          for i in range(p.repeat)])
    else:
      self.CreateChild('body', p.body)

  def _CreateChildrenVariables(self):
    p = self.params
    if p.per_layer_vars:
      super()._CreateChildrenVariables()
      return
    with py_utils.VariableShapePrefixContext(p.repeat):
      self._InstantiateNode('body', self.children['body'])

  def FProp(self, theta, *args):
    p = self.params
    xs = args
    for i in range(p.repeat):
      if p.per_layer_vars:
        layer, th = self.body_iter[i], theta.body_iter[i]
      else:
        layer = self.body
        th = theta.body.Transform(lambda t: t[i])
      if p.remat and not self.do_eval:
        flat_keys = None

        def run(*a, layer=layer, th=th):
          return _ToTuple(layer.FProp(th, *a))
        tens = [a for a in xs]
        xs = _ToTuple(py_utils.RematerializeFn(run, *tens))
      else:
        xs = _ToTuple(layer.FProp(th, *xs))
    return xs[0] if len(xs) == 1 else tuple(xs)

  @classmethod
  def FPropMeta(cls, p, *args):
    meta = p.body.cls.FPropMeta(p.body, *args)
    return NestedMap(flops=p.repeat * meta.flops, out_shapes=meta.out_shapes)


GenericRepeatLayer = RepeatLayer


class ParallelRepeatLayer(RepeatLayer):
  """Runs `repeat` copies of body on leading-dim slices of the inputs."""

  def FProp(self, theta, *args):
    p = self.params
    outs = []
    for i in range(p.repeat):
      th = theta.body.Transform(lambda t: t[i])
      outs.append(_ToTuple(self.body.FProp(th, *[a[i] for a in args])))
    res = tuple(torch.stack([o[j] for o in outs]) for j in range(len(outs[0])))
    return res[0] if len(res) == 1 else res


class SoftCondLayer(base_layer.BaseLayer):
  """Soft mixture over `num_experts` copies of `body` weights (:274-361)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('body', None, 'The param for the main network layer.')
    p.Define('num_experts', None, 'Number of experts.')
    p.Define('cond_dim', None, 'Dimension of the conditioning input.')
    p.Define('nonzeros_mean', False, 'Average over non-padded positions.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.num_experts and p.cond_dim
    self.CreateChild('body', p.body)

  def _CreateChildrenVariables(self):
    with py_utils.VariableShapePrefixContext(self.params.num_experts):
      self._InstantiateNode('body', self.children['body'])

  def _CreateLayerVariables(self):
    p = self.params
    self.CreateVariable('w', WeightParams([p.cond_dim, p.num_experts],
                                          p.params_init, p.dtype))

  def FProp(self, theta, inputs, *args):
    p = self.params
    x = inputs.float()
    if p.nonzeros_mean:
      nz = (x.abs().sum(-1, keepdim=True) > 0).float()
      pooled = (x * nz).sum(1) / torch.clamp(nz.sum(1), min=1.0)
    else:
      pooled = x.mean(1)
    dist = torch.sigmoid(torch.matmul(pooled, theta.w.float())).mean(0)  # [E]
    mixed = theta.body.Transform(
        lambda t: torch.tensordot(dist.to(t.dtype), t, dims=1))
    return self.body.FProp(mixed, inputs, *args)


class SequentialLayer(base_layer.BaseLayer):
  """A layer which connects a few layers in a sequence."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'A list of layers\' params.')
    p.Define('repeat', 1, 'Repeat layers specified in \'sub\' this many times.')
    p.Define('nary', True, 'Layers take/return tuples.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.name
    if p.repeat <= 1:
      self._seq = []
      for sub in p.sub:
        assert sub.name
        self.CreateChild(sub.name, sub)
        self._seq.append((sub.name, self.children[sub.name]))
    else:
      self.CreateChild('rep', RepeatLayer.Params().Set(
          name=p.name, repeat=p.repeat,
          body=SequentialLayer.Params().Set(name='body', sub=p.sub)))

  def FProp(self, theta, *args):
    p = self.params
    if p.repeat > 1:
      return self.rep.FProp(theta.rep, *args)
    out = args
    for name, ch in self._seq:
      out = _ToTuple(ch.FProp(theta[name], *out))
    return out[0] if len(out) == 1 else tuple(out)

  @classmethod
  def FPropMeta(cls, p, *args):
    total = 0
    for _ in range(p.repeat):
      for sub in p.sub:
        meta = sub.cls.FPropMeta(sub, *args)
        total += meta.flops
        args = meta.out_shapes
    return NestedMap(flops=total, out_shapes=args)


class UnarySequentialLayer(base_layer.BaseLayer):
  """Sequence of layers each taking/returning a single tensor."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'A list of layers\' params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._seq = []
    for sub in self.params.sub:
      self.CreateChild(sub.name, sub)
      self._seq.append(sub.name)

  def FProp(self, theta, x):
    for name in self._seq:
      x = self.children[name].FProp(theta[name], x)
    return x


class GraphSignature:
  """Parses `"a,b.c,[d,e],(k=f)->g,h.i"` signatures."""

  def __init__(self, signature: str):
    self._signature = signature
    sig = signature.replace(' ', '')
    ins, outs = sig.split('->')
    self._outputs = [o for o in outs.split(',') if o] if outs else []
    for o in self._outputs:
      if not re.fullmatch(r'[A-Za-z_][\w.\[\]]*', o):
        raise ValueError('Invalid output %r in %r' % (o, signature))
    self._text = ins
    self._pos = 0
    self._inputs = self._ParseItems(end=None) if ins else []
    if self._pos != len(self._text):
      raise ValueError('Could not parse %r' % signature)

  def __str__(self):
    return self._signature

  @property
  def inputs(self):
    return self._inputs

  @property
  def outputs(self):
    return self._outputs

  def _Peek(self):
    return self._text[self._pos] if self._pos < len(self._text) else ''

  def _ParseItems(self, end):
    items = []
    while self._pos < len(self._text) and self._Peek() != end:
      items.append(self._ParseItem())
      if self._Peek() == ',':
        self._pos += 1
    return items

  def _ParseItem(self):
    c = self._Peek()
    if c == '[':
      self._pos += 1
      items = self._ParseItems(']')
      assert self._Peek() == ']', self._signature
      self._pos += 1
      return items
    if c == '(':
      self._pos += 1
      out = {}
      while self._Peek() != ')':
        m = re.match(r'([A-Za-z_]\w*)=', self._text[self._pos:])
        if not m:
          raise ValueError('Bad map in %r' % self._signature)
        self._pos += m.end()
        out[m.group(1)] = self._ParseItem()
        if self._Peek() == ',':
          self._pos += 1
      self._pos += 1
      return out
    m = re.match(r'[A-Za-z_][\w.]*(?:\[\d+\])*', self._text[self._pos:])
    if not m:
      raise ValueError('Bad path at %d in %r' % (self._pos, self._signature))
    self._pos += m.end()
    return m.group(0)


class GraphTensors:
  """Name → tensor store supporting dotted NestedMap paths."""

  def __init__(self):
    self._named = NestedMap()

  def StoreTensor(self, path: str, tensor):
    if isinstance(tensor, NestedMap) and self._named.Has(path) and isinstance(
        self._named.GetItem(path), NestedMap):
      cur = self._named.GetItem(path)
      for k, v in tensor.FlattenItems():
        cur.Set(k, v)
      return
    self._named.Set(path, tensor)

  def GetTensor(self, path):
    if isinstance(path, str):
      return self._named.GetItem(path)
    if isinstance(path, list):
      return [self.GetTensor(p) for p in path]
    if isinstance(path, dict):
      return NestedMap({k: self.GetTensor(v) for k, v in path.items()})
    raise TypeError(path)


class GraphLayer(base_layer.BaseLayer):
  """Dataflow graph of sub-layers wired by string signatures (:886-1003)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_endpoints', [], 'Names of the graph inputs.')
    p.Define('output_endpoints', [], 'Names of the graph outputs.')
    p.Define('sub', [], 'A list of (signature, layer params).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.name
    assert p.input_endpoints
    self._seq = []
    for i, (signature, sub) in enumerate(p.sub):
      assert sub.name
      sig = GraphSignature(signature)
      assert len(sig.inputs) <= 1 or True
      name = sub.name
      if name in self.children:
        raise ValueError('Duplicate sub-layer name %s in GraphLayer %s' %
                         (name, p.name))
      self.CreateChild(name, sub)
      self._seq.append((name, sig, self.children[name]))

  def FProp(self, theta, *args):
    p = self.params
    graph = GraphTensors()
    assert len(p.input_endpoints) == len(args), (p.input_endpoints, len(args))
    for n, t in zip(p.input_endpoints, args):
      if isinstance(t, torch.Tensor) or not isinstance(t, NestedMap):
        graph.StoreTensor(n, t)
      else:
        graph.StoreTensor(n, t.DeepCopy())
    for name, sig, ch in self._seq:
      ins = [graph.GetTensor(x) for x in sig.inputs]
      outs = _ToTuple(ch.FProp(theta[name], *ins))
      assert len(outs) == len(sig.outputs), (name, len(outs), sig.outputs)
      for n, t in zip(sig.outputs, outs):
        graph.StoreTensor(n, t)
    res = tuple(graph.GetTensor(x) for x in p.output_endpoints)
    return res[0] if len(res) == 1 else res

  @classmethod
  def FPropMeta(cls, p, *args):
    total = 0
    shapes = dict(zip(p.input_endpoints, args))
    for signature, sub in p.sub:
      sig = GraphSignature(signature)
      ins = [shapes[x] for x in sig.inputs if isinstance(x, str)]
      meta = sub.cls.FPropMeta(sub, *ins)
      total += meta.flops
      for n, s in zip(sig.outputs, meta.out_shapes):
        shapes[n] = s
    return NestedMap(flops=total,
                     out_shapes=tuple(shapes[x] for x in p.output_endpoints))


class ParallelLayer(base_layer.BaseLayer):
  """Runs sub-layers on the same inputs and merges their outputs."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'A list of layers\' params.')
    p.Define('merge', None, 'fn(list of output tuples) → outputs.')
    p.Define('merge_meta', None, 'FPropMeta of merge.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._seq = []
    for sub in self.params.sub:
      self.CreateChild(sub.name, sub)
      self._seq.append(sub.name)

  def FProp(self, theta, *args):
    p = self.params
    outs = [_ToTuple(self.children[n].FProp(theta[n], *args))
            for n in self._seq]
    rets = p.merge(outs) if p.merge else tuple(
        sum(o[i] for o in outs) for i in range(len(outs[0])))
    rets = _ToTuple(rets)
    return rets[0] if len(rets) == 1 else rets


class MapLayer(base_layer.BaseLayer):
  """Applies fn to each arg."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fn', None, 'A callable tensor→tensor.')
    p.Define('fn_meta', None, 'A callable shape→(flops, shape).')
    p.Define('kwargs', {}, 'Keyword args to fn.')
    return p

  def FProp(self, theta, *args):
    p = self.params
    r = tuple(p.fn(a, **p.kwargs) for a in args)
    return r[0] if len(r) == 1 else r

  @classmethod
  def FPropMeta(cls, p, *args):
    flops, out = 0, []
    for a in args:
      if p.fn_meta:
        f, s = p.fn_meta(a)
        flops += f
        out.append(s)
      else:
        out.append(a)
    return NestedMap(flops=flops, out_shapes=tuple(out))


class LinearLayer(quant_utils.QuantizableLayer):
  """y = x·w, `w [input_dims, output_dims]` (sharding-aware)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_dims', 0, 'Depth of the input.')
    p.Define('output_dims', 0, 'Depth of the output.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.TrackQTensor('act')

  def _CreateLayerVariables(self):
    p = self.params
    self.CreateVariable('w', WeightParams(
        [p.input_dims, p.output_dims], p.params_init, p.dtype,
        device_mesh=p.device_mesh,
        tensor_split_dims_mapping=p.weight_split_dims_mapping))

  def FProp(self, theta, inputs):
    from lingvo_b200.ops import gemm
    w = self.QWeight(theta.w)
    out = gemm.linear(self._CastToFPropDtype(inputs), w.to(self.fprop_dtype)
                      if w.dtype != inputs.dtype else w)
    return self.QAct('act', out)

  @classmethod
  def FPropMeta(cls, p, inputs):
    from lingvo_b200.core import tshape
    return NestedMap(
        flops=inputs.num_elements() * p.output_dims * 2,
        out_shapes=(tshape.Shape(inputs[:-1] + [p.output_dims]),))


class BiasLayer(base_layer.BaseLayer):
  """y = x + b."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dims', 0, 'Depth of the input.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    self.CreateVariable('b', WeightParams(
        [p.dims], WeightInit.Constant(0.0), p.dtype, device_mesh=p.device_mesh,
        tensor_split_dims_mapping=p.weight_split_dims_mapping))

  def FProp(self, theta, inputs):
    return inputs + theta.b.to(inputs.dtype)

  @classmethod
  def FPropMeta(cls, p, inputs):
    return NestedMap(flops=inputs.num_elements(), out_shapes=(inputs,))


class BranchLayer(base_layer.BaseLayer):
  """Runs `body` and returns both inputs-passthrough and fetched outputs."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('body', None, 'The param for the main network layer.')
    p.Define('fetches', [], 'Names of fetch points in body.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('body', self.params.body)

  def FProp(self, theta, *args):
    p = self.params
    outs = _ToTuple(self.body.FProp(theta.body, *args))
    fetched = []
    for f in p.fetches:
      fetched.append(self.body.GetDescendant(f).activation)
    return outs + tuple(fetched)


class BatchParallelLayer(base_layer.BaseLayer):
  """Splits the batch over this process's devices / streams (:1283)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'A layer param.')
    p.Define('num_splits', None, 'Defaults to cluster num_devices_per_split.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('sub', self.params.sub)

  def FProp(self, theta, *args):
    p = self.params
    n = p.num_splits or self.cluster.num_devices_per_split
    if n <= 1:
      return self.sub.FProp(theta.sub, *args)
    chunks = [torch.chunk(a, n, dim=0) for a in args]
    outs = [_ToTuple(self.sub.FProp(theta.sub, *[c[i] for c in chunks]))
            for i in range(n)]
    res = tuple(torch.cat([o[j] for o in outs], dim=0)
                for j in range(len(outs[0])))
    return res[0] if len(res) == 1 else res


class FnLayer(base_layer.BaseLayer):
  """A layer applying `fn(*args)`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fn', None, 'A callable.')
    p.Define('fn_meta', None, 'Callable(shapes) → NestedMap(flops, out_shapes).')
    return p

  def FProp(self, theta, *args):
    r = self.params.fn(*args)
    return r

  @classmethod
  def FPropMeta(cls, p, *args):
    if p.fn_meta:
      return p.fn_meta(*args)
    return NestedMap(flops=0, out_shapes=args)


class RematerializationLayer(base_layer.BaseLayer):
  """Recomputes `body` activations in the backward pass."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('body', None, 'The main layer whose FProp is rematerialised.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('body', self.params.body)

  def FProp(self, theta, *xs):
    if self.do_eval:
      return self.body.FProp(theta.body, *xs)

    def fn(*a):
      return self.body.FProp(theta.body, *a)
    return py_utils.RematerializeFn(fn, *xs)

  @classmethod
  def FPropMeta(cls, p, *args):
    meta = p.body.cls.FPropMeta(p.body, *args)
    return NestedMap(flops=meta.flops * 2, out_shapes=meta.out_shapes)


class PrintShapeLayer(base_layer.BaseLayer):
  """Logs the shapes of its inputs (debugging)."""

  def FProp(self, theta, *args):
    import logging
    for i, a in enumerate(args):
      logging.info('PrintShapeLayer %s[%d]: %s', self.params.name, i,
                   tuple(getattr(a, 'shape', ())))
    return args[0] if len(args) == 1 else args

  @classmethod
  def FPropMeta(cls, p, *args):
    return NestedMap(flops=0, out_shapes=args)


class ReshapeLayer(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('shape', None, 'Target shape (may contain -1).')
    return p

  def FProp(self, theta, inp):
    return inp.reshape(list(self.params.shape))


class ConcatLayer(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('axis', -1, 'Concat axis.')
    return p

  def FProp(self, theta, *inps):
    return torch.cat(list(inps), dim=self.params.axis)


class SliceLayer(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('slices', None, 'Tuple of slice objects / ints.')
    return p

  def FProp(self, theta, inp):
    return inp[self.params.slices]


class SliceHelper:

  def __getitem__(self, args):
    return args
