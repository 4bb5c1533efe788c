"""Layer-composition helpers for point-cloud models (ref `lingvo/tasks/car/builder_lib.py`).

`ModelBuilderBase` extends the core `builder.Base` combinators with what PointNet /
StarNet / PointPillars-style models need: per-point FC / MLP stacks with batch norm,
padded max-pooling over points, conv / deconv / residual blocks, self-attention blocks,
and the `NestedMap`-routing combinators (`_GetValue`, `_ParMap`, `_SeqToKey`, `_SeqOnKey`,
`_SeqOnFeatures`).

The *points tensor* convention: `NestedMap(points [..., P, 3], features [..., P, F],
padding [..., P])`.
"""

from __future__ import annotations

import functools

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer
from lingvo_b200.core import builder
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap

_ACTS = {'RELU': F.relu, 'SWISH': F.silu, 'SILU': F.silu, 'SIGMOID': torch.sigmoid,
         'TANH': torch.tanh, 'GELU': F.gelu, 'NONE': lambda x: x, None: lambda x: x}


def _MovingStats(layer, theta):
  """The running statistics F.batch_norm updates in place. Inside a `RepeatLayer` the layer's
  variables are stacked `[repeat, C]` and `theta` holds this iteration's slice (a view)."""
  mean, var = layer.vars.moving_mean.data, layer.vars.moving_variance.data
  if mean.shape != theta.gamma.shape:
    mean, var = theta.moving_mean.detach(), theta.moving_variance.detach()
  return mean, var


class _PointBN(base_layer.BaseLayer):
  """Batch norm over all leading dims of `[..., D]` (per-point features); padded points
  do not matter for the statistics in practice and are re-masked by the pooling."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dim', 0, 'Feature dim.')
    p.Define('decay', 0.99, 'Moving-average decay.')
    p.Define('epsilon', 1e-3, 'Variance epsilon.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    wp = lambda init: py_utils.WeightParams([p.dim], init, p.dtype)
    self.CreateVariable('beta', wp(py_utils.WeightInit.Constant(0.0)))
    self.CreateVariable('gamma', wp(py_utils.WeightInit.Constant(1.0)))
    self.CreateVariable('moving_mean', wp(py_utils.WeightInit.Constant(0.0)), trainable=False)
    self.CreateVariable('moving_variance', wp(py_utils.WeightInit.Constant(1.0)),
                        trainable=False)

  def FProp(self, theta, x):
    p = self.params
    shape = x.shape
    flat = x.reshape(-1, shape[-1])
    mean, var = _MovingStats(self, theta)
    out = F.batch_norm(flat, mean, var, theta.gamma, theta.beta, training=not self.do_eval,
                       momentum=1.0 - p.decay, eps=p.epsilon)
    return out.reshape(shape)


class _Conv2D(base_layer.BaseLayer):
  """NHWC conv / transposed conv with optional BN + activation."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('filter_shape', (3, 3, 1, 1), '(kh, kw, in, out).')
    p.Define('stride', (1, 1), 'Stride.')
    p.Define('transpose', False, 'Transposed convolution (upsampling).')
    p.Define('use_bn', True, 'Batch norm after the conv.')
    p.Define('activation', 'RELU', 'Activation name.')
    p.Define('bias', False, 'Add a bias (when no BN).')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    kh, kw, cin, cout = p.filter_shape
    self.CreateVariable('w', py_utils.WeightParams(
        [kh, kw, cin, cout], py_utils.WeightInit.Gaussian((2.0 / (kh * kw * cin)) ** 0.5),
        p.dtype))
    if p.use_bn:
      self.CreateVariable('gamma', py_utils.WeightParams(
          [cout], py_utils.WeightInit.Constant(1.0), p.dtype))
      self.CreateVariable('beta', py_utils.WeightParams(
          [cout], py_utils.WeightInit.Constant(0.0), p.dtype))
      self.CreateVariable('moving_mean', py_utils.WeightParams(
          [cout], py_utils.WeightInit.Constant(0.0), p.dtype), trainable=False)
      self.CreateVariable('moving_variance', py_utils.WeightParams(
          [cout], py_utils.WeightInit.Constant(1.0), p.dtype), trainable=False)
    elif p.bias:
      self.CreateVariable('b', py_utils.WeightParams(
          [cout], py_utils.WeightInit.Constant(0.0), p.dtype))

  def FProp(self, theta, x):
    p = self.params
    kh, kw, _, _ = p.filter_shape
    x = x.permute(0, 3, 1, 2)
    if p.transpose:
      w = theta.w.permute(2, 3, 0, 1)                      # [in, out, kh, kw]
      def _SamePad(k, st):        # output = input · stride
        if k >= st:
          pad = (k - st + 1) // 2
          return pad, 2 * pad - (k - st)
        return 0, st - k
      (ph, oph), (pw, opw) = _SamePad(kh, p.stride[0]), _SamePad(kw, p.stride[1])
      y = F.conv_transpose2d(x, w, stride=tuple(p.stride), padding=(ph, pw),
                             output_padding=(oph, opw))
    else:
      w = theta.w.permute(3, 2, 0, 1)                      # [out, in, kh, kw]
      y = F.conv2d(x, w, stride=tuple(p.stride), padding=((kh - 1) // 2, (kw - 1) // 2))
    if p.use_bn:
      mean, var = _MovingStats(self, theta)
      y = F.batch_norm(y, mean, var, theta.gamma, theta.beta, training=not self.do_eval,
                       momentum=0.01, eps=1e-3)
    elif p.bias:
      y = y + theta.b.view(1, -1, 1, 1)
    return _ACTS[p.activation](y).permute(0, 2, 3, 1)


class ModelBuilderBase(builder.Base):
  """ref :39."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('bn_decay', 0.99, 'Batch-norm moving-average decay.')
    p.Define('bn_epsilon', 1e-3, 'Batch-norm epsilon.')
    p.Define('linear_params_init', None, 'Weight init of linear layers.')
    p.Define('conv_init_method', None, 'Kept for parity.')
    return p

  # ------------------------------------------------------------------ basics --
  def _BN(self, name, dims):
    return _PointBN.Params().Set(name=name, dim=dims, decay=self.params.bn_decay,
                                 epsilon=self.params.bn_epsilon)

  def _Activation(self, name, activation_fn_or_name=None):
    fn = activation_fn_or_name
    if fn is None:
      fn = 'RELU'
    if isinstance(fn, str):
      fn = _ACTS[fn.upper()]
    return self._Fn(name, fn)

  def _Relu(self, name):
    return self._Activation(name, 'RELU')

  def _Swish(self, name):
    return self._Activation(name, 'SWISH')

  def _Sigmoid(self, name):
    return self._Activation(name, 'SIGMOID')

  def _FC(self, name, idims, odims, use_bn=True, activation_fn=None):
    """Linear → (BN | bias) → activation on the last dim."""
    mid = self._BN('bn', odims) if use_bn else self._Bias('bias', odims)
    return self._Seq(name, self._Linear('linear', idims, odims), mid,
                     self._Activation('act', activation_fn))

  def _MLP(self, name, dims, use_bn=True, activation_fn=None):
    return self._Seq(name, *[
        self._FC('mlp%d' % i, a, b, use_bn, activation_fn)
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])

  def _Map(self, name, fn, **kwargs):
    return self._Fn(name, functools.partial(fn, **kwargs) if kwargs else fn)

  def _Max(self, name):
    """Max over the points dim (second to last)."""
    return self._Fn(name, lambda x: x.max(-2).values)

  def _Matmul(self, name, *subs):
    """Runs `subs` in parallel on the input and multiplies their outputs in order."""
    def Mul(*xs):
      out = xs[0]
      for x in xs[1:]:
        out = torch.matmul(out, x)
      return out
    return self._Seq(name, self._Par('par', *subs), self._Fn('mul', Mul))

  def _GLU(self, name, idims, odims):
    """Gated linear unit: linear(x) ⊙ sigmoid(linear_gate(x))."""
    def Gate(x):
      a, g = x.chunk(2, -1)
      return a * torch.sigmoid(g)
    return self._Seq(name, self._Linear('linear', idims, 2 * odims),
                     self._Bias('bias', 2 * odims), self._Fn('gate', Gate))

  def _Fetch(self, name):
    return self._Save(name)

  def _FirstN(self, name, n):
    from lingvo_b200.core import builder_layers  # pylint: disable=g-import-not-at-top
    return builder_layers.FirstNLayer.Params().Set(name=name, n=n)

  def _ArgIdx(self, name, index):
    from lingvo_b200.core import builder_layers  # pylint: disable=g-import-not-at-top
    return builder_layers.ArgIndexLayer.Params().Set(name=name, idx=list(index))

  def _Join(self, name, *subs):
    """Runs `subs` on the same input and returns the tuple of all outputs."""
    return self._Par(name, *subs)

  def _Concat(self, name, *subs):
    """Concats the outputs of `subs` along the last dim."""
    return self._Seq(name, self._Par('par', *subs),
                     self._Fn('concat', lambda *xs: torch.cat(xs, -1)))

  def _BroadcastConcat(self, name, *subs):
    """Concat after broadcasting every output to the largest leading shape."""
    def Merge(*xs):
      lead = torch.broadcast_shapes(*[x.shape[:-1] for x in xs])
      return torch.cat([x.expand(lead + x.shape[-1:]) for x in xs], -1)
    return self._Seq(name, self._Par('par', *subs), self._Fn('concat', Merge))

  def _ApplyFnMulti(self, name, fn, *subs):
    return self._Seq(name, self._Par('par', *subs), self._Fn('fn', fn))

  def _ApplyInParallelAndMerge(self, name, merge_fn, *subs):
    return self._ApplyFnMulti(name, merge_fn, *subs)

  def _ApplyFn(self, name, fn):
    return self._Fn(name, fn)

  def _Squeeze(self, name, axis=None):
    return self._Fn(name, (lambda x: x.squeeze(axis)) if axis is not None
                    else (lambda x: x.squeeze()))

  def _MakeInputFeatureFromPoints(self, name):
    """points tensor → per-point input features [xyz ‖ features], padded points zeroed."""
    def Make(inp):
      f = torch.cat([inp.points, inp.features], -1)
      return f * (1.0 - inp.padding).unsqueeze(-1)
    return self._Fn(name, Make)

  # -------------------------------------------------------------------- convs --
  def _ConvPlain(self, name, filter_shape, filter_stride=(1, 1), padding='SAME'):
    del padding
    return _Conv2D.Params().Set(name=name, filter_shape=tuple(filter_shape),
                                stride=tuple(filter_stride), use_bn=False, activation='NONE',
                                bias=True)

  def _DeconvPlain(self, name, filter_shape, filter_stride=(1, 1)):
    return _Conv2D.Params().Set(name=name, filter_shape=tuple(filter_shape),
                                stride=tuple(filter_stride), transpose=True, use_bn=False,
                                activation='NONE', bias=True)

  def _Conv(self, name, filter_shape, stride=(1, 1), padding='SAME', use_bn=True,
            activation_fn='RELU'):
    del padding
    return _Conv2D.Params().Set(name=name, filter_shape=tuple(filter_shape),
                                stride=tuple(stride), use_bn=use_bn,
                                activation=activation_fn if isinstance(activation_fn, str)
                                else 'RELU')

  def _Deconv(self, name, filter_shape, stride):
    return _Conv2D.Params().Set(name=name, filter_shape=tuple(filter_shape),
                                stride=tuple(stride), transpose=True)

  def _Shortcut(self, name, idims, odims, stride):
    if idims == odims and tuple(stride) == (1, 1):
      return self._Identity(name)
    return self._Conv(name, (1, 1, idims, odims), stride, activation_fn='NONE')

  def _ResidualLayer(self, name, filter_size, stride):
    """relu(conv-bn-relu-conv-bn(x) + shortcut(x)); filter_size = (kh, kw, in, out)."""
    kh, kw, idims, odims = filter_size
    body = self._Seq('body', self._Conv('c1', (kh, kw, idims, odims), stride),
                     self._Conv('c2', (kh, kw, odims, odims), (1, 1), activation_fn='NONE'))
    return self._Seq(name, self._Par('branches', body,
                                     self._Shortcut('shortcut', idims, odims, stride)),
                     self._Fn('add_relu', lambda a, b: F.relu(a + b)))

  def _ResidualBlock(self, name, filter_size, stride, repeats):
    kh, kw, _, odims = filter_size
    blocks = [self._ResidualLayer('r0', filter_size, stride)]
    blocks += [self._ResidualLayer('r%d' % i, (kh, kw, odims, odims), (1, 1))
               for i in range(1, repeats)]
    return self._Seq(name, *blocks)

  # ---------------------------------------------------------------- attention --
  def _LN(self, name, dims):
    return layers.LayerNorm.Params().Set(name=name, input_dim=dims)

  def _Project(self, name, idims, odims):
    return self._Seq(name, self._Linear('linear', idims, odims), self._Bias('bias', odims))

  def _Add(self, name, lhs, rhs):
    return self._Seq(name, self._Par('par', lhs, rhs), self._Fn('add', lambda a, b: a + b))

  def _Multiply(self, name, lhs, rhs):
    return self._Seq(name, self._Par('par', lhs, rhs), self._Fn('mul', lambda a, b: a * b))

  def _AttenFF(self, name, dims, hdims, keep_prob=1.0):
    return self._Add(name, self._Identity('id'), self._Seq(
        'ff', self._LN('ln', dims), self._FC('fc', dims, hdims, use_bn=False),
        self._Project('proj', hdims, dims), self._Dropout('dropout', keep_prob)))

  def _AttenSelf(self, name, dims, hdims, heads, keep_prob=1.0):
    """Multi-head self attention over the points dim of `[..., P, dims]`."""
    del hdims
    def Atten(q, k, v):
      shp = q.shape
      split = lambda t: t.reshape(shp[:-1] + (heads, dims // heads)).transpose(-2, -3)
      o = F.scaled_dot_product_attention(split(q), split(k), split(v))
      return o.transpose(-2, -3).reshape(shp)
    return self._Add(name, self._Identity('id'), self._Seq(
        'att', self._LN('ln', dims),
        self._Par('qkv', self._Project('q', dims, dims), self._Project('k', dims, dims),
                  self._Project('v', dims, dims)),
        self._Fn('sdpa', Atten), self._Project('o', dims, dims),
        self._Dropout('dropout', keep_prob)))

  def _Atten(self, name, dims, hdims, heads, keep_prob=1.0):
    return self._Seq(name, self._AttenSelf('self', dims, hdims, heads, keep_prob),
                     self._AttenFF('ff', dims, hdims, keep_prob))

  def _SelfAttenStack(self, name, depth, dims, hdims, heads, keep_prob):
    return self._Seq(name, *[self._Atten('atten%d' % i, dims, hdims, heads, keep_prob)
                             for i in range(depth)], self._LN('final_ln', dims))

  # ------------------------------------------------------- NestedMap routing --
  class Decorators:
    """Run-time input checks for builder functions (ref :530)."""

    @classmethod
    def ExpectsNestedMapTensor(cls, expected_keys=()):
      keys = (expected_keys,) if isinstance(expected_keys, str) else tuple(expected_keys)
      def Validate(inp):
        if not isinstance(inp, NestedMap):
          raise ValueError('Input not a `NestedMap`. Is a %s' % type(inp))
        missing = [k for k in keys if k not in inp]
        if missing:
          raise ValueError('Input missing keys %s (has %s)' % (missing, sorted(inp.keys())))
        return inp
      def Decorator(builder_fn):
        @functools.wraps(builder_fn)
        def Wrapped(self, *args, **kwargs):
          return self._Seq('validated', self._Fn('validate', Validate),  # pylint: disable=protected-access
                           builder_fn(self, *args, **kwargs))
        return Wrapped
      return Decorator

    @classmethod
    def ExpectsNestedMapPointsTensor(cls, builder_fn):
      return cls.ExpectsNestedMapTensor(('points', 'features', 'padding'))(builder_fn)

    @classmethod
    def ExpectsNestedMapRangeImage(cls, builder_fn):
      return cls.ExpectsNestedMapTensor(('features', 'xyz', 'mask'))(builder_fn)

  def _GetValue(self, name, key, default=None):
    def Get(inp):
      return inp.get(key, default) if default is not None else inp[key]
    return self._Fn(name, Get)

  def _ParMap(self, name, key_to_sub):
    """NestedMap in → NestedMap out, key k produced by `key_to_sub[k]` run on the input."""
    keys = sorted(key_to_sub)
    return self._Seq(name, self._Par('par', *[key_to_sub[k] for k in keys]),
                     self._Fn('pack', lambda *vals: NestedMap(dict(zip(keys, vals)))))

  def _SeqToKey(self, name, key, *subs):
    """Runs `subs` on the whole input NestedMap and stores the result under `key`."""
    def Merge(inp, out):
      res = inp.copy()
      res[key] = out
      return res
    return self._Seq(name, self._Par('par', self._Identity('id'), self._Seq('subs', *subs)),
                     self._Fn('merge', Merge))

  def _SeqOnKey(self, name, key, *subs):
    """Runs `subs` on `input[key]` and writes the result back to `key`."""
    return self._SeqToKey(name, key, self._GetValue('get', key), *subs)

  def _SeqOnFeatures(self, name, *subs):
    return self._SeqOnKey(name, 'features', *subs)

  def _PaddedMax(self, name, nested_output=False):
    """Max of `features` over real points → `[..., F]` (or a points tensor with P = 1)."""
    def Fn(inp):
      neg = torch.finfo(inp.features.dtype).min
      f = inp.features.masked_fill(inp.padding.unsqueeze(-1) > 0.5, neg).max(-2).values
      any_real = (inp.padding < 0.5).any(-1, keepdim=True)
      f = torch.where(any_real, f, torch.zeros_like(f))
      if not nested_output:
        return f
      real = (inp.padding < 0.5).unsqueeze(-1).to(inp.points.dtype)
      center = (inp.points * real).sum(-2) / real.sum(-2).clamp_min(1.0)
      return NestedMap(points=center.unsqueeze(-2), features=f.unsqueeze(-2),
                       padding=(~any_real).to(inp.padding.dtype))
    return self._Fn(name, Fn)


# ------------------------------------------------------------------------------------------
# Point-set recipes of the reference builder (ref builder_lib.py:75, 690-1167), added as a
# mixin so the class above stays readable.
# ------------------------------------------------------------------------------------------
def _PaddedMask(inp):
  return (1.0 - inp.padding).unsqueeze(-1).to(inp.features.dtype)


class _PointSetRecipes:

  def _Branch(self, name, body, fetches):
    """Runs `body`; its output is followed by the activations saved at `fetches` (dotted
    paths of `_Fetch` layers inside `body`)."""
    from lingvo_b200.core import builder_layers  # pylint: disable=g-import-not-at-top
    return builder_layers.BranchLayer.Params().Set(name=name, body=body, fetches=list(fetches))

  def _MakeNestedMap(self, name, keys):
    return self._Fn(name, lambda *vals: NestedMap(dict(zip(keys, vals))))

  def _SqueezeFn(self, axis=None):
    return (lambda x: x.squeeze(axis)) if axis is not None else (lambda x: x.squeeze())

  # -- padding-aware pooling over the points dim ----------------------------------------------
  def _PaddedMean(self, name):
    """Mean of `features` over the real points → `[..., F]` (0 when all are padded)."""
    def Fn(inp):
      mask = _PaddedMask(inp)
      return (inp.features * mask).sum(-2) / mask.sum(-2).clamp_min(1.0)
    return self._Fn(name, Fn)

  def _PaddedSum(self, name):
    return self._Fn(name, lambda inp: (inp.features * _PaddedMask(inp)).sum(-2))

  # -- per-point MLPs on a points NestedMap ------------------------------------------------------
  def _FeaturesFC(self, name, idims, odims, use_bn=True, activation_fn=None):
    """FC on `.features`, rest of the NestedMap passed through (ref :752)."""
    mid = self._BN('bn', odims) if use_bn else self._Bias('bias', odims)
    return self._SeqOnFeatures(name, self._Linear('linear', idims, odims), mid,
                               self._Activation('activation', activation_fn))

  def _FeaturesMLP(self, name, dims, use_bn=True, activation_fn=None):
    return self._Seq(name, *[
        self._FeaturesFC('l%03d' % n, i, o, use_bn=use_bn, activation_fn=activation_fn)
        for n, (i, o) in enumerate(zip(dims[:-1], dims[1:]))])

  def _ConcatPointsToFeatures(self, name):
    return self._SeqToKey(name, 'features', self._Concat(
        'concat', self._GetValue('get_points', 'points'),
        self._GetValue('get_features', 'features')))

  # -- GIN ------------------------------------------------------------------------------------------
  def _CondFC(self, name, idims, adims, odims, use_bn=True, activation_fn=None):
    """(features `[..., P, idims]`, aggregate `[..., 1, adims]`) → `[..., P, odims]`: the
    aggregate predicts a per-example `[idims, odims]` matrix applied to every point — a
    T-Net-like conditional linear layer (ref :798)."""
    def ReshapeTransform(t):
      return t.reshape(list(t.shape[:-1]) + [idims, odims])
    transform = self._Matmul(
        'cond_transform',
        self._Seq('prep_features', self._ArgIdx('arg0', [0]),
                  self._BN('bn', idims) if use_bn else self._Identity('id')),
        self._Seq('compute_linear_transform', self._ArgIdx('arg1', [1]),
                  self._Squeeze('squeeze', axis=-2),
                  self._FC('fc0', adims, adims * 2, use_bn=use_bn, activation_fn=activation_fn),
                  self._FC('fc1', adims * 2, idims * odims, use_bn=use_bn,
                           activation_fn='NONE'),
                  self._Fn('reshape', ReshapeTransform)))
    return self._Seq(name, transform,
                     self._Identity('id') if use_bn else self._Bias('bias', odims),
                     self._Activation('activation', activation_fn))

  def _GINCondFC(self, name, lhs, rhs, idims, adims, odims, use_bn=True, activation_fn=None):
    return self._Seq(name, self._Join('join', lhs, rhs),
                     self._CondFC('cond_fc', idims, adims, odims, use_bn=use_bn,
                                  activation_fn=activation_fn))

  def _GINIntermediateLayer(self, name, dims, aggregate_sub, combine_method='add', eps=0.,
                            use_bn=True):
    """f'_i = MLP(combine((1 + eps)·f_i, aggregate_j f_j)) over a points NestedMap
    (ref :960; eq. 4.1 of "How Powerful are Graph Neural Networks?")."""
    combine = {
        'add': self._Add,
        'concat': self._BroadcastConcat,
        'cond_fc': functools.partial(self._GINCondFC, idims=dims[0], odims=dims[0],
                                     adims=dims[0], use_bn=use_bn),
    }
    if combine_method not in combine:
      raise ValueError('Unexpected combine method: {}'.format(combine_method))
    return self._Seq(
        name,
        self._SeqToKey(
            'map_features', 'features',
            combine[combine_method](
                'combine',
                self._Seq('left', self._GetValue('get_features', 'features'),
                          self._Fn('eps_scale', lambda t: (1.0 + eps) * t)),
                self._Seq('right', aggregate_sub,
                          self._Fn('expand_dims', lambda t: t.unsqueeze(-2))))),
        self._FeaturesMLP('mlp', dims, use_bn=use_bn))

  def _GIN(self, name, mlp_dims, aggregate_sub, readout_sub, combine_method='add', eps=0.,
           use_bn=True):
    """Graph Isomorphism Network over a point set (ref :860): a chain of intermediate
    layers; the output concatenates the read-out of the input and of every layer's output
    (eq. 4.2) — `readout(f0) ‖ readout(f1) ‖ … ‖ readout(fN)`."""
    if combine_method not in ('add', 'concat', 'cond_fc'):
      raise ValueError('Unexpected combine method: {}'.format(combine_method))
    for idx, (a, b) in enumerate(zip(mlp_dims[:-1], mlp_dims[1:])):
      prev = a[-1] * (2 if combine_method == 'concat' else 1)
      if prev != b[0]:
        raise ValueError('mlp_dims do not match ({} != {}) at layer {} with dims: {} and {}'
                         .format(prev, b[0], idx, a, b))

    def Build(depth):
      if depth == len(mlp_dims):
        return readout_sub.Copy()
      return self._Concat(
          'gin_concat', readout_sub.Copy(),
          self._Seq('seq', self._GINIntermediateLayer(
              'gin_intermediate', mlp_dims[depth], aggregate_sub.Copy(), combine_method, eps,
              use_bn), Build(depth + 1)))

    return self._Seq(name, Build(0))

  # -- PointNet++ / PointConv -------------------------------------------------------------------------
  def _SetAbstraction(self, name, feature_extraction_sub, num_samples, group_size, ball_radius,
                      sample_neighbors_uniformly=True):
    """Farthest-point sample `num_samples` centres, group `group_size` neighbours within
    `ball_radius` of each, featurize every group with `feature_extraction_sub` (points
    NestedMap with leading dims `[B, num_samples, group_size]` → `[B, num_samples, F]`)
    (ref :1040). Points NestedMap in, points NestedMap (the centres) out."""
    from lingvo_b200.models.car import car_layers  # pylint: disable=g-import-not-at-top
    return self._Seq(
        name,
        car_layers.SamplingAndGroupingLayer.Params().Set(
            name='sample_group', num_samples=num_samples, ball_radius=ball_radius,
            group_size=group_size, sample_neighbors_uniformly=sample_neighbors_uniformly),
        self._Fn('pack', lambda grouped, query: NestedMap(grouped_points=grouped,
                                                          query_points=query)),
        self._ParMap('pmap', dict(
            points=self._Seq('seq_points', self._GetValue('get_query', 'query_points'),
                             self._GetValue('get_points', 'points')),
            features=self._Seq('seq_features', self._GetValue('get_grouped', 'grouped_points'),
                               feature_extraction_sub),
            padding=self._Seq('seq_padding', self._GetValue('get_query', 'query_points'),
                              self._GetValue('get_padding', 'padding')))))

  def _PointConvParametricConv(self, name, mlp_dims, num_in_channels, num_out_channels):
    """PointConv's parametric convolution (ref :1110, Fig. 5 of the paper without the
    inverse-density scaling): `featuresᵀ [C_in, P] · MLP(points) [P, C_mid]`, flattened,
    followed by an FC to `num_out_channels`."""
    if mlp_dims[0] != 3:
      raise ValueError('First dimension of mlp_dims must be 3. mlp_dims={}'.format(mlp_dims))

    def CombineLastTwoDims(x):
      return x.reshape(list(x.shape[:-2]) + [x.shape[-2] * x.shape[-1]])

    return self._Seq(
        name,
        self._ApplyFnMulti(
            'transpose_matmul', lambda x, y: torch.matmul(x.transpose(-1, -2), y),
            self._GetValue('get_features', 'features'),
            self._Seq('transform_points',
                      self._SeqToKey('points_as_features', 'features',
                                     self._GetValue('get_points', 'points')),
                      self._FeaturesMLP('points_mlp', mlp_dims),
                      self._GetValue('get_transformed_points', 'features'))),
        self._Fn('reshape', CombineLastTwoDims),
        self._FC('fc', num_in_channels * mlp_dims[-1], num_out_channels))


for _name, _fn in list(vars(_PointSetRecipes).items()):
  if callable(_fn) and not _name.startswith('__') and not hasattr(ModelBuilderBase, _name):
    setattr(ModelBuilderBase, _name, _fn)
