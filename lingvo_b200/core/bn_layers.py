"""Normalisation layers with padding awareness.

Reference `lingvo/core/bn_layers.py`: `BatchNormLayer` (padding-aware
moments, optional cross-replica stats :114-131), `CategoricalBN`,
`BatchNormLayerNoPadding`, `GroupNormLayer` (:747, incl. cumulative /
streaming mode). Cross-replica statistics use NCCL all-reduce on the small
`[C]` vectors.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core import summary_utils
from lingvo_b200.core.nested_map import NestedMap


def _AllReduceMoments(tensors, group=None):
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(
      group) > 1:
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.all_reduce(flat, group=group)
    out, off = [], 0
    for t in tensors:
      n = t.numel()
      out.append(flat[off:off + n].reshape(t.shape).to(t.dtype))
      off += n
    return out
  return tensors


def ComputeMoments(inputs, padding, reduce_over_dims, cumulative_axis=None,
                   enable_cross_replica_sum_on_tpu=False, keepdims=False):
  """Mean/variance over `reduce_over_dims` ignoring padded positions."""
  mask = 1.0 - padding.to(inputs.dtype)
  while mask.dim() < inputs.dim():
    mask = mask.unsqueeze(-1)
  x = inputs * mask
  if cumulative_axis is None:
    sum_v = x.sum(dim=reduce_over_dims, keepdim=keepdims)
    cnt = (mask.expand_as(inputs)).sum(dim=reduce_over_dims, keepdim=keepdims)
    if enable_cross_replica_sum_on_tpu:
      sum_v, cnt = _AllReduceMoments([sum_v, cnt])
    cnt = torch.clamp(cnt, min=1.0)
    mean = sum_v / cnt
    m = mean if keepdims else mean.reshape(
        [1 if i in reduce_over_dims else s for i, s in enumerate(inputs.shape)])
    sum_vv = (((inputs - m) * mask)**2).sum(dim=reduce_over_dims,
                                            keepdim=keepdims)
    if enable_cross_replica_sum_on_tpu:
      (sum_vv,) = _AllReduceMoments([sum_vv])
    var = sum_vv / cnt
    return mean, var
  # Cumulative (streaming) statistics along `cumulative_axis`.
  dims = [d for d in reduce_over_dims if d != cumulative_axis]
  sum_v = torch.cumsum(x.sum(dim=dims, keepdim=True), dim=cumulative_axis)
  cnt = torch.cumsum(mask.expand_as(inputs).sum(dim=dims, keepdim=True),
                     dim=cumulative_axis)
  cnt = torch.clamp(cnt, min=1.0)
  mean = sum_v / cnt
  sum_vv = torch.cumsum((x * x).sum(dim=dims, keepdim=True),
                        dim=cumulative_axis)
  var = torch.clamp(sum_vv / cnt - mean * mean, min=0.0)
  return mean, var


class AddingAccumulator(base_layer.Accumulator):
  """Accumulator summing sufficient statistics across pipeline micro-batches (ref :31)."""

  def __init__(self, shape, dtype, device=None):
    super().__init__()
    self.dtype, self.shape, self.device = dtype, shape, device

  def DefaultValue(self):
    return torch.zeros(list(self.shape), dtype=self.dtype, device=self.device)

  def Update(self, value):
    cur = self.GetValue()
    self.SetValue(cur + value.to(device=cur.device, dtype=self.dtype))


class BatchNormLayer(base_layer.BaseLayer):
  """Batch normalization with moving statistics."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dim', 0, 'Depth of the input/output.')
    p.Define('decay', 0.999, 'Decay in updating the mean and variance.')
    p.Define('enable_cross_replica_sum_on_tpu', True,
             'Sync moments across data-parallel ranks.')
    p.Define('use_moving_avg_in_training', False,
             'Use moving averages while training (freeze BN).')
    p.Define('freeze_bn_stats', False, 'Do not update moving stats.')
    p.Define('gamma_zero_init', False, 'Init gamma such that scale = 0.')
    p.Define('gamma_one_init', False, 'gamma var itself holds the scale.')
    p.Define('set_padded_output_to_zero', True, 'Zero padded outputs.')
    p.Define('use_fused_batch_norm_for_eval', False, 'Kept for parity.')
    p.Define('add_stats_to_moving_average_variables', None, 'Kept for parity.')
    p.Define('epsilon', 1e-3, 'Small float added to variance.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._epsilon = self.params.epsilon
    self._pending_stats = None

  def _GetWeightShape(self):
    return [self.params.dim]

  def _CreateLayerVariables(self):
    p = self.params
    pc = py_utils.WeightParams(
        shape=self._GetWeightShape(), init=py_utils.WeightInit.Constant(0.0),
        dtype=p.dtype, collections=[self.__class__.__name__ + '_vars'])
    if not p.use_moving_avg_in_training or True:
      self.CreateVariable('beta', pc)
      gamma_p = pc.Copy()
      if p.gamma_zero_init:
        gamma_p.init = py_utils.WeightInit.Constant(-1.0)
      if p.gamma_one_init:
        gamma_p.init = py_utils.WeightInit.Constant(1.0)
      self.CreateVariable('gamma', gamma_p)
    mva = py_utils.WeightParams(
        shape=self._GetWeightShape(), init=py_utils.WeightInit.Constant(0.0),
        dtype=torch.float32, collections=[self.__class__.__name__ + '_vars'])
    self.CreateVariable('moving_mean', mva, trainable=False)
    mvv = mva.Copy()
    mvv.init = py_utils.WeightInit.Constant(1.0)
    self.CreateVariable('moving_variance', mvv, trainable=False)

  def _GetBetaGamma(self, theta, inputs, **kwargs):
    p = self.params
    beta = theta.beta
    gamma = theta.gamma if p.gamma_one_init else 1.0 + theta.gamma
    return beta, gamma

  def GetCurrentMoments(self, theta):
    return theta.moving_mean, theta.moving_variance, *self._GetBetaGamma(
        theta, None)

  def ComputeAndUpdateMoments(self, theta, inputs, paddings=None, **kwargs):
    p = self.params
    if paddings is None:
      paddings = torch.zeros(list(inputs.shape[:-1]) + [1], dtype=inputs.dtype,
                             device=inputs.device)
    beta, gamma = self._GetBetaGamma(theta, inputs, **kwargs)
    if self.do_eval or p.use_moving_avg_in_training:
      mean, var = theta.moving_mean, theta.moving_variance
    else:
      rdims = list(range(inputs.dim() - 1))
      mean, var = ComputeMoments(
          inputs.float(), paddings, rdims,
          enable_cross_replica_sum_on_tpu=p.enable_cross_replica_sum_on_tpu)
      if not p.freeze_bn_stats:
        # Moving-average update applied now (eager) — the reference defers it
        # to the train op; equivalent since FProp precedes the optimizer.
        with torch.no_grad():
          mm, mv = self.vars.moving_mean, self.vars.moving_variance
          mm.sub_((mm - mean.detach().to(mm.dtype)) * (1.0 - p.decay))
          mv.sub_((mv - var.detach().to(mv.dtype)) * (1.0 - p.decay))
      summary_utils.histogram('%s_mean' % p.name, mean)
    return mean, var, beta, gamma

  def FProp(self, theta, inputs, paddings=None):
    p = self.params
    if paddings is None:
      paddings = torch.zeros(list(inputs.shape[:-1]) + [1], dtype=inputs.dtype,
                             device=inputs.device)
    mean, var, beta, gamma = self.ComputeAndUpdateMoments(theta, inputs,
                                                          paddings)
    inv = torch.rsqrt(var.float() + self._epsilon)
    scale = (inv * gamma.float()).to(inputs.dtype)
    shift = (beta.float() - mean.float() * inv * gamma.float()).to(inputs.dtype)
    out = inputs * scale + shift
    if p.set_padded_output_to_zero:
      out = py_utils.ApplyPadding(paddings, out)
    return out

  @classmethod
  def FPropMeta(cls, p, inputs, padding=None):
    return NestedMap(flops=inputs.num_elements() * 10, out_shapes=(inputs,))


class CategoricalBN(BatchNormLayer):
  """BN with per-class (domain) statistics and affine params."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('class_emb_dim', None, 'Number of classes.')
    p.use_moving_avg_in_training = False
    return p

  def _GetWeightShape(self):
    return [self.params.class_emb_dim, self.params.dim]

  def FProp(self, theta, inputs, paddings, class_emb):
    """class_emb: [B, class_emb_dim] one-hot (or soft) class membership."""
    p = self.params
    if paddings is None:
      paddings = torch.zeros(list(inputs.shape[:-1]) + [1], dtype=inputs.dtype,
                             device=inputs.device)
    emb = class_emb.to(torch.float32)
    while emb.dim() < inputs.dim():
      emb = emb.unsqueeze(1)
    beta = torch.matmul(emb, theta.beta.float())
    gamma = 1.0 + torch.matmul(emb, theta.gamma.float())
    if self.do_eval:
      mean = torch.matmul(emb, theta.moving_mean)
      var = torch.matmul(emb, theta.moving_variance)
    else:
      mask = 1.0 - paddings.float()
      w = emb.unsqueeze(-1) * mask.unsqueeze(-1)          # [..., K, 1]
      x = inputs.float().unsqueeze(-2)                     # [..., 1, D]
      rd = list(range(inputs.dim() - 1))
      cnt = torch.clamp(w.sum(dim=rd), min=1.0)            # [K, 1]
      cmean = (x * w).sum(dim=rd) / cnt                    # [K, D]
      cvar = (((x - cmean) ** 2) * w).sum(dim=rd) / cnt
      with torch.no_grad():
        mm, mv = self.vars.moving_mean, self.vars.moving_variance
        mm.sub_((mm - cmean.detach()) * (1.0 - p.decay))
        mv.sub_((mv - cvar.detach()) * (1.0 - p.decay))
      mean = torch.matmul(emb, cmean)
      var = torch.matmul(emb, cvar)
    out = (inputs.float() - mean) * torch.rsqrt(var + self._epsilon) * gamma + beta
    return py_utils.ApplyPadding(paddings, out.to(inputs.dtype))


class BatchNormLayerNoPadding(base_layer.BaseLayer):
  """BN without padding support (image models)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dim', 0, 'Depth of the input/output.')
    p.Define('decay', 0.997, 'Decay in updating the mean and variance.')
    p.Define('epsilon', 1e-3, 'Small float added to variance.')
    p.Define('bn_group_size', 1, 'Ranks per cross-replica BN group.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    pc = py_utils.WeightParams([p.dim], py_utils.WeightInit.Constant(0.0),
                               p.dtype, ['BatchNormLayerNoPadding_vars'])
    self.CreateVariable('beta', pc)
    self.CreateVariable('gamma', pc)
    mva = py_utils.WeightParams([p.dim], py_utils.WeightInit.Constant(0.0),
                                torch.float32)
    self.CreateVariable('moving_mean', mva, trainable=False)
    mvv = py_utils.WeightParams([p.dim], py_utils.WeightInit.Constant(1.0),
                                torch.float32)
    self.CreateVariable('moving_variance', mvv, trainable=False)

  def FProp(self, theta, inputs):
    p = self.params
    if self.do_eval:
      mean, var = theta.moving_mean, theta.moving_variance
    else:
      rd = list(range(inputs.dim() - 1))
      x = inputs.float()
      mean = x.mean(dim=rd)
      msq = (x * x).mean(dim=rd)
      if p.bn_group_size > 1:
        mean, msq = _AllReduceMoments([mean, msq])
        n = float(dist.get_world_size()) if dist.is_initialized() else 1.0
        mean, msq = mean / n, msq / n
      var = torch.clamp(msq - mean * mean, min=0.0)
      with torch.no_grad():
        mm, mv = self.vars.moving_mean, self.vars.moving_variance
        mm.sub_((mm - mean.detach()) * (1.0 - p.decay))
        mv.sub_((mv - var.detach()) * (1.0 - p.decay))
    inv = torch.rsqrt(var + p.epsilon) * (1.0 + theta.gamma.float())
    return (inputs.float() * inv + (theta.beta.float() - mean * inv)).to(
        inputs.dtype)

  @classmethod
  def FPropMeta(cls, p, inputs):
    return NestedMap(flops=inputs.num_elements() * 10, out_shapes=(inputs,))


class GroupNormLayer(base_layer.BaseLayer):
  """Group normalization (reference :747); rank-3 [B,T,C] or rank-4 [B,T,F,C]."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dim', 0, 'Depth of the input/output.')
    p.Define('num_groups', 32, 'Number of groups.')
    p.Define('min_group_size', 1, 'Minimum group size.')
    p.Define('cumulative', False, 'Cumulative (causal/streaming) statistics.')
    p.Define('input_rank', 4, 'Rank of input: 3 or 4.')
    p.Define('epsilon', 1e-3, 'Epsilon.')
    p.Define('set_padded_output_to_zero', True, 'Zero padded outputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.name
    assert p.num_groups > 0 and p.dim % p.num_groups == 0
    assert p.dim // p.num_groups >= p.min_group_size

  def _CreateLayerVariables(self):
    p = self.params
    shape = [1, 1, 1, p.dim] if p.input_rank == 4 else [1, 1, p.dim]
    pc = py_utils.WeightParams(shape, py_utils.WeightInit.Constant(0.0),
                               p.dtype, ['GroupNormLayer_vars'])
    self.CreateVariable('beta', pc)
    self.CreateVariable('gamma', pc)

  @property
  def group_size(self):
    return self.params.dim // self.params.num_groups

  def zero_state(self, batch_size):
    p = self.params
    z = torch.zeros([batch_size, 1, 1, p.num_groups, 1])
    return NestedMap(cached_sum=z, cached_count=z.clone(), cached_var=z.clone())

  def _Normalize(self, theta, grouped, mean, var):
    p = self.params
    x = (grouped - mean) * torch.rsqrt(var + p.epsilon)
    x = x.reshape(self._in_shape)
    return x * (1.0 + theta.gamma.float()) + theta.beta.float()

  def FProp(self, theta, inputs, paddings=None):
    p = self.params
    self._in_shape = inputs.shape
    b, t = inputs.shape[0], inputs.shape[1]
    x = inputs.float()
    if p.input_rank == 4:
      grouped = x.reshape(b, t, inputs.shape[2], p.num_groups, self.group_size)
      rdims = [1, 2, 4]
    else:
      grouped = x.reshape(b, t, p.num_groups, self.group_size)
      rdims = [1, 3]
    if paddings is None:
      pad = torch.zeros([b, t], device=inputs.device)
    else:
      pad = paddings.reshape(b, t).float()
    mean, var = ComputeMoments(grouped, pad, rdims,
                               cumulative_axis=1 if p.cumulative else None,
                               keepdims=True)
    out = self._Normalize(theta, grouped, mean, var).to(inputs.dtype)
    if paddings is None:
      return out
    if p.set_padded_output_to_zero:
      out = py_utils.ApplyPadding(pad, out)
    return out, paddings

  def StreamStep(self, theta, inputs, paddings, state0):
    """Cumulative statistics carried across chunks (streaming inference)."""
    p = self.params
    assert p.cumulative
    self._in_shape = inputs.shape
    b, t = inputs.shape[0], inputs.shape[1]
    x = inputs.float()
    if p.input_rank == 4:
      grouped = x.reshape(b, t, inputs.shape[2], p.num_groups, self.group_size)
      dims = [2, 4]
    else:
      grouped = x.reshape(b, t, 1, p.num_groups, self.group_size)
      dims = [2, 4]
    mask = (1.0 - paddings.reshape(b, t).float()).reshape(b, t, 1, 1, 1)
    gx = grouped * mask
    s = torch.cumsum(gx.sum(dim=dims, keepdim=True), 1) + state0.cached_sum.to(x.device)
    c = torch.cumsum((mask.expand_as(grouped)).sum(dim=dims, keepdim=True), 1) + (
        state0.cached_count.to(x.device))
    cc = torch.clamp(c, min=1.0)
    mean = s / cc
    v = torch.cumsum((gx * gx).sum(dim=dims, keepdim=True), 1) + (
        state0.cached_var.to(x.device))
    var = torch.clamp(v / cc - mean * mean, min=0.0)
    out = ((grouped - mean) * torch.rsqrt(var + p.epsilon)).reshape(inputs.shape)
    out = out * (1.0 + theta.gamma.float()) + theta.beta.float()
    out = py_utils.ApplyPadding(paddings.reshape(b, t), out.to(inputs.dtype))
    state1 = NestedMap(cached_sum=s[:, -1:], cached_count=c[:, -1:],
                       cached_var=v[:, -1:])
    return out, paddings, state1

  @classmethod
  def FPropMeta(cls, p, inputs, paddings=None):
    return NestedMap(flops=inputs.num_elements() * 10, out_shapes=(inputs,))
