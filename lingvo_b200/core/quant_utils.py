"""Fake-quantization domains and `QuantizableLayer`.

Reference `lingvo/core/quant_utils.py` (1839 LoC): `QuantizableLayer` mixin
(`QWeight/QAct/QTensor/QRAct/…` hooks), `QDomain` base,
`SymmetricScheduledClipQDomain`, `PassiveAsymQDomain`, clipping-cap schedule
`FakeQuantizationSchedule`. Layers call the hooks unconditionally; without a
qdomain they are identities, so non-quantised models pay nothing.
"""

from __future__ import annotations

import math

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def FakeQuant(x, min_v, max_v, bits=8, narrow=False):
  """Straight-through fake quantisation of `x` into [min_v, max_v]."""
  levels = 2**bits - (2 if narrow else 1)
  scale = (max_v - min_v) / levels
  scale = torch.clamp(torch.as_tensor(scale, dtype=x.dtype, device=x.device),
                      min=1e-12)
  q = torch.round((torch.clamp(x, min_v, max_v) - min_v) / scale) * scale + min_v
  return x + (q - x).detach()


class QuantizableLayer(base_layer.BaseLayer):
  """Layer base with quantisation hooks."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('qdomain', py_utils.Params(), 'Container for quantization domains.')
    p.qdomain.Define('default', None, 'Default quantization domain.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._tracked_tensors = {}
    self._qstate = None
    p = self.params
    for name, qp in p.qdomain.IterParams():
      if qp is not None:
        self.CreateChild('qdomain_' + name, qp.Copy().Set(
            name='qdomain_' + name))

  def _QDomain(self, domain='default'):
    key = 'qdomain_' + domain
    if key in self.children:
      return self.children[key]
    if 'qdomain_default' in self.children:
      return self.children['qdomain_default']
    return None

  def TrackQTensor(self, *t_names, **kwargs):
    for n in t_names:
      self._tracked_tensors[n] = kwargs.get('domain', 'default')

  TrackQActs = TrackQTensor
  TrackQWeight = TrackQTensor

  def QWeight(self, w, domain='default'):
    qd = self._QDomain(domain)
    return qd.QuantizeWeight(w) if qd is not None else w

  def ToAqtWeight(self, w_name, w, feature_axis=-1, expected_scale_shape=None):
    return w

  def FromAqtWeight(self, w_name, out, merge_feature_axes=False):
    return out

  def QAct(self, act_name, act, eval_only=False):
    qd = self._QDomain(self._tracked_tensors.get(act_name, 'default'))
    if qd is None:
      return act
    return qd.QuantizeAct(act_name, act, eval_only=eval_only)

  QTensor = QAct
  QRAct = lambda self, act, dist, domain='default': act  # pylint: disable=invalid-name

  def QRSoftmax(self, logits, dim=-1, domain='softmax'):
    return torch.softmax(logits, dim=dim)

  def QRTanh(self, x, domain='fullyconnected'):
    return torch.tanh(x)

  def QRSigmoid(self, x, domain='fullyconnected'):
    return torch.sigmoid(x)

  def QRRelu(self, x, domain='default'):
    return torch.relu(x)

  def QTensorMulti(self, t_name, *ts):
    return ts

  def QMatmul(self, lhs, rhs, **kwargs):
    return torch.matmul(lhs, rhs)

  def QConv1D(self, *args, **kwargs):
    raise NotImplementedError()

  def GetQDomainParams(self, domain='default'):
    p = self.params.qdomain
    return p.Get(domain) if domain in p else p.default


class QDomain(base_layer.BaseLayer):
  """Base class for a quantization domain."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'qdomain'
    return p

  def QuantizeWeight(self, w):
    return w

  def QuantizeAct(self, act_name, act, eval_only=False):
    return act

  def FProp(self, theta, x):
    return x

  @property
  def bits(self):
    return 8


class FakeQuantizationSchedule(base_layer.BaseLayer):
  """Clipping-cap schedule: ramps clip range then enables quantisation."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'FQSchedule'
    p.Define('clip_start_step', 0, 'Step at which clipping starts.')
    p.Define('clip_end_step', -1, 'Step at which the clip cap reaches end_cap.')
    p.Define('quant_start_step', -1, 'Step at which quantisation starts.')
    p.Define('start_cap', 8.0, 'Initial clipping cap.')
    p.Define('end_cap', 1.0, 'Final clipping cap.')
    return p

  def Value(self, step=None):
    p = self.params
    t = float(py_utils.GetGlobalStep() if step is None else step)
    if p.clip_end_step <= p.clip_start_step:
      cap = p.end_cap
    else:
      r = min(max((t - p.clip_start_step) /
                  (p.clip_end_step - p.clip_start_step), 0.0), 1.0)
      cap = p.start_cap + r * (p.end_cap - p.start_cap)
    quant = p.quant_start_step >= 0 and t >= p.quant_start_step
    return cap, quant


class SymmetricScheduledClipQDomain(QDomain):
  """Symmetric clip → fake-quant with a scheduled cap."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cc_schedule', FakeQuantizationSchedule.Params(), 'Cap schedule.')
    p.Define('bits', 8, 'Quantisation bits.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('cc_schedule', self.params.cc_schedule)

  def _Q(self, x):
    cap, quant = self.cc_schedule.Value()
    x = torch.clamp(x, -cap, cap)
    if quant:
      x = FakeQuant(x, -cap, cap, self.params.bits, narrow=True)
    return x

  def QuantizeWeight(self, w):
    return self._Q(w)

  def QuantizeAct(self, act_name, act, eval_only=False):
    if eval_only and not self.do_eval:
      return act
    return self._Q(act)


class PassiveAsymQDomain(QDomain):
  """Tracks running min/max of activations (EMA) and fake-quantises."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('bits', 8, 'Default quantized bit depth.')
    p.Define('ema_decay', 0.99, 'Moving-average decay.')
    p.Define('delay_start_steps', 0, 'Delay quantisation until this step.')
    p.Define('quantize_weight_epsilon', 0.0, 'Epsilon for weight range.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._ranges = {}

  def QuantizeWeight(self, w):
    if py_utils.GetGlobalStep() < self.params.delay_start_steps:
      return w
    lo, hi = w.detach().min(), w.detach().max()
    return FakeQuant(w, lo, hi, self.params.bits)

  def QuantizeAct(self, act_name, act, eval_only=False):
    p = self.params
    lo, hi = act.detach().min().float(), act.detach().max().float()
    if act_name in self._ranges and not self.do_eval:
      plo, phi = self._ranges[act_name]
      lo = plo * p.ema_decay + lo * (1 - p.ema_decay)
      hi = phi * p.ema_decay + hi * (1 - p.ema_decay)
    if not self.do_eval:
      self._ranges[act_name] = (lo, hi)
    elif act_name in self._ranges:
      lo, hi = self._ranges[act_name]
    if py_utils.GetGlobalStep() < p.delay_start_steps:
      return act
    return FakeQuant(act, lo.to(act.dtype), hi.to(act.dtype), p.bits)
