"""AdaGraft: step *magnitude* from one optimizer, step *direction* from another
(ref `lingvo/core/adagraft.py` + wrapper `optimizer.py:803`; "Disentangling Adaptive Gradient
Methods from Learning Rates", Agarwal et al., arXiv:2002.11803).

Per step and per tensor (ref `_internal_apply_dense` :93):
  1. remember the weights (`scratch_copy`),
  2. run the magnitude optimizer in place, measure ‖m_step‖, put the weights back,
  3. run the direction optimizer in place, measure d_step and ‖d_step‖,
  4. w ← w_old + (‖m_step‖ / ‖d_step‖)·d_step     (0 when ‖d_step‖ = 0).
With `use_global_norm` the ratio is taken between the global l2 norms over all tensors
(ref `_finish` :146). `direction_optimizer_lr` gives the direction optimizer a constant
learning rate of its own; `diagnostic` records per-tensor and global step norms as summaries.

Both child optimizers keep their own slots (moments, accumulators) exactly as if they ran
alone; they see the same gradients. The grafting itself is a few norm reductions and one
axpy per tensor, all on the device and without host synchronisation (the norms stay device
scalars).
"""

from __future__ import annotations

import torch

from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core import summary_utils


class AdaGraft(optimizer.Base):
  """Combines the step size of one optimizer with the direction of another."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('magnitude_optimizer', None, 'Optimizer params providing the step size.')
    p.Define('direction_optimizer', None, 'Optimizer params providing the step direction.')
    p.Define('direction_optimizer_lr', None,
             'Constant learning rate of the direction optimizer; None: the scheduled lr for '
             'both.')
    p.Define('use_global_norm', False, 'Whether to graft the global l2 norm.')
    p.Define('diagnostic', False, 'Whether to record norm measurements.')
    p.name = 'AdaGraft'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    mag = p.magnitude_optimizer if p.magnitude_optimizer is not None else optimizer.SGD.Params()
    dire = (p.direction_optimizer if p.direction_optimizer is not None
            else optimizer.Adam.Params())
    self.CreateChild('_mag', mag.Copy().Set(name=(mag.name or 'mag') + '_magnitude'))
    self.CreateChild('_dir', dire.Copy().Set(name=(dire.name or 'dir') + '_direction'))
    self.m_step_norm = {}
    self.d_step_norm = {}

  @property
  def magnitude_optimizer(self):
    return self._mag

  @property
  def direction_optimizer(self):
    return self._dir

  def Apply(self, lr, var_grad, grad_scale=None):
    p = self.params
    pairs = optimizer._Pairs(var_grad)   # pylint: disable=protected-access
    if not pairs:
      return
    variables = [v for v, _ in pairs]
    if grad_scale is not None:
      gs = grad_scale.reshape(())
      pairs = [(v, torch.where(gs == 0, torch.zeros_like(g), g * gs.to(g.dtype)))
               for v, g in pairs]
    vg = [py_utils.VarGrad(v, g) for v, g in pairs]
    dir_lr = lr if p.direction_optimizer_lr is None else p.direction_optimizer_lr
    with torch.no_grad():
      scratch = [self._Slot(v, 'scratch_copy') for v in variables]
      for s, v in zip(scratch, variables):
        s.copy_(v)
    self._mag.Apply(lr, vg)
    with torch.no_grad():
      m_norm = []
      for s, v in zip(scratch, variables):
        m_norm.append((v.detach().float() - s.float()).norm())
        v.copy_(s)
    self._dir.Apply(dir_lr, vg)
    with torch.no_grad():
      d_steps = [v.detach() - s for s, v in zip(scratch, variables)]
      d_norm = [d.float().norm() for d in d_steps]
      if p.use_global_norm:
        mg = torch.sqrt(torch.stack([n * n for n in m_norm]).sum())
        dg = torch.sqrt(torch.stack([n * n for n in d_norm]).sum())
        ratio = [mg / dg.clamp_min(1e-30)] * len(variables)
      else:
        ratio = [m / d.clamp_min(1e-30) for m, d in zip(m_norm, d_norm)]
      for v, s, d, dn, r in zip(variables, scratch, d_steps, d_norm, ratio):
        step = torch.where(dn > 0, r.to(d.dtype) * d, torch.zeros_like(d))
        v.copy_(s + step)
      py_utils.RefreshComputeCopies(variables)
      if p.diagnostic or p.use_global_norm:
        for v, m, d in zip(variables, m_norm, d_norm):
          key = optimizer._VarKey(v)   # pylint: disable=protected-access
          self.m_step_norm[key] = m
          self.d_step_norm[key] = d
    self._step_count += 1
    if p.add_summary_in_apply:
      self.AddSummary(lr, self, var_grad)

  def AddSummary(self, lr, optimizer_obj, var_grad):   # pylint: disable=arguments-renamed
    summary_utils.scalar('adagraft_lr', lr)
    if not self.params.diagnostic:
      return
    m_tot, d_tot = 0.0, 0.0
    for key in self.m_step_norm:
      summary_utils.scalar('optimizer/m_step_norm_%s' % key, self.m_step_norm[key])
      summary_utils.scalar('optimizer/d_step_norm_%s' % key, self.d_step_norm[key])
      m_tot = m_tot + self.m_step_norm[key] ** 2
      d_tot = d_tot + self.d_step_norm[key] ** 2
    if self.m_step_norm:
      summary_utils.scalar('optimizer/m_step_norm', m_tot ** 0.5)
      summary_utils.scalar('optimizer/d_step_norm', d_tot ** 0.5)
      summary_utils.scalar('optimizer/norm_correction', (m_tot / (d_tot + 1e-60)) ** 0.5)

  # checkpointing: own scratch is transient; the children carry the real state
  def GetOptimizerSlots(self):
    out = {k: v for k, v in super().GetOptimizerSlots().items()
           if not k.endswith('/scratch_copy')}
    out.update(self._mag.GetOptimizerSlots())
    out.update(self._dir.GetOptimizerSlots())
    return out

  def LoadOptimizerSlots(self, tensors):
    used = super().LoadOptimizerSlots(tensors)
    used += self._mag.LoadOptimizerSlots(tensors)
    used += self._dir.LoadOptimizerSlots(tensors)
    return used

  def to(self, device=None, dtype=None):   # pylint: disable=invalid-name
    super().to(device, dtype)
    self._mag.to(device, dtype)
    self._dir.to(device, dtype)
    return self


AdaGraftOptimizer = AdaGraft
