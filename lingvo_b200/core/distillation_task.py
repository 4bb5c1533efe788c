"""Teacher/student distillation (ref `lingvo/core/distillation_task.py`).

loss = w_gt · student_loss + w_distill · CE(softmax(teacher_logits/T), student_logits/T).
The teacher runs under `no_grad` and is excluded from training
(`train_teacher=False`)."""

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import schedule
from lingvo_b200.core.nested_map import NestedMap


class DistillationTask(base_model.BaseTask):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('teacher', None, 'Teacher task params.')
    p.Define('student', None, 'Student task params.')
    p.Define('distillation_loss_weight', schedule.Constant.Params().Set(value=1.0),
             'Schedule of the distillation-loss weight.')
    p.Define('teacher_target_type', 'truth', 'truth | beam (kept for parity).')
    p.Define('beam_search_temperature', 1.0, 'Softmax temperature T.')
    p.Define('train_teacher', False, 'Also train the teacher.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    for sub in (p.teacher, p.student):
      sub.input = p.input
      sub.train = p.train.Copy() if hasattr(p.train, 'Copy') else p.train
    self.CreateChild('teacher', p.teacher)
    self.CreateChild('student', p.student)
    self.CreateChild('distillation_loss_weight', p.distillation_loss_weight)

  def _CreateChildrenVariables(self):
    # Variables are materialised after __init__; the teacher is frozen once they exist.
    super()._CreateChildrenVariables()
    if not self.params.train_teacher:
      for v in self.teacher.vars.Flatten():
        v.requires_grad_(False)

  def ComputePredictions(self, theta, input_batch):
    p = self.params
    if p.train_teacher:
      teacher = self.teacher.ComputePredictions(theta.teacher, input_batch)
    else:
      with torch.no_grad():
        teacher = self.teacher.ComputePredictions(theta.teacher, input_batch)
    student = self.student.ComputePredictions(theta.student, input_batch)
    return NestedMap(teacher=teacher, student=student)

  def ComputeLoss(self, theta, predictions, input_batch):
    p = self.params
    metrics, per_ex = self.student.ComputeLoss(theta.student, predictions.student, input_batch)
    t = p.beam_search_temperature
    tl = predictions.teacher.logits.float() / t
    sl = predictions.student.logits.float() / t
    soft = -(torch.softmax(tl, -1) * torch.log_softmax(sl, -1)).sum(-1).mean()
    w = float(self.distillation_loss_weight.Value())
    gt, gw = metrics['loss']
    metrics['groundtruth_loss'] = (gt, gw)
    metrics['distillation_loss'] = (soft, gw)
    metrics['loss'] = ((1.0 - w) * gt + w * soft, gw)
    return metrics, per_ex
