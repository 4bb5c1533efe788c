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
    p.Define('teacher_target_type', 'truth', 'truth: train on the labels; beam: on the teacher beam hypotheses.')
    p.Define('beam_search_temperature', 1.0, 'Softmax temperature T.')
    p.Define('train_teacher', False, 'Also train the teacher.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    for sub in (p.teacher, p.student):
      assert isinstance(sub.cls, type) and issubclass(sub.cls, base_model.BaseTask)
      assert not issubclass(sub.cls, DistillationTask)
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
    """teacher_target_type 'truth': both nets see the ground-truth batch. 'beam': the
    teacher's `ComputeBeamPredictions(theta, batch, temperature)` returns its predictions,
    a batch rewritten with its beam hypotheses as targets and the hypotheses' probabilities;
    the student is trained on that batch (ref :83)."""
    p = self.params

    def _Teacher(fn, *args):
      if p.train_teacher:
        return fn(*args)
      with torch.no_grad():
        return fn(*args)

    if p.teacher_target_type == 'truth':
      teacher = _Teacher(self.teacher.ComputePredictions, theta.teacher, input_batch)
      student = self.student.ComputePredictions(theta.student, input_batch)
      return NestedMap(teacher=teacher, student=student)
    if p.teacher_target_type == 'beam':
      teacher, teacher_batch, beam_prob = _Teacher(
          self.teacher.ComputeBeamPredictions, theta.teacher, input_batch,
          p.beam_search_temperature)
      student = self.student.ComputePredictions(theta.student, teacher_batch)
      return NestedMap(teacher=teacher, student=student, teacher_beam_prob=beam_prob)
    raise ValueError('teacher target type not defined properly: %s' % p.teacher_target_type)

  def ComputeDistillationLoss(self, theta, predictions, input_batch):
    """→ ({'loss': (value, weight), …}, per_example). The default is the soft-target cross
    entropy CE(softmax(teacher_logits/T), student_logits/T) averaged over positions (over the
    non-padded ones when the batch carries `tgt.paddings` / `paddings` of the logits' leading
    shape); tasks without `.logits` predictions override this (ref :151)."""
    del theta
    t = self.params.beam_search_temperature
    tl = predictions.teacher.logits.float() / t
    sl = predictions.student.logits.float() / t
    ce = -(torch.softmax(tl, -1).detach() if not self.params.train_teacher
           else torch.softmax(tl, -1))
    ce = (ce * torch.log_softmax(sl, -1)).sum(-1)
    pad = None
    for src in (input_batch.get('tgt'), input_batch):
      if src is not None and hasattr(src, 'get') and src.get('paddings') is not None and \
          tuple(src.paddings.shape) == tuple(ce.shape):
        pad = src.paddings
        break
    if pad is None:
      weight = torch.tensor(float(ce.numel()), device=ce.device)
      loss = ce.mean()
    else:
      w = 1.0 - pad.to(ce.dtype)
      weight = w.sum().clamp_min(1e-8)
      loss = (ce * w).sum() / weight
    return {'loss': (loss, weight)}, {}

  def ComputeLoss(self, theta, predictions, input_batch):
    """(1 − w) · ground-truth metrics + w · distillation metrics, w from the
    `distillation_loss_weight` schedule; with `train_teacher` the ground-truth loss is the
    weighted average of the teacher's and the student's (ref :108)."""
    p = self.params
    per_example = {}
    groundtruth, gt_per_example = self.student.ComputeLoss(theta.student, predictions.student,
                                                           input_batch)
    groundtruth = dict(groundtruth)
    groundtruth['student_groundtruth_loss'] = groundtruth['loss']
    per_example.update(gt_per_example or {})
    if p.train_teacher:
      teacher_gt, _ = self.teacher.ComputeLoss(theta.teacher, predictions.teacher, input_batch)
      groundtruth['teacher_groundtruth_loss'] = teacher_gt['loss']
      (tv, tw), (sv, sw) = teacher_gt['loss'], groundtruth['student_groundtruth_loss']
      total = tw + sw
      groundtruth['loss'] = ((tv * tw + sv * sw) / total, total)
    distill, distill_per_example = self.ComputeDistillationLoss(theta, predictions, input_batch)
    distill = dict(distill)
    distill['distillation_loss'] = distill['loss']
    per_example.update(distill_per_example or {})
    w = float(self.distillation_loss_weight.Value())
    # Every metric is reported; 'loss' is the (1 − w, w) blend.
    metrics = {}
    for k, v in groundtruth.items():
      if k != 'loss':
        metrics[k] = v
    for k, v in distill.items():
      if k != 'loss':
        metrics[k] = v
    metrics['groundtruth_loss'] = groundtruth['loss']
    gv, gw = groundtruth['loss']
    dv, _ = distill['loss']
    metrics['loss'] = ((1.0 - w) * gv + w * dv, gw)
    return metrics, per_example

  def BProp(self):
    """Only the student's variables are optimised unless `train_teacher` (ref :154): the
    teacher's are frozen (`requires_grad=False`), so the learners skip them."""
    return super().BProp()

  def Decode(self, input_batch):
    return self.student.Decode(input_batch)

  def Inference(self):
    return self.student.Inference()

  def CreateDecoderMetrics(self):
    return self.student.CreateDecoderMetrics()

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    return self.student.PostProcessDecodeOut(dec_out_dict, dec_metrics_dict)
