"""Multi-task models (shared encoders, regex-shared variables) and distillation."""

import torch

from lingvo_b200.core import base_input_generator, base_model, distillation_task, layers
from lingvo_b200.core import multitask_model, optimizer, schedule
from lingvo_b200.core.nested_map import NestedMap


class _ClsInput(base_input_generator.BaseInputGenerator):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('flip', False, 'Second task: labels are inverted.')
    return p

  def _InputBatch(self):
    x = torch.randn(32, 6)
    y = (x[:, 0] + x[:, 1] > 0).long()
    return NestedMap(x=x, y=1 - y if self.params.flip else y)


class _ToyTask(base_model.BaseTask):
  """enc (FC) → dec (FC → 2 logits), softmax cross-entropy."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('hidden', 8, 'Encoder width.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('enc', layers.FCLayer.Params().Set(name='enc', input_dim=6, output_dim=p.hidden,
                                                        activation='TANH'))
    self.CreateChild('dec', layers.FCLayer.Params().Set(name='dec', input_dim=p.hidden, output_dim=2,
                                                        activation='NONE'))

  def ComputePredictions(self, theta, input_batch):
    return NestedMap(logits=self.dec.FProp(theta.dec, self.enc.FProp(theta.enc, input_batch.x)))

  def ComputeLoss(self, theta, predictions, input_batch):
    loss = torch.nn.functional.cross_entropy(predictions.logits.float(), input_batch.y)
    acc = (predictions.logits.argmax(-1) == input_batch.y).float().mean()
    w = torch.tensor(float(input_batch.y.shape[0]))
    return {'loss': (loss, w), 'accuracy': (acc, w)}, {}


def _TaskParams(name, flip=False, hidden=8):
  p = _ToyTask.Params().Set(name=name, hidden=hidden)
  p.input = _ClsInput.Params().Set(name='in', batch_size=32, flip=flip)
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 2e-2
  p.train.lr_schedule = schedule.Constant.Params()
  return p


def _Model(cls, **kw):
  mp = cls.Params().Set(name='mt', **kw)
  mp.task_params.Define('a', _TaskParams('a'), '')
  mp.task_params.Define('b', _TaskParams('b', flip=True), '')
  mp.task_probs.Define('a', 0.5, '')
  mp.task_probs.Define('b', 0.5, '')
  return mp.Instantiate()


def test_shared_encoder_model_shares_weights_and_trains_both_tasks():
  torch.manual_seed(0)
  m = _Model(multitask_model.SharedEncoderModel)
  a, b = m.GetTask('a'), m.GetTask('b')
  assert a.enc is b.enc and a.dec is not b.dec
  assert a.enc.vars.w is b.enc.vars.w
  seen = set()
  for _ in range(150):
    m.ConstructFPropBPropGraph()
    seen.add(m.last_task_name)
  assert seen == {'a', 'b'}
  for t in (a, b):
    metrics, _ = t.FPropDefaultTheta()
    assert float(metrics['accuracy'][0]) > 0.85            # one encoder serves both label conventions


def test_shared_encoder_decoder_model():
  m = _Model(multitask_model.SharedEncoderDecoderModel)
  a, b = m.GetTask('a'), m.GetTask('b')
  assert a.enc is b.enc and a.dec is b.dec


def test_regex_shared_variables():
  m = _Model(multitask_model.RegExSharedVariableModel,
             variable_renaming_rules=[(r'^(?:mt/)?[ab]/enc/(.*)$', 'shared/enc/%s')])
  a, b = m.GetTask('a'), m.GetTask('b')
  assert a.enc.vars.w.data_ptr() == b.enc.vars.w.data_ptr()      # same storage
  assert a.dec.vars.w.data_ptr() != b.dec.vars.w.data_ptr()
  with torch.no_grad():
    a.enc.vars.w.add_(1.0)
  torch.testing.assert_close(a.enc.vars.w, b.enc.vars.w)


def test_distillation_student_follows_teacher():
  torch.manual_seed(0)
  p = distillation_task.DistillationTask.Params().Set(name='distill')
  p.input = _ClsInput.Params().Set(name='in', batch_size=32)
  p.teacher = _TaskParams('teacher', hidden=16)
  p.student = _TaskParams('student', hidden=4)
  p.distillation_loss_weight = schedule.Constant.Params().Set(value=1.0)   # soft targets only
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 3e-2
  p.train.lr_schedule = schedule.Constant.Params()
  task = p.Instantiate()
  # give the (frozen) teacher a decisive, known solution: logit = ±4·(x0 + x1)
  with torch.no_grad():
    tw = task.teacher.enc.vars.w; tw.zero_(); tw[0, 0] = tw[1, 0] = 0.2
    dw = task.teacher.dec.vars.w; dw.zero_(); dw[0, 0] = -20.0; dw[0, 1] = 20.0
  before = task.teacher.enc.vars.w.detach().clone()
  first = None
  for _ in range(200):
    metrics, _ = task.TrainStep()
    first = first if first is not None else float(metrics['distillation_loss'][0])
  assert float(metrics['distillation_loss'][0]) < 0.7 * first
  assert float(metrics['accuracy'][0]) > 0.9                      # learned from soft targets alone
  torch.testing.assert_close(task.teacher.enc.vars.w.detach(), before)   # teacher untouched
  assert not task.teacher.enc.vars.w.requires_grad


def test_distillation_train_teacher_blends_ground_truth_and_delegates_decoding():
  torch.manual_seed(0)
  p = distillation_task.DistillationTask.Params().Set(name='distill2', train_teacher=True)
  p.input = _ClsInput.Params().Set(name='in', batch_size=16)
  p.teacher = _TaskParams('teacher', hidden=8)
  p.student = _TaskParams('student', hidden=4)
  p.distillation_loss_weight = schedule.Constant.Params().Set(value=0.25)
  p.train.optimizer = optimizer.Adam.Params()
  p.train.learning_rate = 1e-2
  p.train.lr_schedule = schedule.Constant.Params()
  task = p.Instantiate()
  metrics, _ = task.FPropDefaultTheta()
  for k in ('student_groundtruth_loss', 'teacher_groundtruth_loss', 'groundtruth_loss',
            'distillation_loss', 'loss'):
    assert k in metrics, k
  (tv, tw), (sv, sw) = metrics['teacher_groundtruth_loss'], metrics['student_groundtruth_loss']
  gt = (float(tv) * float(tw) + float(sv) * float(sw)) / (float(tw) + float(sw))
  assert abs(float(metrics['groundtruth_loss'][0]) - gt) < 1e-5
  want = 0.75 * gt + 0.25 * float(metrics['distillation_loss'][0])
  assert abs(float(metrics['loss'][0]) - want) < 1e-5
  before = task.teacher.enc.vars.w.detach().clone()
  task.BProp()
  assert not torch.equal(task.teacher.enc.vars.w.detach(), before)      # teacher trains too
  # custom distillation loss hook
  class _L2(distillation_task.DistillationTask):
    def ComputeDistillationLoss(self, theta, predictions, input_batch):
      d = (predictions.teacher.logits.detach() - predictions.student.logits) ** 2
      return {'loss': (d.mean(), torch.tensor(float(d.shape[0])))}, {'sq': d}
  p2 = p.Copy().Set(name='distill3', train_teacher=False)
  p2.cls = _L2
  t2 = p2.Instantiate()
  m2, per_ex = t2.FPropDefaultTheta()
  assert 'sq' in per_ex and float(m2['distillation_loss'][0]) >= 0
  import pytest
  bad = p.Copy().Set(name='distill4', teacher_target_type='nope')
  with pytest.raises(ValueError):
    bad.Instantiate().FPropDefaultTheta()
