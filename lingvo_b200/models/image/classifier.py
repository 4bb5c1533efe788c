"""Image classification tasks (reference `tasks/image/classifier.py`).

`ModelV1`: [conv → (BN) → maxpool → (dropout)]×N → fc → softmax (:100-222).
`ModelV2`: conv extractor stack → fc → softmax (:225-343).
"""

import numpy as np
import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import layers
from lingvo_b200.core import metrics
from lingvo_b200.core import py_utils
from lingvo_b200.core import schedule
from lingvo_b200.core.nested_map import NestedMap


def TopKAccuracy(k, logits, labels, weights):
  """Weighted top-k accuracy. logits [N, C], labels [N], weights [N]."""
  logits = py_utils.HasRank(logits, 2)
  topk = logits.float().topk(min(k, logits.shape[1]), dim=-1).indices
  correct = (topk == labels.long().unsqueeze(-1)).any(-1).to(weights.dtype)
  return (correct * weights).sum() / torch.clamp(weights.sum(), min=1e-8)


class BaseClassifier(base_model.BaseTask):
  """Base class for image classifiers."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax layer.')
    p.Define('add_image_summary', True, 'Image summaries in evaler/decoder.')
    return p

  def _Accuracy(self, k, logits, labels, weights):
    return TopKAccuracy(k, logits, labels, weights)

  def Decode(self, input_batch):
    return self.FPropDefaultTheta(input_batch)[0]

  def CreateDecoderMetrics(self):
    return {'num_samples_in_batch': metrics.AverageMetric()}

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    v = dec_out_dict['num_samples_in_batch'][0]
    dec_metrics_dict['num_samples_in_batch'].Update(float(v))

  def Inference(self):
    """'default' subgraph: normalized_image → logits/probs/prediction."""
    def default(normalized_image):
      theta = self.theta
      logits = self._Logits(theta, normalized_image)
      return NestedMap(logits=logits, probs=torch.softmax(logits.float(), -1),
                       prediction=logits.argmax(-1))
    return {'default': default}


class ModelV1(BaseClassifier):
  """CNNs with max-pooling followed by a softmax."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('filter_shapes', [(0, 0, 0, 0)], 'Conv filter shapes (h,w,in,out).')
    p.Define('window_shapes', [(0, 0)], 'Max pooling window shapes (h,w).')
    p.Define('fc_tpl', layers.FCLayer.Params(), 'FC template for the logits.')
    p.Define('batch_norm', False, 'Apply BN after the conv.')
    p.Define('dropout_prob', 0.0, 'Dropout applied after pooling.')
    tp = p.train
    tp.learning_rate = 1e-4
    tp.lr_schedule = (
        schedule.LinearRampupExponentialDecayScaledByNumSplitSchedule.Params()
        .Set(warmup=100, decay_start=100000, decay_end=1000000, min=0.1))
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.name
    assert len(p.filter_shapes) == len(p.window_shapes)
    shape = [None] + list(p.input.data_shape)
    conv_params, pool_params = [], []
    for i, (kernel, window) in enumerate(zip(p.filter_shapes, p.window_shapes)):
      conv_params.append(layers.ConvLayer.Params().Set(
          name='conv%d' % i, filter_shape=kernel, filter_stride=(1, 1),
          batch_norm=p.batch_norm))
      pool_params.append(layers.PoolingLayer.Params().Set(
          name='pool%d' % i, window_shape=window, window_stride=window))
    self.CreateChildren('conv', conv_params)
    self.CreateChildren('pool', pool_params)
    for i in range(len(self.conv)):
      shape = self.conv[i].OutShape(shape)
      shape = self.pool[i].OutShape(shape)
    self.CreateChild('fc', p.fc_tpl.Copy().Set(
        name='fc', input_dim=int(np.prod(shape[1:])),
        output_dim=p.softmax.input_dim))
    self.CreateChild('softmax', p.softmax)

  def _Logits(self, theta, data):
    p = self.params
    batch = data.shape[0]
    h, w, d = p.input.data_shape
    act = data.reshape(batch, h, w, d)
    for i in range(len(self.conv)):
      act, _ = self.conv[i].FProp(theta.conv[i], act)
      act = self.pool[i].FProp(theta.pool[i], act)
      if p.dropout_prob > 0.0 and not self.do_eval:
        act = torch.nn.functional.dropout(act, p.dropout_prob, training=True)
    return self.fc.FProp(theta.fc, act.reshape(batch, -1))

  def FPropTower(self, theta, input_batch):
    act = self._Logits(theta, input_batch.data)
    batch = torch.tensor(float(input_batch.data.shape[0]), device=act.device)
    labels = input_batch.label.long()
    xent = self.softmax.FProp(theta.softmax, act,
                              class_weights=input_batch.weight,
                              class_ids=labels)
    rets = {
        'loss': (xent.avg_xent, batch),
        'log_pplx': (xent.avg_xent, batch),
        'num_preds': (batch, torch.tensor(1.0)),
    }
    if self.do_eval:
      acc1 = self._Accuracy(1, xent.logits, labels, input_batch.weight)
      acc5 = self._Accuracy(5, xent.logits, labels, input_batch.weight)
      rets.update(accuracy=(acc1, batch), acc5=(acc5, batch),
                  error=(1. - acc1, batch), error5=(1. - acc5, batch))
    return rets, {'loss': xent.per_example_xent}


class ModelV2(BaseClassifier):
  """Generic extractor → fc → softmax (reference :225)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('extract', None, 'Param for the layer extracting image features.')
    p.Define('label_smoothing', 0., 'Smooth the labels towards 1/num_classes.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.name
    self.CreateChild('extract', p.extract)
    self.CreateChild('softmax', p.softmax)

  def _Features(self, theta, data):
    act = self.extract.FProp(theta.extract, data)
    if isinstance(act, tuple):
      act = act[0]
    return act.reshape(act.shape[0], -1)

  def _Logits(self, theta, data):
    """Class logits (what the serving subgraph of `BaseClassifier.Inference` exposes)."""
    return self.softmax.Logits(theta.softmax, self._Features(theta, data))

  def ComputePredictions(self, theta, input_batch):
    act = self._Features(theta, input_batch.data)
    logits = self.softmax.Logits(theta.softmax, act)
    return NestedMap(logits=logits, act=act)

  def ComputeLoss(self, theta, predictions, input_batch):
    p = self.params
    batch = torch.tensor(float(input_batch.data.shape[0]),
                         device=predictions.logits.device)
    labels = input_batch.label.long()
    num_classes = predictions.logits.shape[-1]
    probs = torch.nn.functional.one_hot(labels, num_classes).float()
    if p.label_smoothing > 0.:
      probs = probs * (1 - p.label_smoothing) + p.label_smoothing / num_classes
    xent = self.softmax.XentLossFromLogits(
        theta.softmax, predictions.logits, input_batch.weight,
        class_probabilities=probs)
    rets = {'loss': (xent.avg_xent, batch), 'log_pplx': (xent.avg_xent, batch),
            'num_preds': (batch, torch.tensor(1.0))}
    acc1 = self._Accuracy(1, predictions.logits, labels, input_batch.weight)
    acc5 = self._Accuracy(5, predictions.logits, labels, input_batch.weight)
    rets.update(accuracy=(acc1, batch), acc5=(acc5, batch),
                error=(1. - acc1, batch), error5=(1. - acc5, batch))
    return rets, {'loss': xent.per_example_xent}
