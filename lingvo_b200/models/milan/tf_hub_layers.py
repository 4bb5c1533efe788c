"""Pretrained image towers (ref `lingvo/tasks/milan/tf_hub_layers.py`).

The reference wraps TF-Hub modules (Inception / EfficientNet feature vectors). There is no
TF-Hub here; `ImageModule` is an in-repo residual conv tower with the same contract —
`[B, 3, S, S]` (or `[B, S, S, 3]`) images in, one feature vector per image out, optional
batch-norm statistics updates while fine-tuning — whose weights can be loaded from a
`module_path` checkpoint (torch `state_dict` file or tensor-bundle prefix).
`ImageModuleV2` is the Keras-style variant: a `trainable` switch freezes the whole tower.
"""

from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils

EFFICIENTNET_B4_INPUT_SHAPE = 380
EFFICIENTNET_B4_OUTPUT_FEATURE_DIM = 1792


class ImageModule(base_layer.BaseLayer):
  """ref :110."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('module_path', '', 'Optional checkpoint with the tower weights.')
    p.Define('signature', 'default', 'Kept for parity.')
    p.Define('training_graph_tags', {'train'}, 'Kept for parity.')
    p.Define('eval_graph_tags', set(), 'Kept for parity.')
    p.Define('run_update_ops', True, 'Update batch-norm statistics while fine-tuning.')
    p.Define('stage_channels', [32, 64, 128, 256], 'Channels of each stride-2 stage.')
    p.Define('blocks_per_stage', 1, 'Residual blocks per stage.')
    p.Define('output_dim', 512, 'Feature vector size.')
    p.Define('bn_decay', 0.99, 'Batch-norm moving-average decay.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    wp = lambda shape, init: py_utils.WeightParams(shape, init, p.dtype)
    he = lambda fan_in: py_utils.WeightInit.Gaussian((2.0 / fan_in) ** 0.5)
    cin = 3
    self._convs = []
    for s, c in enumerate(p.stage_channels):
      for b in range(p.blocks_per_stage + 1):          # block 0 = strided entry conv
        name = 's%d_b%d' % (s, b)
        i = cin if b == 0 else c
        self.CreateVariable(name + '_w', wp([c, i, 3, 3], he(i * 9)))
        self.CreateVariable(name + '_gamma', wp([c], py_utils.WeightInit.Constant(1.0)))
        self.CreateVariable(name + '_beta', wp([c], py_utils.WeightInit.Constant(0.0)))
        self.CreateVariable(name + '_mean', wp([c], py_utils.WeightInit.Constant(0.0)),
                            trainable=False)
        self.CreateVariable(name + '_var', wp([c], py_utils.WeightInit.Constant(1.0)),
                            trainable=False)
        self._convs.append((name, 2 if b == 0 else 1, b > 0))
      cin = c
    self.CreateVariable('head_w', wp([cin, p.output_dim], he(cin)))
    self.CreateVariable('head_b', wp([p.output_dim], py_utils.WeightInit.Constant(0.0)))

  def _InstantiateSelfAndChildren(self):
    super()._InstantiateSelfAndChildren()
    p = self.params
    if p.module_path and os.path.exists(p.module_path):
      state = torch.load(p.module_path, map_location='cpu')
      with torch.no_grad():
        for k, v in self.vars.items():
          if k in state:
            v.copy_(state[k])

  @property
  def losses(self):
    return []

  def _Trainable(self):
    return True

  def FProp(self, theta, images):
    p = self.params
    x = images
    if x.dim() == 4 and x.shape[-1] == 3 and x.shape[1] != 3:
      x = x.permute(0, 3, 1, 2)
    x = x.to(theta.head_w.dtype)
    update = (not self.do_eval) and p.run_update_ops and self._Trainable()
    for name, stride, residual in self._convs:
      y = F.conv2d(x, theta[name + '_w'], stride=stride, padding=1)
      mean_v, var_v = self.vars[name + '_mean'], self.vars[name + '_var']
      y = F.batch_norm(y, mean_v.data, var_v.data, theta[name + '_gamma'],
                       theta[name + '_beta'], training=update, momentum=1.0 - p.bn_decay)
      y = F.silu(y)
      x = x + y if residual else y
    feat = x.mean((2, 3))
    return torch.matmul(feat, theta.head_w) + theta.head_b


class ImageModuleV2(ImageModule):
  """ref :186: `trainable=False` freezes the tower (weights and BN statistics)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('trainable', True, 'Fine-tune the tower.')
    return p

  def _Trainable(self):
    return self.params.trainable

  def _InstantiateSelfAndChildren(self):
    super()._InstantiateSelfAndChildren()
    if not self.params.trainable:
      for v in self.vars.Flatten():
        v.requires_grad_(False)


def EfficientNetB4Params():
  """A tower with EfficientNet-B4's interface: 380×380 input, 1792-d features (ref :253)."""
  return ImageModuleV2.Params().Set(
      name='efficientnet_b4', stage_channels=[48, 32, 56, 112, 272], blocks_per_stage=2,
      output_dim=EFFICIENTNET_B4_OUTPUT_FEATURE_DIM)
