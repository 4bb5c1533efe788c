"""MNIST input generators (reference `tasks/image/input_generator.py:26-102`)."""

import os

import numpy as np
import torch

from lingvo_b200.core import base_input_generator


class _MnistInputBase(base_input_generator.BaseTinyDatasetInput):
  """Base input params for MNIST."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.data_dtype = torch.uint8
    p.data_shape = (28, 28, 1)
    p.label_dtype = torch.uint8
    return p


class MnistTrainInput(_MnistInputBase):
  """MNIST training set."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.data = 'x_train'
    p.label = 'y_train'
    p.num_samples = 60000
    p.batch_size = 256
    p.repeat = True
    return p


class MnistTestInput(_MnistInputBase):
  """MNIST test set."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.data = 'x_test'
    p.label = 'y_test'
    p.num_samples = 10000
    p.batch_size = 256
    p.repeat = False
    return p


def FakeMnistData(tmpdir, train_size=60000, test_size=10000):
  """Writes a fake MNIST data file (reference :84-102); returns its path."""
  rng = np.random.RandomState(0)
  data_path = os.path.join(tmpdir, 'ckpt.npz')
  np.savez(data_path,
           x_train=rng.randint(0, 256, size=(train_size, 28, 28, 1)).astype(np.uint8),
           y_train=rng.randint(0, 10, size=(train_size,)).astype(np.uint8),
           x_test=rng.randint(0, 256, size=(test_size, 28, 28, 1)).astype(np.uint8),
           y_test=rng.randint(0, 10, size=(test_size,)).astype(np.uint8))
  return data_path
