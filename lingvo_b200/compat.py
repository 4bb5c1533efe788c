"""Compatibility surface (ref `lingvo/compat.py`, the TF1/TF2 shim imported everywhere as
`import lingvo.compat as tf`).

There is no TensorFlow here; this module offers the small set of names that user
configs and scripts most often reach for through `compat` (dtypes, logging, flags,
io helpers) mapped onto their PyTorch / stdlib equivalents, so ported experiment
files keep working with `from lingvo_b200 import compat as tf`.
"""
import glob as _glob
import os as _os

import torch as _torch
from absl import flags  # noqa: F401
from absl import logging  # noqa: F401

float32 = _torch.float32
float64 = _torch.float64
float16 = _torch.float16
bfloat16 = _torch.bfloat16
int32 = _torch.int32
int64 = _torch.int64
int8 = _torch.int8
uint8 = _torch.uint8
bool = _torch.bool  # pylint: disable=redefined-builtin
complex64 = _torch.complex64
Tensor = _torch.Tensor


class _Gfile:
  """`tf.io.gfile` subset on the local filesystem."""

  @staticmethod
  def Exists(path):
    return _os.path.exists(path)
  exists = Exists

  @staticmethod
  def Glob(pattern):
    return sorted(_glob.glob(pattern))
  glob = Glob

  @staticmethod
  def MakeDirs(path):
    _os.makedirs(path, exist_ok=True)
  makedirs = MakeDirs

  @staticmethod
  def IsDirectory(path):
    return _os.path.isdir(path)
  isdir = IsDirectory

  @staticmethod
  def ListDirectory(path):
    return sorted(_os.listdir(path))
  listdir = ListDirectory

  @staticmethod
  def Remove(path):
    _os.remove(path)
  remove = Remove

  @staticmethod
  def Rename(src, dst, overwrite=False):
    if overwrite:
      _os.replace(src, dst)
    else:
      _os.rename(src, dst)
  rename = Rename

  @staticmethod
  def GFile(path, mode='r'):
    return open(path, mode)


class _Io:
  gfile = _Gfile


io = _Io
gfile = _Gfile


def executing_eagerly():  # pylint: disable=invalid-name
  return True


def enable_eager_execution():  # pylint: disable=invalid-name
  pass


def disable_v2_behavior():  # pylint: disable=invalid-name
  pass
