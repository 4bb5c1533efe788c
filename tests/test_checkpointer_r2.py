"""SaverWrapper / eager checkpointer flavours / numpy bundles (ref checkpointer_test.py,
saver_test.py)."""
import os

import numpy as np
import pytest
import torch

from lingvo_b200.core import checkpointer
from lingvo_b200.core import hyperparams
from lingvo_b200.core import saver as saver_lib
from lingvo_b200.core.nested_map import NestedMap


def test_sort_checkpoint_paths():
  assert checkpointer.SortCheckpointPaths(['a/ckpt-00000010', 'a/ckpt-00000002', 'a/ckpt-9']) == [
      'a/ckpt-00000002', 'a/ckpt-9', 'a/ckpt-00000010']


def _TrainParams(**kw):
  p = hyperparams.Params()
  for k, v in dict(max_steps=100, tpu_steps_per_loop=10, checkpoint_finite_check=True,
                   save_max_to_keep=2, save_keep_checkpoint_every_n_hours=None).items():
    p.Define(k, kw.get(k, v), '')
  return p


def test_saver_wrapper_roundtrip_and_sanity_checks(tmp_path):
  w, b = torch.randn(3, 2), torch.randn(2)
  sw = checkpointer.SaverWrapper(str(tmp_path), _TrainParams(),
                                 variables_to_restore_dict={'m/w': w, 'm/b': b})
  path = sw.Save(None, 7)
  assert path.endswith('ckpt-00000007') and os.path.exists(path + '.index')
  want_w = w.clone()
  w.zero_(); b.zero_()
  assert sw.Restore(None, path) == path
  torch.testing.assert_close(w, want_w)
  for step in (8, 9):
    sw.Save(None, step)
  sw.Sync()
  assert [os.path.basename(p) for p in saver_lib.AllCheckpoints(str(tmp_path))] == [
      'ckpt-00000008', 'ckpt-00000009']                      # keep_latest_n = 2
  w[0, 0] = float('nan')
  with pytest.raises(saver_lib.SanityCheckFailed):
    sw.Save(None, 10)
  w[0, 0] = 0.0
  with pytest.raises(saver_lib.SanityCheckFailed):
    sw.Save(None, 500)                                       # global_step out of range
  # restoring under other names: {checkpoint name: tensor}
  w2 = torch.zeros(3, 2)
  sw2 = checkpointer.SaverWrapper(str(tmp_path), _TrainParams(),
                                  variables_to_restore_dict={'m/w': w2})
  sw2.Restore(None, saver_lib.LatestCheckpoint(str(tmp_path)))
  torch.testing.assert_close(w2[1:], want_w[1:])


def test_np_arrays_bundle_roundtrip(tmp_path):
  nmap = NestedMap(a=np.arange(6, dtype=np.float32).reshape(2, 3),
                   b=NestedMap(c=np.asarray([1, 2, 3], np.int64)))
  prefix = str(tmp_path / 'arrays')
  saver_lib.WriteNpArrays(prefix, nmap)
  got = saver_lib.ReadNpArrays(prefix, NestedMap(a=np.float32, b=NestedMap(c=np.int32)))
  np.testing.assert_array_equal(got.a, nmap.a)
  assert got.b.c.dtype == np.int32 and got.b.c.tolist() == [1, 2, 3]


def test_eager_checkpointer_flavours(tmp_path):
  from lingvo_b200 import model_registry
  import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa
  cfg = model_registry.GetParams('lm.synthetic_packed_input.MoELm8ETiny', 'Train')
  cfg.task.fprop_dtype = torch.float32
  cfg.task.builder.fprop_dtype = torch.float32
  model = cfg.Instantiate()
  assert checkpointer.EagerCheckpointerV1 is checkpointer.Checkpointer
  v2 = checkpointer.EagerCheckpointerV2(str(tmp_path), model,
                                        experimental_enable_async_checkpoint=True)
  assert v2.checkpoint_dir.endswith('ckpt_V2') and v2.async_checkpointing
  path = v2.Save(gsteps=3)
  v2.Sync()
  assert os.path.dirname(path).endswith('ckpt_V2') and os.path.exists(path + '.index')
  v1 = checkpointer.EagerCheckpointerV1(str(tmp_path), model)
  p1 = v1.Save(gsteps=4)
  assert os.path.dirname(p1) == str(tmp_path)
  first = model.vars.Flatten()[0]
  want = first.detach().clone()
  with torch.no_grad():
    first.zero_()
  assert checkpointer.EagerCheckpointerV2(str(tmp_path), model).Restore().endswith('00000003')
  torch.testing.assert_close(first.detach(), want)
