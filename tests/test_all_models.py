"""Every registered model's params build for every dataset it defines, round-trip through
text, and (for the small ones) instantiate (ref `lingvo/models_test.py`)."""

import pytest
import torch

from lingvo_b200 import model_imports
from lingvo_b200 import model_registry
from lingvo_b200 import datasets

model_imports.ImportAllParams()
_ALL = sorted(model_registry.GetAllRegisteredClasses().keys())
_INSTANTIATE = {
    'image.mnist.LeNet5', 'lm.synthetic_packed_input.MoELm8ETiny',
    'lm.one_billion_wds.WordLevelOneBwdsSimpleSampledSoftmaxTiny',
    'car.kitti.PointPillarsCarTiny', 'mt.wmt14_en_de.WmtEnDeTransformerSmall',
}


def test_registry_is_populated():
  tasks = {n.split('.')[0] for n in _ALL}
  assert {'image', 'lm', 'mt', 'asr', 'car', 'punctuator'} <= tasks, tasks
  assert len(_ALL) >= 25, _ALL


@pytest.mark.parametrize('name', _ALL)
def test_model_params(name):
  cls = model_registry.GetClass(name)
  ds = datasets.GetDatasets(cls) or ['Train']
  for d in ds:
    try:
      p = model_registry.GetParams(name, d)
    except (datasets.DatasetFunctionError, NotImplementedError):
      continue
    assert p.cls is not None
    text = p.ToText()
    assert text
    assert p.input is not None or p.task.input is not None
  p = model_registry.GetParams(name, 'Train')
  if name in _INSTANTIATE:
    p.cluster.worker.gpus_per_replica = 0
    p.cluster.mode = 'sync'
    try:
      from lingvo_b200.core import cluster_factory
      with cluster_factory.Cluster(p.cluster):
        model = p.Instantiate()
    except (FileNotFoundError, RuntimeError) as e:
      if 'No such file' in str(e) or 'cannot open' in str(e) or 'no files match' in str(e):
        pytest.skip('dataset / vocab file not present: %s' % e)
      raise
    assert len(model.tasks) >= 1
    n = sum(v.numel() for v in model.vars.Flatten())
    assert n > 0
