"""KITTI-style PointPillars configs (ref `lingvo/tasks/car/params/kitti.py`)."""

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.core import optimizer
from lingvo_b200.core import schedule
from lingvo_b200.models.car import input_generator
from lingvo_b200.models.car import pillars


@model_registry.RegisterSingleTaskModel
class PointPillarsCar(base_model_params.SingleTaskModelParams):
  """PointPillars on a 432×496 BEV grid, 12k pillars × 100 points (ref PillarsModelV1)."""

  GRID_X = (0.0, 69.12, 432)
  GRID_Y = (-39.68, 39.68, 496)
  MAX_PILLARS = 12000
  POINTS_PER_PILLAR = 100

  def _Input(self, seed):
    p = input_generator.SyntheticPillarsInput.Params()
    p.pillars.Set(grid_x=self.GRID_X, grid_y=self.GRID_Y, max_pillars=self.MAX_PILLARS,
                  points_per_pillar=self.POINTS_PER_PILLAR)
    p.seed = seed
    p.batch_size = 2
    return p

  def Train(self):
    return self._Input(0)

  def Dev(self):
    return self._Input(1)

  def Task(self):
    p = pillars.ModelV1.Params()
    p.featurizer.Set(grid_x=self.GRID_X, grid_y=self.GRID_Y)
    p.train.optimizer = optimizer.Adam.Params()
    p.train.learning_rate = 2e-4
    p.train.lr_schedule = schedule.Constant.Params()
    return p


@model_registry.RegisterSingleTaskModel
class PointPillarsCarTiny(PointPillarsCar):
  GRID_X = (-16.0, 16.0, 16)
  GRID_Y = (-16.0, 16.0, 16)
  MAX_PILLARS = 128
  POINTS_PER_PILLAR = 16

  def Task(self):
    p = super().Task()
    p.featurizer.num_features = 16
    p.backbone.Set(block_dims=(16, 32), block_layers=(1, 1), upsample_dim=16)
    p.train.learning_rate = 2e-3
    return p
