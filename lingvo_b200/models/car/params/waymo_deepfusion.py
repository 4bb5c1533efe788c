"""Waymo anchor-free pillars → CenterPoint → DeepFusion lineage (ref
`lingvo/tasks/car/params/waymo_deepfusion.py`; reported pedestrian L1 mAP 65.9 → 69.5 → 71
→ 80.4 → 81.5, vehicle 65.2 / 76.5).

Input: Waymo lasers → range crop → (augmentation) → pillar grid → point assignment on the
grid centres. Model: `pillars_anchor_free.ModelV1` with progressively more of: dynamic
voxelisation, augmentation, centerness, heat-map NMS + stronger backbone ("CenterPoint
improved"), Laplace-KL uncertainty, camera fusion.
"""

import math

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.core import hyperparams
from lingvo_b200.core import optimizer
from lingvo_b200.models.car import deep_fusion
from lingvo_b200.models.car import input_preprocessors
from lingvo_b200.models.car import lr_util
from lingvo_b200.models.car import pillars
from lingvo_b200.models.car import pillars_anchor_free
from lingvo_b200.models.car.params import waymo as waymo_params
from lingvo_b200.models.car.waymo import waymo_decoder
from lingvo_b200.models.car.waymo import waymo_metadata
from lingvo_b200.models.car.waymo import waymo_open_input_generator


class AnchorFreePillarsModelV1Base(base_model_params.SingleTaskModelParams):
  """ref :145."""

  INCLUDED_CLASSES = ['Vehicle', 'Pedestrian', 'Cyclist']
  GRID_X_RANGE = (-76.8, 76.8)
  GRID_Y_RANGE = (-76.8, 76.8)
  GRID_Z_RANGE = (-3.0, 3.0)
  GRID = (512, 512, 1)
  NUM_PILLARS = 32000
  POINTS_PER_PILLAR = 20
  TRAIN_BATCH = 2
  AUGMENT = False
  DYNAMIC_VOXELIZATION = False
  CENTERNESS_WEIGHT = 0.0
  NMS_DECODER = pillars_anchor_free.NMSDecoderType.NMS_DECODER
  LOCATION_LOSS = pillars_anchor_free.HuberLoss
  TOTAL_EPOCH = 75
  CAMERA_FUSION = False

  def _Input(self, spec, train):
    ip = input_preprocessors
    p = spec(waymo_open_input_generator.WaymoSparseLaser.Params())
    waymo_params._FilterKeepLabels(p, self.INCLUDED_CLASSES)   # pylint: disable=protected-access
    pre = hyperparams.Params()
    pre.Define('filter_nlz_points', waymo_open_input_generator.FilterNLZPoints.Params(), '')
    pre.Define('viz_copy', ip.CreateDecoderCopy.Params().Set(
        pad_lasers=ip.PadLaserFeatures.Params().Set(max_num_points=240000)), '')
    order = ['filter_nlz_points', 'viz_copy']
    if train and self.AUGMENT:
      pre.Define('random_flip', ip.RandomFlipY.Params(), '')
      pre.Define('world_rot', ip.RandomWorldRotationAboutZAxis.Params().Set(
          max_rotation=math.pi / 4), '')
      pre.Define('world_scale', ip.WorldScaling.Params().Set(scaling=(0.95, 1.05)), '')
      pre.Define('translate', ip.GlobalTranslateNoise.Params().Set(noise_std=[0.2, 0.2, 0.2]), '')
      order += ['random_flip', 'world_rot', 'world_scale', 'translate']
    pre.Define('keep_xyz_range', ip.DropLaserPointsOutOfRange.Params().Set(
        keep_x_range=self.GRID_X_RANGE, keep_y_range=self.GRID_Y_RANGE,
        keep_z_range=self.GRID_Z_RANGE), '')
    pre.Define('drop_boxes', ip.DropBoxesOutOfRange.Params().Set(
        keep_x_range=self.GRID_X_RANGE, keep_y_range=self.GRID_Y_RANGE), '')
    order += ['keep_xyz_range', 'drop_boxes']
    if not self.DYNAMIC_VOXELIZATION:
      pre.Define('points_to_grid', ip.PointsToGrid.Params().Set(
          num_points_per_cell=self.POINTS_PER_PILLAR, grid_size=self.GRID,
          grid_range_x=self.GRID_X_RANGE, grid_range_y=self.GRID_Y_RANGE,
          grid_range_z=self.GRID_Z_RANGE), '')
      pre.Define('grid_to_pillars', ip.GridToPillars.Params().Set(
          num_pillars=self.NUM_PILLARS, num_points_per_pillar=self.POINTS_PER_PILLAR), '')
      pre.Define('pillar_centers', ip.PerPillarPointCloudCenters.Params().Set(
          grid_size=self.GRID, grid_range_x=self.GRID_X_RANGE, grid_range_y=self.GRID_Y_RANGE,
          grid_range_z=self.GRID_Z_RANGE), '')
      order += ['points_to_grid', 'grid_to_pillars', 'pillar_centers']
    pre.Define('anchor_centers', ip.GridAnchorCenters.Params().Set(
        grid_size=self.GRID, grid_range_x=self.GRID_X_RANGE, grid_range_y=self.GRID_Y_RANGE,
        grid_range_z=(0.0, 0.0)), '')
    n_cls = 1 + max(waymo_metadata.WaymoMetadata().ClassNames().index(n)
                    for n in self.INCLUDED_CLASSES)
    pre.Define('assign_points', ip.PointAssignment.Params().Set(num_classes=n_cls), '')
    pre.Define('pad_lasers', ip.PadLaserFeatures.Params().Set(max_num_points=240000), '')
    order += ['anchor_centers', 'assign_points', 'pad_lasers']
    p.preprocessors = pre
    p.preprocessors_order = order
    if self.CAMERA_FUSION:
      p.extractors.Define('images', waymo_open_input_generator.WaymoImageExtractor.Params()
                          .Set(camera_names=['FRONT']), '')
    p.batch_size = self.TRAIN_BATCH if train else 4
    return p

  def Train(self):
    return self._Input(waymo_params.WaymoTrainSpec, True)

  def Dev(self):
    return self._Input(waymo_params.WaymoValSpec, False)

  def Minidev(self):
    return self._Input(waymo_params.WaymoMinivalSpec, False)

  def Test(self):
    return self._Input(waymo_params.WaymoTestSpec, False)

  def Task(self):
    n_cls = 1 + max(waymo_metadata.WaymoMetadata().ClassNames().index(n)
                    for n in self.INCLUDED_CLASSES)
    p = pillars_anchor_free.ModelV1.Params(grid_size_z=self.GRID[2], num_classes=n_cls,
                                           num_laser_features=3)
    p.name = 'anchor_free_pillars'
    if self.DYNAMIC_VOXELIZATION:
      p.input_featurizer = pillars.DynamicVoxelizationFeaturizer.Params(3, 64)
      p.input_featurizer.dynamic_voxelization.Set(
          grid_size=self.GRID, grid_range_x=self.GRID_X_RANGE, grid_range_y=self.GRID_Y_RANGE,
          grid_range_z=self.GRID_Z_RANGE)
    else:
      p.input_featurizer.grid_size = self.GRID
    if self.CAMERA_FUSION:
      p.input_featurizer = deep_fusion.MultiModalFeaturizer.Params(3, 64)
      p.input_featurizer.lidar_featurizer.grid_size = self.GRID
    p.centerness_loss_weight = self.CENTERNESS_WEIGHT
    p.nms_decoder_type = self.NMS_DECODER
    p.location_loss = self.LOCATION_LOSS.Params()
    p.output_decoder = waymo_decoder.WaymoOpenDatasetDecoder.Params()
    p.use_oriented_per_class_nms = True
    p.max_nms_boxes = 512
    p.nms_iou_threshold = [0.0] + [0.2] * (n_cls - 1)
    p.nms_score_threshold = [1.0] + [0.05] * (n_cls - 1)
    tp = p.train
    tp.optimizer = optimizer.Adam.Params()
    tp.clip_gradient_norm_to_value = 5
    tp.learning_rate = 1e-3
    lr_util.SetCosineLR(train_p=tp, train_input_p=self.Train(), total_epoch=self.TOTAL_EPOCH,
                        warmup_epoch=1)
    return p


@model_registry.RegisterSingleTaskModel
class AnchorFreePillarsModelV1Ped(AnchorFreePillarsModelV1Base):
  """Pedestrian baseline: 65.9 L1 mAP in the reference (ref :340)."""
  INCLUDED_CLASSES = ['Pedestrian']
  GRID_X_RANGE = (-74.88, 74.88)
  GRID_Y_RANGE = (-74.88, 74.88)
  GRID = (468, 468, 1)


@model_registry.RegisterSingleTaskModel
class AnchorFreePillarsModelV1PedDV(AnchorFreePillarsModelV1Ped):
  """+ dynamic voxelisation (69.5; ref :417)."""
  DYNAMIC_VOXELIZATION = True


@model_registry.RegisterSingleTaskModel
class AnchorFreePillarsModelV1PedAug(AnchorFreePillarsModelV1PedDV):
  """+ world augmentation (71; ref :440)."""
  AUGMENT = True


@model_registry.RegisterSingleTaskModel
class AnchorFreePillarsModelV1PedCenterNess(AnchorFreePillarsModelV1PedAug):
  """+ centerness head (ref :472)."""
  CENTERNESS_WEIGHT = 1.0


@model_registry.RegisterSingleTaskModel
class AnchorFreePillarsModelV1PedCenterNessRelated(AnchorFreePillarsModelV1PedCenterNess):
  """Centerness with a tighter label range (ref :493)."""

  def _Input(self, spec, train):
    p = super()._Input(spec, train)
    p.preprocessors.assign_points.centerness_range = (0.5, 1.0)
    return p


@model_registry.RegisterSingleTaskModel
class AnchorFreePillarsModelV1VehicleCenterNess(AnchorFreePillarsModelV1PedCenterNess):
  """Vehicle version (65.2; ref :514)."""
  INCLUDED_CLASSES = ['Vehicle']
  GRID_X_RANGE = (-76.8, 76.8)
  GRID_Y_RANGE = (-76.8, 76.8)
  GRID = (512, 512, 1)


@model_registry.RegisterSingleTaskModel
class CenterPointImprovedVehicle(AnchorFreePillarsModelV1VehicleCenterNess):
  """CenterPoint-style decode (max-pool heat-map NMS), wider backbone, longer schedule
  (76.5; ref :555)."""
  NMS_DECODER = pillars_anchor_free.NMSDecoderType.HEATMAP_NMS_DECODER
  TOTAL_EPOCH = 100

  def Task(self):
    p = super().Task()
    b = pillars.Builder(pillars.Builder.Params())
    p.backbone = b.Backbone(64, dims=(64, 128, 256), repeats=(3, 5, 5),
                            up_dims=pillars_anchor_free.AnchorFreePillarsBase.NUM_OUTPUT_CHANNELS,
                            first_stride=1)
    p.heatmap_nms_kernel_size = [1, 3, 3, 1]
    p.heatmap_nms_score_threshold = 0.1
    p.corner_loss_weight = 0.1
    return p


@model_registry.RegisterSingleTaskModel
class CenterPointImprovedPedestrian(CenterPointImprovedVehicle):
  """80.4 (ref :750)."""
  INCLUDED_CLASSES = ['Pedestrian']
  GRID_X_RANGE = (-74.88, 74.88)
  GRID_Y_RANGE = (-74.88, 74.88)
  GRID = (468, 468, 1)


@model_registry.RegisterSingleTaskModel
class UncertaintyCenterPointPed(CenterPointImprovedPedestrian):
  """+ Laplace-KL localisation loss with predicted uncertainty (ref :768)."""
  LOCATION_LOSS = pillars_anchor_free.LaplaceKL


@model_registry.RegisterSingleTaskModel
class DeepFusionCenterPointPed(UncertaintyCenterPointPed):
  """+ LearnableAlign camera fusion (81.5; ref :797)."""
  CAMERA_FUSION = True
  DYNAMIC_VOXELIZATION = False
