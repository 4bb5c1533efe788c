"""Waymo Open Dataset experiments: StarNet (ref `lingvo/tasks/car/params/waymo.py`)."""

import math
import os

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.core import optimizer
from lingvo_b200.models.car import input_preprocessors
from lingvo_b200.models.car import lr_util
from lingvo_b200.models.car import starnet
from lingvo_b200.models.car.waymo import waymo_decoder
from lingvo_b200.models.car.waymo import waymo_metadata
from lingvo_b200.models.car.waymo import waymo_open_input_generator

WAYMO_DIR = os.environ.get('LINGVO_B200_WAYMO', '/tmp/waymo/')


def _Spec(params, pattern, n):
  p = params.Copy()
  p.file_pattern = 'tfrecord:' + os.path.join(WAYMO_DIR, pattern)
  p.num_samples = n
  return p


def WaymoTrainSpec(params):
  return _Spec(params, 'train.tfr-*-of-01000', 158361)


def WaymoMiniTrainSpec(params):
  """Training shards used for decoding."""
  return _Spec(params, 'train.tfr-000[0-2]?-of-01000', 4773)


def WaymoValSpec(params):
  """Validation (no run segment overlaps training)."""
  return _Spec(params, 'valid.tfr-*-of-01000', 40077)


def WaymoMinivalSpec(params):
  """10 % of validation."""
  return _Spec(params, 'valid.tfr-000??-of-01000', 4109)


def WaymoTestSpec(params):
  return WaymoValSpec(params)


def _DatasetClass(name, spec):
  base = waymo_open_input_generator.WaymoSparseLaser
  def Params(cls):
    return spec(super(klass, cls).Params())
  klass = type(name, (base,), {'__doc__': 'WaymoSparseLaser over %s.' % spec.__name__})
  klass.Params = classmethod(Params)
  return klass


WaymoSparseLaserTrain = _DatasetClass('WaymoSparseLaserTrain', WaymoTrainSpec)
WaymoSparseLaserValidation = _DatasetClass('WaymoSparseLaserValidation', WaymoValSpec)
WaymoSparseLaserMinival = _DatasetClass('WaymoSparseLaserMinival', WaymoMinivalSpec)
WaymoSparseLaserTest = _DatasetClass('WaymoSparseLaserTest', WaymoTestSpec)


def _FilterKeepLabels(params, label_names):
  """Only keep ground truth of the named classes (ref :119)."""
  names = waymo_metadata.WaymoMetadata().ClassNames()
  params.extractors.labels.filter_labels = [names.index(n) for n in label_names]


class StarNetBase(base_model_params.SingleTaskModelParams):
  """StarNet on Waymo (ref :129)."""

  RUN_LOCALLY = False
  NUM_ANCHOR_BBOX_OFFSETS = 25
  NUM_ANCHOR_BBOX_ROTATIONS = 4
  NUM_ANCHOR_BBOX_DIMENSIONS = 1
  FOREGROUND_ASSIGNMENT_THRESHOLD = 0.6
  BACKGROUND_ASSIGNMENT_THRESHOLD = 0.45
  INCLUDED_CLASSES = ['Vehicle', 'Pedestrian', 'Cyclist']
  NUM_CELL_CENTERS = 1024
  NUM_POINTS_PER_CELL = 128
  MAX_DISTANCE = 2.75
  TRAIN_BATCH = 2
  ANCHOR_DIMENSIONS = [(4.7, 2.1, 1.7)]
  ANCHOR_XY_OFFSETS = [-1.5, -0.75, 0.0, 0.75, 1.5]

  def _configure_input(self, p, split):   # pylint: disable=invalid-name
    ip = input_preprocessors
    _FilterKeepLabels(p, self.INCLUDED_CLASSES)
    pre = p.preprocessors
    pre.select_centers.num_cell_centers = self.NUM_CELL_CENTERS
    pre.gather_features.Set(num_points_per_cell=self.NUM_POINTS_PER_CELL,
                            max_distance=self.MAX_DISTANCE)
    settings = ip.MakeAnchorBoxSettings(
        self.ANCHOR_DIMENSIONS,
        [i * math.pi / self.NUM_ANCHOR_BBOX_ROTATIONS for i in range(self.NUM_ANCHOR_BBOX_ROTATIONS)],
        self.ANCHOR_XY_OFFSETS, self.ANCHOR_XY_OFFSETS, [0.0])
    settings.Update(pre.tile_anchors)
    pre.assign_anchors.Set(
        foreground_assignment_threshold=self.FOREGROUND_ASSIGNMENT_THRESHOLD,
        background_assignment_threshold=self.BACKGROUND_ASSIGNMENT_THRESHOLD)
    if split == 'Train':
      pre.Define('random_flip', ip.RandomFlipY.Params(), '')
      pre.Define('world_rot', ip.RandomWorldRotationAboutZAxis.Params().Set(
          max_rotation=math.pi / 4), '')
      pre.Define('world_scale', ip.WorldScaling.Params().Set(scaling=(0.95, 1.05)), '')
      order = list(p.preprocessors_order)
      at = order.index('select_centers')
      p.preprocessors_order = order[:at] + ['random_flip', 'world_rot', 'world_scale'] + order[at:]
      p.batch_size = self.TRAIN_BATCH
    else:
      p.batch_size = 4
      p.file_parallelism = 1
    p.file_buffer_size = 32
    p.num_batcher_threads = 8
    return p

  def Train(self):
    return self._configure_input(WaymoSparseLaserTrain.Params(), 'Train')

  def Minitrain(self):
    p = WaymoMiniTrainSpec(waymo_open_input_generator.WaymoSparseLaser.Params())
    return self._configure_input(p, 'Minitrain')

  def Test(self):
    return self._configure_input(WaymoSparseLaserTest.Params(), 'Test')

  def Dev(self):
    return self._configure_input(WaymoSparseLaserValidation.Params(), 'Dev')

  def Minidev(self):
    return self._configure_input(WaymoSparseLaserMinival.Params(), 'Minidev')

  def Task(self):
    num_classes = 1 + max(waymo_metadata.WaymoMetadata().ClassNames().index(n)
                          for n in self.INCLUDED_CLASSES)
    p = starnet.ModelV2.Params(
        num_classes, num_anchor_bboxes_offsets=self.NUM_ANCHOR_BBOX_OFFSETS,
        num_anchor_bboxes_rotations=self.NUM_ANCHOR_BBOX_ROTATIONS,
        num_anchor_bboxes_dimensions=self.NUM_ANCHOR_BBOX_DIMENSIONS, num_laser_features=3)
    p.name = 'sparse_detector'
    p.output_decoder = waymo_decoder.WaymoOpenDatasetDecoder.Params()
    tp = p.train
    tp.optimizer = optimizer.Adam.Params()
    tp.clip_gradient_norm_to_value = 5
    tp.learning_rate = 1e-3
    lr_util.SetExponentialLR(train_p=tp, train_input_p=self.Train(), exp_start_epoch=5,
                             total_epoch=75)
    p.dimension_loss_weight = 0.3
    p.location_loss_weight = 3.0
    p.loss_weight_classification = 1.0
    p.loss_weight_localization = 3.0
    p.rotation_loss_weight = 0.3
    p.use_oriented_per_class_nms = True
    p.max_nms_boxes = 512
    p.nms_iou_threshold = [0.0] + [0.2] * (num_classes - 1)
    p.nms_score_threshold = [1.0] + [0.03] * (num_classes - 1)
    return p


@model_registry.RegisterSingleTaskModel
class StarNetVehicle(StarNetBase):
  """Vehicles only (ref :354)."""
  INCLUDED_CLASSES = ['Vehicle']
  NUM_CELL_CENTERS = 1024

  def Task(self):
    p = super().Task()
    p.nms_iou_threshold = [0.0, 0.03] + [0.0] * (p.num_classes - 2)
    p.nms_score_threshold = [1.0, 0.06] + [1.0] * (p.num_classes - 2)
    return p


@model_registry.RegisterSingleTaskModel
class StarNetPed(StarNetBase):
  """Pedestrians only (ref :398)."""
  INCLUDED_CLASSES = ['Pedestrian']
  FOREGROUND_ASSIGNMENT_THRESHOLD = 0.5
  BACKGROUND_ASSIGNMENT_THRESHOLD = 0.4
  NUM_ANCHOR_BBOX_OFFSETS = 9
  MAX_DISTANCE = 2.0
  ANCHOR_DIMENSIONS = [(0.9, 0.9, 1.75)]
  ANCHOR_XY_OFFSETS = [-0.55, 0.0, 0.55]

  def Task(self):
    p = super().Task()
    p.nms_iou_threshold = [0.0, 0.0, 0.46] + [0.0] * (p.num_classes - 3)
    p.nms_score_threshold = [1.0, 1.0, 0.01] + [1.0] * (p.num_classes - 3)
    return p


@model_registry.RegisterSingleTaskModel
class StarNetPedFused(StarNetPed):
  """Pedestrian model with camera-aware input (adds the image extractor and the
  cell-centre → camera association) (ref :444)."""

  def _configure_input(self, p, split):   # pylint: disable=invalid-name
    p = super()._configure_input(p, split)
    p.extractors.Define('images', waymo_open_input_generator.WaymoImageExtractor.Params().Set(
        camera_names=['FRONT']), '')
    p.preprocessors.Define('best_camera', waymo_open_input_generator.CellCenterToBestCamera
                           .Params().Set(camera_names=['FRONT']), '')
    order = list(p.preprocessors_order)
    order.insert(order.index('gather_features'), 'best_camera')
    p.preprocessors_order = order
    return p
