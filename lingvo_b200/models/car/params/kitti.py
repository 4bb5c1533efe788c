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


# --------------------------------------------------------------------------------------
# KITTI dataset specs + StarNet experiments (ref :40-620)
# --------------------------------------------------------------------------------------
import math  # pylint: disable=g-import-not-at-top,wrong-import-position
import os  # pylint: disable=wrong-import-position

from lingvo_b200.core import hyperparams  # pylint: disable=wrong-import-position
from lingvo_b200.models.car import input_preprocessors  # pylint: disable=wrong-import-position
from lingvo_b200.models.car import kitti_input_generator  # pylint: disable=wrong-import-position
from lingvo_b200.models.car import lr_util  # pylint: disable=wrong-import-position
from lingvo_b200.models.car import starnet  # pylint: disable=wrong-import-position

KITTI_DIR = os.environ.get('LINGVO_B200_KITTI', '/tmp/kitti/')


def _Spec(params, pattern, n):
  p = params.Copy()
  p.file_pattern = 'tfrecord:' + os.path.join(KITTI_DIR, pattern)
  p.num_samples = n
  return p


def KITTITrainSpec(params):
  return _Spec(params, 'kitti_object_3dop_train.tfrecord-*-of-00100', 3712)


def KITTIValSpec(params):
  return _Spec(params, 'kitti_object_3dop_val.tfrecord-*-of-00100', 3769)


def KITTITestSpec(params):
  return _Spec(params, 'kitti_object_test.tfrecord-*-of-00100', 7518)


def _DatasetClass(name, base, spec):
  """Creates `class <name>(base)` whose Params() is the base's with `spec` applied."""
  def Params(cls):
    return spec(super(klass, cls).Params())
  klass = type(name, (base,), {'__doc__': '%s over the %s split.' % (base.__name__, spec.__name__)})
  klass.Params = classmethod(Params)
  return klass


KITTITrain = _DatasetClass('KITTITrain', kitti_input_generator.KITTILaser, KITTITrainSpec)
KITTIValidation = _DatasetClass('KITTIValidation', kitti_input_generator.KITTILaser, KITTIValSpec)
KITTITest = _DatasetClass('KITTITest', kitti_input_generator.KITTILaser, KITTITestSpec)
KITTIGridTrain = _DatasetClass('KITTIGridTrain', kitti_input_generator.KITTIGrid, KITTITrainSpec)
KITTIGridValidation = _DatasetClass('KITTIGridValidation', kitti_input_generator.KITTIGrid,
                                    KITTIValSpec)
KITTIGridTest = _DatasetClass('KITTIGridTest', kitti_input_generator.KITTIGrid, KITTITestSpec)
KITTISparseLaserTrain = _DatasetClass('KITTISparseLaserTrain',
                                      kitti_input_generator.KITTISparseLaser, KITTITrainSpec)
KITTISparseLaserValidation = _DatasetClass('KITTISparseLaserValidation',
                                           kitti_input_generator.KITTISparseLaser, KITTIValSpec)
KITTISparseLaserTest = _DatasetClass('KITTISparseLaserTest',
                                     kitti_input_generator.KITTISparseLaser, KITTITestSpec)


def _MaybeRemove(values, key):
  if key in values:
    values.remove(key)


def AddLaserAndCamera(params):
  """Makes every laser / label preprocessor also see the camera calibration (ref :151)."""
  params.extractors.images.decode_image = True
  return params


class StarNetCarsBase(base_model_params.SingleTaskModelParams):
  """StarNet on KITTI cars (ref :185)."""

  RUN_LOCALLY = False
  NUM_ANCHOR_BBOX_OFFSETS = 25
  NUM_ANCHOR_BBOX_ROTATIONS = 4
  NUM_ANCHOR_BBOX_DIMENSIONS = 1
  FOREGROUND_ASSIGNMENT_THRESHOLD = 0.6
  BACKGROUND_ASSIGNMENT_THRESHOLD = 0.45
  INCLUDED_CLASSES = ['Car']
  NUM_CELL_CENTERS = 768
  NUM_POINTS_PER_CELL = 32
  MAX_DISTANCE = 3.0
  TRAIN_BATCH = 2

  class AnchorBoxSettings(input_preprocessors.SparseCarV1AnchorBoxSettings):
    ROTATIONS = [0, math.pi / 2, 3 * math.pi / 4, math.pi / 4]
    CENTER_X_OFFSETS = [-1.5, -0.75, 0.0, 0.75, 1.5]
    CENTER_Y_OFFSETS = [-1.5, -0.75, 0.0, 0.75, 1.5]

  def _configure_generic_input(self, p):   # pylint: disable=invalid-name
    names = kitti_input_generator.KITTI_CLASS_NAMES
    p.extractors.labels.filter_labels = [names.index(n) for n in self.INCLUDED_CLASSES]
    pre = p.preprocessors
    pre.select_centers.num_cell_centers = self.NUM_CELL_CENTERS
    pre.gather_features.Set(num_points_per_cell=self.NUM_POINTS_PER_CELL,
                            max_distance=self.MAX_DISTANCE)
    self.AnchorBoxSettings.Update(pre.tile_anchors)
    pre.assign_anchors.Set(
        foreground_assignment_threshold=self.FOREGROUND_ASSIGNMENT_THRESHOLD,
        background_assignment_threshold=self.BACKGROUND_ASSIGNMENT_THRESHOLD)
    p.file_buffer_size = 32
    p.file_parallelism = 8
    p.num_batcher_threads = 8
    return p

  def _configure_trainer_input(self, p):   # pylint: disable=invalid-name
    """World augmentation in front of the sampling stage (ref :265)."""
    ip = input_preprocessors
    pre = p.preprocessors
    pre.Define('random_flip', ip.RandomFlipY.Params(), '')
    pre.Define('world_rot', ip.RandomWorldRotationAboutZAxis.Params().Set(
        max_rotation=math.pi / 4), '')
    pre.Define('world_scale', ip.WorldScaling.Params().Set(scaling=(0.95, 1.05)), '')
    pre.Define('bbox_aug', ip.RandomBBoxTransform.Params().Set(
        max_rotation=math.pi / 20, noise_std=[0.5, 0.5, 0.0]), '')
    pre.Define('frustum_dropout', ip.FrustumDropout.Params().Set(
        theta_width=0.03, phi_width=0.0, keep_prob=0.0), '')
    order = list(p.preprocessors_order)
    at = order.index('select_centers')
    p.preprocessors_order = order[:at] + ['random_flip', 'world_rot', 'world_scale', 'bbox_aug',
                                          'frustum_dropout'] + order[at:]
    p.batch_size = self.TRAIN_BATCH
    return p

  def _configure_decoder_input(self, p):   # pylint: disable=invalid-name
    p.batch_size = 4
    p.file_parallelism = 1
    p.file_random_seed = 9
    return p

  def _configure_evaler_input(self, p):   # pylint: disable=invalid-name
    p.batch_size = 4
    return p

  def Train(self):
    p = KITTISparseLaserTrain.Params()
    return self._configure_trainer_input(self._configure_generic_input(p))

  def Test(self):
    p = KITTISparseLaserTest.Params()
    return self._configure_decoder_input(self._configure_generic_input(p))

  def Dev(self):
    p = KITTISparseLaserValidation.Params()
    return self._configure_decoder_input(self._configure_generic_input(p))

  def Task(self):
    num_classes = 1 + len(self.INCLUDED_CLASSES)
    p = starnet.ModelV2.Params(
        num_classes, num_anchor_bboxes_offsets=self.NUM_ANCHOR_BBOX_OFFSETS,
        num_anchor_bboxes_rotations=self.NUM_ANCHOR_BBOX_ROTATIONS,
        num_anchor_bboxes_dimensions=self.NUM_ANCHOR_BBOX_DIMENSIONS)
    p.name = 'sparse_detector'
    tp = p.train
    tp.optimizer = optimizer.Adam.Params()
    tp.clip_gradient_norm_to_value = 5
    tp.learning_rate = 1e-3
    lr_util.SetExponentialLR(train_p=tp, train_input_p=self.Train(), exp_start_epoch=150,
                             total_epoch=650)
    p.dimension_loss_weight = 0.3
    p.location_loss_weight = 3.0
    p.loss_weight_classification = 1.0
    p.loss_weight_localization = 3.0
    p.rotation_loss_weight = 0.3
    p.nms_iou_threshold = [0.0, 0.1][:num_classes] + [0.1] * max(0, num_classes - 2)
    p.nms_score_threshold = [1.0] + [0.05] * (num_classes - 1)
    p.use_oriented_per_class_nms = True
    p.max_nms_boxes = 512
    p.output_decoder.filter_predictions_outside_frustum = True
    p.output_decoder.truncation_threshold = 0.0
    return p


@model_registry.RegisterSingleTaskModel
class StarNetCarModel0701(StarNetCarsBase):
  """The released car model: 512 centres at decode time, GIN featurizer (ref :402)."""

  NUM_CELL_CENTERS = 512
  NUM_POINTS_PER_CELL = 64

  def _configure_generic_input(self, p):   # pylint: disable=invalid-name
    p = super()._configure_generic_input(p)
    p.preprocessors.keep_xyz_range.Set(keep_x_range=(0.0, 70.4), keep_y_range=(-40.0, 40.0),
                                       keep_z_range=(-3.0, 1.0))
    return p

  def Task(self):
    p = super().Task()
    p.num_attention_layers = 0
    return p


@model_registry.RegisterSingleTaskModel
class StarNetPedCycModel0704(StarNetCarsBase):
  """Pedestrians + cyclists (ref :499)."""

  INCLUDED_CLASSES = ['Pedestrian', 'Cyclist']
  FOREGROUND_ASSIGNMENT_THRESHOLD = 0.48
  BACKGROUND_ASSIGNMENT_THRESHOLD = 0.28
  NUM_ANCHOR_BBOX_OFFSETS = 9
  NUM_ANCHOR_BBOX_ROTATIONS = 2
  NUM_ANCHOR_BBOX_DIMENSIONS = 2
  MAX_DISTANCE = 2.0

  class AnchorBoxSettings(input_preprocessors.SparseCarV1AnchorBoxSettings):
    DIMENSION_PRIORS = [(0.6, 0.8, 1.73), (0.6, 1.76, 1.73)]
    ROTATIONS = [0, math.pi / 2]
    CENTER_X_OFFSETS = [-0.6, 0.0, 0.6]
    CENTER_Y_OFFSETS = [-0.6, 0.0, 0.6]

  def Task(self):
    p = super().Task()
    p.nms_iou_threshold = [0.0, 0.46, 0.46]
    p.nms_score_threshold = [1.0, 0.01, 0.01]
    return p


@model_registry.RegisterSingleTaskModel
class StarNetCarTiny(StarNetCarsBase):
  """Unit-test sized StarNet."""
  NUM_CELL_CENTERS = 16
  NUM_POINTS_PER_CELL = 8
  NUM_ANCHOR_BBOX_OFFSETS = 4
  NUM_ANCHOR_BBOX_ROTATIONS = 3

  class AnchorBoxSettings(input_preprocessors.SparseCarV1AnchorBoxSettings):
    pass

  def _configure_generic_input(self, p):   # pylint: disable=invalid-name
    p = super()._configure_generic_input(p)
    p.preprocessors.pad_lasers.max_num_points = 1024
    p.preprocessors.viz_copy.pad_lasers.max_num_points = 1024
    p.Set(file_parallelism=1, num_batcher_threads=1, file_buffer_size=4)
    return p

  def Task(self):
    p = super().Task()
    p.max_nms_boxes = 8
    return p



