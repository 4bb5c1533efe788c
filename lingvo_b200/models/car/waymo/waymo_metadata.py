"""Waymo Open Dataset evaluation metadata (ref
`lingvo/tasks/car/waymo/waymo_metadata.py`)."""

from lingvo_b200.models.car import evaluation_metadata


class WaymoMetadata(evaluation_metadata.EvaluationMetadata):

  SPEC = dict(
      class_names=['Unknown', 'Vehicle', 'Pedestrian', 'Sign', 'Cyclist'],
      difficulty_levels={'UNKNOWN': 0, 'LEVEL_1': 1, 'LEVEL_2': 2},
      iou_thresholds={'Vehicle': 0.7, 'Pedestrian': 0.5, 'Cyclist': 0.5},
      eval_classes=['Vehicle', 'Pedestrian', 'Cyclist'],
      ignore_neighbors={},
      pr_points=101, max_distance=80.0, distance_bin_width=5.0, max_num_points=30000.0,
      num_points_bins=20, rotation_bins=10, calibration_bins=15,
      min_height_2d={'UNKNOWN': 0, 'LEVEL_1': 0, 'LEVEL_2': 0})

  def __init__(self):
    super().__init__('waymo')
