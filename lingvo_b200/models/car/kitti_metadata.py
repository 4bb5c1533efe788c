"""KITTI evaluation metadata (ref `lingvo/tasks/car/kitti_metadata.py`)."""

from lingvo_b200.models.car import evaluation_metadata


class KITTIMetadata(evaluation_metadata.EvaluationMetadata):

  SPEC = dict(
      class_names=['Background', 'Car', 'Van', 'Truck', 'Pedestrian', 'Person_sitting',
                   'Cyclist', 'Tram', 'Misc', 'DontCare'],
      difficulty_levels={'hard': 1, 'moderate': 2, 'easy': 3},
      iou_thresholds={'Car': 0.7, 'Pedestrian': 0.5, 'Cyclist': 0.5},
      # a detection matching a neighbour-class box is neither TP nor FP
      ignore_neighbors={'Car': ['Van'], 'Pedestrian': ['Person_sitting']},
      pr_points=41, max_distance=80.0, distance_bin_width=10.0, max_num_points=3000.0,
      num_points_bins=20, rotation_bins=10, calibration_bins=15,
      min_height_2d={'hard': 25, 'moderate': 25, 'easy': 40})

  def __init__(self):
    super().__init__('kitti')
