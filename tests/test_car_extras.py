"""Car breadth: geometry, car_lib, native AP op, metrics, preprocessors, KITTI / Waymo
input pipelines, StarNet / PointPillars / anchor-free / DeepFusion models, params."""

import math
import os

import numpy as np
import pytest
import torch

from lingvo_b200 import model_registry
from lingvo_b200 import ops
from lingvo_b200.core import cluster_factory
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import ap_metric  # noqa: F401
from lingvo_b200.models.car import calibration_processing
from lingvo_b200.models.car import car_lib
from lingvo_b200.models.car import deep_fusion
from lingvo_b200.models.car import detection_3d_metrics as dm
from lingvo_b200.models.car import detection_decoder as dd
from lingvo_b200.models.car import geometry
from lingvo_b200.models.car import input_preprocessors as ip
from lingvo_b200.models.car import kitti_ap_metric
from lingvo_b200.models.car import kitti_input_generator as kig
from lingvo_b200.models.car import kitti_metadata
from lingvo_b200.models.car import ops as car_ops
from lingvo_b200.models.car import pillars
from lingvo_b200.models.car import pillars_anchor_free as paf
from lingvo_b200.models.car import pointnet
from lingvo_b200.models.car import starnet
from lingvo_b200.models.car import summary
from lingvo_b200.models.car import transform_util
from lingvo_b200.models.car.waymo import waymo_ap_metric
from lingvo_b200.models.car.waymo import waymo_metadata
from lingvo_b200.models.car.waymo import waymo_open_input_generator as wig
from lingvo_b200.utils import tf_example


# ------------------------------------------------------------------------------ fixtures
def WriteKitti(path, num=6):
  w = ops.host().TFRecordWriter(path)
  rng = np.random.RandomState(0)
  for i in range(num):
    n = 3
    xyz = np.concatenate([rng.uniform(5, 60, (500, 1)), rng.uniform(-30, 30, (500, 1)),
                          rng.uniform(-2, 0.5, (500, 1))], 1).astype(np.float32)
    loc = np.stack([rng.uniform(10, 50, n), rng.uniform(-20, 20, n), np.full(n, -1.0)],
                   1).astype(np.float32)
    xyz[:60] = np.repeat(loc, 20, 0) + rng.uniform(-0.5, 0.5, (60, 3))
    w.write(tf_example.MakeExample({
        'pointcloud/xyz': xyz.reshape(-1),
        'pointcloud/reflectance': rng.rand(500).astype(np.float32),
        'image/format': [b'png'], 'image/height': np.asarray([375]),
        'image/width': np.asarray([1242]), 'image/source_id': [b'%06d' % i],
        'transform/velo_to_image_plane': np.array(
            [[700, 0, 600, 0], [0, 700, 180, 0], [0, 0, 1, 0]], np.float32).reshape(-1),
        'transform/velo_to_camera': np.eye(4, dtype=np.float32).reshape(-1),
        'transform/camera_to_velo': np.eye(4, dtype=np.float32).reshape(-1),
        'object/image/bbox/xmin': np.array([10, 50, 90], np.float32),
        'object/image/bbox/xmax': np.array([60, 100, 140], np.float32),
        'object/image/bbox/ymin': np.array([100, 100, 100], np.float32),
        'object/image/bbox/ymax': np.array([150, 130, 160], np.float32),
        'object/label': [b'Car', b'Pedestrian', b'Car'],
        'object/has_3d_info': np.array([1, 1, 1]), 'object/occlusion': np.array([0, 1, 2]),
        'object/truncation': np.array([0.0, 0.2, 0.4], np.float32),
        'object/velo/bbox/xyz': loc.reshape(-1),
        'object/velo/bbox/dim_xyz': np.tile([3.9, 1.6, 1.56], n).astype(np.float32),
        'object/velo/bbox/phi': np.zeros(n, np.float32)}))
  w.close()


@pytest.fixture(scope='module')
def kitti_file(tmp_path_factory):
  path = str(tmp_path_factory.mktemp('kitti') / 'kitti.tfrecord')
  WriteKitti(path)
  return 'tfrecord:' + path


def _Small(p, path):
  p.Set(name='inp', file_pattern=path, batch_size=2, file_parallelism=1, file_buffer_size=4,
        num_batcher_threads=1)
  p.preprocessors.pad_lasers.max_num_points = 1024
  p.preprocessors.viz_copy.pad_lasers.max_num_points = 1024
  return p


def _Feats(P=400, L=6):
  g = torch.Generator().manual_seed(0)
  xyz = torch.rand(P, 3, generator=g) * torch.tensor([60., 60., 3.]) + torch.tensor([0., -30., -2.])
  boxes = torch.cat([torch.rand(L, 2, generator=g) * torch.tensor([50., 40.]) +
                     torch.tensor([5., -20.]), torch.full((L, 1), -1.0),
                     torch.tensor([[3.9, 1.6, 1.56]]).expand(L, 3),
                     torch.rand(L, 1, generator=g) * 3 - 1.5], 1)
  xyz[:20] = boxes[0, :3] + (torch.rand(20, 3, generator=g) - 0.5)
  return NestedMap(
      lasers=NestedMap(points_xyz=xyz, points_feature=torch.rand(P, 1, generator=g)),
      labels=NestedMap(labels=torch.tensor([1, 1, 2, 1, 0, 0]), bboxes_3d=boxes,
                       bboxes_3d_mask=torch.tensor([1., 1, 1, 1, 0, 0]),
                       difficulties=torch.tensor([1, 2, 3, 1, 0, 0])))


# --------------------------------------------------------------------------------- tests
def test_geometry_and_transforms():
  box = torch.tensor([[0., 0, 0, 2, 2, 2, 0]])
  pts = torch.tensor([[0., 0, 0], [0.9, 0.9, 0], [1.2, 0, 0]])
  assert geometry.IsWithinBBox3D(pts, box).squeeze(1).tolist() == [True, True, False]
  assert geometry.BBoxCorners(box).shape == (1, 8, 3)
  rot = geometry.BatchMakeRotationMatrix(torch.tensor(math.pi / 2))
  torch.testing.assert_close(rot @ torch.tensor([1., 0, 0]), torch.tensor([0., 1, 0]),
                             atol=1e-6, rtol=0)
  t = torch.eye(4); t[0, 3] = 5.0
  moved = geometry.TransformBBoxes3D(box.unsqueeze(0), t.unsqueeze(0))
  assert float(moved[0, 0, 0]) == 5.0
  xywh = torch.tensor([[3., 4, 2, 6]])
  torch.testing.assert_close(geometry.BBoxesToXYWH(geometry.XYWHToBBoxes(xywh)), xywh)
  sph = geometry.SphericalCoordinatesTransform(torch.tensor([[0., 0, 2]]))
  assert abs(float(sph[0, 0]) - 2) < 1e-6 and abs(float(sph[0, 1])) < 1e-6
  b2 = transform_util.Box2D(1, 2, 4, 2, math.pi / 2)
  np.testing.assert_allclose(b2.Extrema(), (0, 2, 0, 4), atol=1e-6)
  quad = torch.tensor([[0., 0], [2, 0], [2, 2], [0, 2]])
  inside = geometry.IsWithinBBox(torch.tensor([[1., 1], [3, 1]]), quad)
  assert inside.tolist() == [True, False]
  assert geometry.IsWithinBBox(torch.tensor([[1., 1]]), quad.flip(0)).tolist() == [True]


def test_car_lib():
  p = torch.randn(2, 50, 3)
  pad = torch.zeros(2, 50); pad[1, 30:] = 1
  s, cl = car_lib.FarthestPointSampler(p, pad, 8, random_seed=1)
  assert (s[1] < 30).all() and len(set(s[0].tolist())) == 8 and cl.shape == (2, 50)
  idx, ipad = car_lib.NeighborhoodIndices(p, p[:, :5], 4, pad > 0.5, max_distance=0.5)
  assert idx.shape == (2, 5, 4) and float(ipad[:, :, 0].sum()) == 0      # a point finds itself
  dv = car_lib.DynamicVoxelization(p, pad, [4, 4, 2], (-2, 2), (-2, 2), (-2, 2))
  st = car_lib.DynamicVoxelStatistics(p, dv)
  assert st.centroids.shape == (2, 50, 3) and dv.num_voxels == 32
  _, pf = car_lib.SegmentPool3D(p, torch.randn(2, 50, 6), s, cl, 'max')
  assert pf.shape == (2, 8, 6)
  d1 = car_lib.SquaredDistanceMatrix(p, p[:, :7])
  d2 = car_lib.SquaredDistanceMatrix(p, p[:, :7], mem_optimized=True)
  torch.testing.assert_close(d1, d2, atol=1e-4, rtol=1e-4)
  cn = car_lib.GenerateCenternessLabel(torch.tensor([[0., 0, 0], [0.9, 0, 0]]),
                                       torch.tensor([[0., 0, 0, 2, 2, 2, 0]] * 2), (0.0, 1.0))
  assert float(cn[0]) == 1.0 and 0 < float(cn[1]) < 1.0


def test_native_average_precision():
  gt = np.array([[0, 0, 0, 2, 2, 2, 0], [10, 0, 0, 2, 2, 2, 0], [20, 0, 0, 2, 2, 2, 0]], np.float32)
  pd = np.array([[0, 0, 0, 2, 2, 2, 0], [10.2, 0, 0, 2, 2, 2, 0], [50, 0, 0, 2, 2, 2, 0],
                 [20, 0, 0, 2, 2, 2, 0]], np.float32)
  ap, pr, sh = car_ops.average_precision3d(0.5, gt, [0, 0, 0], [0, 0, 0], pd, [0, 0, 0, 0],
                                           [0, 0, 0, 0], [0.9, 0.8, 0.7, 0.6],
                                           num_recall_points=10)
  assert sh[:, 1].tolist() == [1, 1, 0, 1]
  assert abs(ap - (7 + 4 * 0.75) / 11) < 1e-5 and pr.shape == (10, 2)
  # an "ignore all matches" region absorbs the false positive
  gt2 = np.concatenate([gt, [[50, 0, 0, 2, 2, 2, 0]]]).astype(np.float32)
  ap2, _, _ = car_ops.average_precision3d(0.5, gt2, [0] * 4, [0, 0, 0, 2], pd, [0] * 4, [0] * 4,
                                          [0.9, 0.8, 0.7, 0.6], num_recall_points=10)
  assert abs(ap2 - 1.0) < 1e-6
  iou = car_ops.pairwise_iou3d(gt[:1], np.array([[1, 0, 0, 2, 2, 2, 0]], np.float32))
  assert abs(float(iou) - 1 / 3) < 1e-5


def test_ap_metrics_and_calibration():
  md = kitti_metadata.KITTIMetadata()
  m = kitti_ap_metric.KITTIAPMetrics(kitti_ap_metric.KITTIAPMetrics.Params(md).Set(
      breakdown_metrics=['distance', 'num_points', 'rotation']))
  rng = np.random.RandomState(0)
  c = md.NumClasses()
  for s in range(6):
    g = 5
    boxes = np.concatenate([rng.uniform(-30, 30, (g, 2)), np.zeros((g, 1)),
                            np.tile([4, 2, 1.5], (g, 1)), rng.uniform(-3, 3, (g, 1))], 1).astype(
                                np.float32)
    labels = np.array([1, 1, 4, 6, 2])
    ds, db = np.zeros((c, 8), np.float32), np.zeros((c, 8, 7), np.float32)
    for i, (l, b) in enumerate(zip(labels, boxes)):
      if l in (1, 4, 6):
        ds[l, i], db[l, i] = rng.uniform(0.5, 1.0), b + rng.normal(0, 0.02, 7)
    ds[1, 7], db[1, 7] = 0.3, [50, 50, 0, 4, 2, 1.5, 0]
    m.Update('scene%d' % s, NestedMap(
        groundtruth_labels=labels, groundtruth_bboxes=boxes,
        groundtruth_difficulties=rng.randint(1, 4, g),
        groundtruth_num_points=rng.randint(1, 2000, g), detection_scores=ds,
        detection_boxes=db, detection_heights_in_pixels=np.full((c, 8), 50.0)))
  assert 0.8 < m.value <= 1.0
  sc = m.Scalars('kitti')
  assert 'kitti/AP_car_moderate' in sc and any('ByDistance' in k for k in sc)
  r = calibration_processing.CalibrationCalculator(md).Calculate(m)
  assert set(r) == {'Car', 'Cyclist', 'Pedestrian'} and 0 <= r['Car']['ece'] <= 1
  wm = waymo_metadata.WaymoMetadata()
  w = waymo_ap_metric.WaymoAPMetrics(waymo_ap_metric.WaymoAPMetrics.Params(wm).Set(
      waymo_breakdown_metrics=['RANGE']))
  cw = wm.NumClasses()
  for s in range(3):
    boxes = np.concatenate([rng.uniform(-60, 60, (4, 2)), np.zeros((4, 1)),
                            np.tile([4, 2, 1.5], (4, 1)), rng.uniform(-3, 3, (4, 1))], 1).astype(
                                np.float32)
    labels = np.array([1, 1, 2, 4])
    ds, db = np.zeros((cw, 6), np.float32), np.zeros((cw, 6, 7), np.float32)
    for i, (l, b) in enumerate(zip(labels, boxes)):
      ds[l, i], db[l, i] = 0.9, b
      db[l, i, 6] += 0.2
    w.Update('s%d' % s, NestedMap(groundtruth_labels=labels, groundtruth_bboxes=boxes,
                                  groundtruth_difficulties=np.array([1, 2, 1, 2]),
                                  groundtruth_num_points=np.full(4, 10), detection_scores=ds,
                                  detection_boxes=db))
  assert w.value > 0.99
  aph = w.Scalars('waymo')['waymo/APH_vehicle_LEVEL_2']
  assert abs(aph - (1 - 0.2 / math.pi)) < 1e-3


def test_visualisation():
  img = summary.DrawTopDown(np.random.randn(2000, 3) * 10)
  assert img.shape == (1024, 1024, 3) and img.sum() > 0
  td = summary.TransformBBoxesToTopDown(np.array([[[5, 0, 0, 4, 2, 1.5, 0.3]]], np.float32))
  out = summary.DrawBBoxesOnImages(img[None], td, np.ones((1, 1)), np.ones((1, 1), np.int32),
                                   {1: 'Car'}, groundtruth=False)
  assert out.shape == (1, 1024, 1024, 3)
  m = dm.TopDownVisualizationMetric(image_height=128, image_width=128)
  m.Update(NestedMap(
      source_ids=['a', 'b'], visualization_labels=np.ones((2, 3), np.int32),
      predicted_bboxes=np.random.rand(2, 3, 5) * 5, visualization_weights=np.random.rand(2, 3),
      gt_bboxes_2d=np.random.rand(2, 2, 5) * 5, gt_bboxes_2d_weights=np.ones((2, 2)),
      labels=np.ones((2, 2), np.int32), points_xyz=np.random.randn(2, 100, 3) * 5,
      points_padding=np.zeros((2, 100))))
  s = m.Summary('td')
  assert len(s) == 2 and s[0][1][:4] == b'\x89PNG'


def test_preprocessor_pipeline_and_inverses():
  f = _Feats()
  grid = ip.PointsToGrid.Params().Set(num_points_per_cell=8, grid_size=(16, 16, 1),
                                      grid_range_x=(0, 64), grid_range_y=(-32, 32),
                                      grid_range_z=(-3, 2))
  pipeline = ip.Sequence.Params().Set(name='all', preprocessors=[
      ip.CountNumberOfPointsInBoxes3D.Params(),
      ip.FilterGroundTruthByNumPoints.Params().Set(min_num_points=1),
      ip.RandomWorldRotationAboutZAxis.Params().Set(max_rotation=math.pi / 4, random_seed=1),
      ip.RandomFlipY.Params().Set(random_seed=2),
      ip.WorldScaling.Params().Set(scaling=(0.95, 1.05), random_seed=3),
      ip.GlobalTranslateNoise.Params().Set(random_seed=4),
      ip.RandomDropLaserPoints.Params().Set(random_seed=5),
      ip.FrustumDropout.Params().Set(theta_width=0.1, phi_width=0.2, random_seed=6),
      ip.PadLaserFeatures.Params().Set(max_num_points=512, random_seed=7),
      ip.DropLaserPointsOutOfRange.Params().Set(keep_x_range=(0, 70)),
      ip.AddPerPointLabels.Params(), grid,
      ip.GridToPillars.Params().Set(num_points_per_pillar=8, num_pillars=64, random_seed=8),
      ip.GridAnchorCenters.Params().Set(grid_size=(8, 8, 1), grid_range_x=(0, 64),
                                        grid_range_y=(-32, 32), grid_range_z=(-1, -1)),
      ip.PointPillarAnchorBoxSettingsCar.Update(ip.TileAnchorBBoxes.Params()),
      ip.AnchorAssignment.Params()]).Instantiate()
  out = pipeline.TransformFeatures(f)
  assert out.pillar_points.shape == (64, 8, 4) and out.anchor_bboxes.shape == (64, 2, 7)
  assert out.lasers.points_xyz.shape == (512, 3) and out.lasers.points_label.shape == (512,)
  assert float(out.labels.bboxes_3d_mask.sum()) >= 1
  f2 = _Feats()
  orig = f2.labels.bboxes_3d.clone()
  fw = ip.Sequence.Params().Set(name='fw', preprocessors=[
      ip.RandomWorldRotationAboutZAxis.Params().Set(max_rotation=1.0, random_seed=1),
      ip.RandomFlipY.Params().Set(flip_probability=1.0),
      ip.WorldScaling.Params().Set(scaling=(0.9, 1.1), random_seed=2),
      ip.GlobalTranslateNoise.Params().Set(random_seed=3)]).Instantiate()
  o = fw.TransformFeatures(f2)
  o.predicted_bboxes = o.labels.bboxes_3d.clone()
  inv = ip.Sequence.Params().Set(name='inv', preprocessors=[
      ip.InverseGlobalTranslateNoise.Params(), ip.InverseWorldScaling.Params(),
      ip.InverseRandomFlipY.Params(),
      ip.InverseRandomWorldRotationAboutZAxis.Params()]).Instantiate()
  o = inv.TransformFeatures(o)
  torch.testing.assert_close(o.predicted_bboxes[:, :6], orig[:, :6], atol=1e-4, rtol=1e-4)
  choice = ip.RandomChoicePreprocessor.Params().Set(name='rc', subprocessors=[
      (ip.IdentityPreprocessor.Params(), 1.0),
      (ip.RandomDropLaserPoints.Params().Set(keep_prob=1.0), 2.0)]).Instantiate()
  assert choice.TransformFeatures(_Feats()).lasers.points_xyz.shape == (400, 3)
  gta = ip.GroundTruthAugmentor.Params().Set(name='gta', random_seed=0, groundtruth_database=[
      dict(bbox_3d=[200., 200, 0, 4, 2, 1.5, 0], label=1,
           points_xyz=np.random.rand(10, 3) + [199.5, 199.5, -0.5],
           points_feature=np.random.rand(10, 1))]).Instantiate()
  aug = gta.TransformFeatures(_Feats())
  assert float(aug.labels.bboxes_3d_mask.sum()) == 5 and aug.lasers.points_xyz.shape[0] == 410


def test_nms_decoders():
  boxes = torch.tensor([[[0, 0, 0, 4, 2, 1.5, 0], [0.2, 0, 0, 4, 2, 1.5, 0],
                         [10, 0, 0, 4, 2, 1.5, 0.3], [20, 5, 0, 1, 1, 1.7, 0]]]).float()
  scores = torch.tensor([[[0.9, 0.1], [0.8, 0.05], [0.7, 0.2], [0.1, 0.95]]])
  _, b, s, m = dd.DecodeWithNMS(boxes, scores, 0.3, 0.05, 4, use_oriented_per_class_nms=True)
  assert m[0, 0].tolist() == [1, 1, 1, 0] and abs(float(s[0, 1, 0]) - 0.95) < 1e-6
  hm = torch.zeros(1, 8, 8, 1)
  hm[0, 2, 3, 0], hm[0, 2, 4, 0], hm[0, 6, 6, 0] = 0.9, 0.5, 0.7
  n = dd.HeatMapNMS(hm, (3, 3), 3, 0.1)
  assert n.top_k_indices[0, 0, :2].tolist() == [[2, 3], [6, 6]]


def test_pointnet():
  pts = NestedMap(points=torch.randn(2, 64, 3), features=torch.randn(2, 64, 1),
                  padding=torch.zeros(2, 64))
  pts.padding[1, 40:] = 1
  b = pointnet.PointNet(pointnet.PointNet.Params())
  c = b.Classifier(input_dims=4, feature_dims=32).Instantiate()
  assert c.FProp(c.theta, pts).shape == (2, 32)
  s = b.Segmentation(input_dims=4).Instantiate()
  assert s.FProp(s.theta, pts).features.shape == (2, 64, 128)


def test_kitti_pipelines_and_models(kitti_file):
  p = _Small(kig.KITTISparseLaser.Params(), kitti_file)
  p.preprocessors.select_centers.num_cell_centers = 16
  p.preprocessors.gather_features.num_points_per_cell = 8
  batch = p.Instantiate().GetPreprocessedInputBatch()
  assert batch.labels.difficulties[0, :3].tolist() == [3, 2, 1]
  assert batch.cell_feature.shape == (2, 16, 8, 1) and batch.anchor_bboxes.shape == (2, 16, 12, 7)
  batch.assigned_gt_labels = batch.assigned_gt_labels.clamp(max=1)
  for cls in (starnet.ModelV1, starnet.ModelV2):
    mp = cls.Params(num_classes=2, num_anchor_bboxes_offsets=4, num_anchor_bboxes_rotations=3)
    mp.max_nms_boxes = 8
    with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
      task = mp.Instantiate()
    pred = task.ComputePredictions(task.theta, batch)
    m, _ = task.ComputeLoss(task.theta, pred, batch)
    m.loss[0].backward()
    assert torch.isfinite(m.loss[0])
    with cluster_factory.SetEval(True):
      out = task.Decode(batch)
    metrics = task.CreateDecoderMetrics()
    task.PostProcessDecodeOut(out, metrics)
    assert out.per_class_predicted_bboxes.shape == (2, 2, 8, 7)
    assert 0.0 <= metrics['kitti_AP_v2'].value <= 1.0
  # PointPillars on the grid input
  g = _Small(kig.KITTIGrid.Params(), kitti_file)
  g.preprocessors.points_to_grid.Set(grid_size=(16, 16, 1), num_points_per_cell=8,
                                     grid_range_x=(0, 64), grid_range_y=(-32, 32))
  g.preprocessors.keep_xyz_range.Set(keep_x_range=(0, 64), keep_y_range=(-32, 32))
  g.preprocessors.grid_to_pillars.Set(num_pillars=64, num_points_per_pillar=8)
  g.preprocessors.grid_anchor_centers.Set(grid_size=(8, 8, 1), grid_range_x=(0, 64),
                                          grid_range_y=(-32, 32))
  gb = g.Instantiate().GetPreprocessedInputBatch()
  mp = pillars.ModelV2.Params()
  mp.input_featurizer.grid_size = (16, 16, 1)
  mp.max_nms_boxes = 8
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = mp.Instantiate()
  pred = task.ComputePredictions(task.theta, gb)
  assert pred.residuals.shape == gb.anchor_bboxes.shape
  m, _ = task.ComputeLoss(task.theta, pred, gb)
  m.loss[0].backward()
  # anchor-free head + heat-map NMS + Laplace-KL on the same pillars
  gp = ip.GridAnchorCenters.Params().Set(grid_size=(16, 16, 1), grid_range_x=(0, 64),
                                         grid_range_y=(-32, 32), grid_range_z=(-1, -1)).Instantiate()
  gb.anchor_centers = gp.TransformFeatures(NestedMap()).anchor_centers.unsqueeze(0).expand(
      2, -1, -1).contiguous()
  pa = ip.PointAssignment.Params().Set(num_classes=2).Instantiate()
  outs = [pa.TransformFeatures(NestedMap(
      anchor_centers=gb.anchor_centers[i],
      labels=NestedMap(bboxes_3d=gb.labels.bboxes_3d[i], labels=gb.labels.labels[i].clamp(max=1),
                       bboxes_3d_mask=gb.labels.bboxes_3d_mask[i]))) for i in range(2)]
  for k in ('target_predictions', 'assigned_gt_bbox', 'assigned_gt_labels',
            'assigned_gt_center_ness', 'assigned_cls_mask', 'assigned_reg_mask'):
    gb[k] = torch.stack([o[k] for o in outs])
  ap = paf.ModelV1.Params(angle_bin_num=4)
  ap.input_featurizer.grid_size = (16, 16, 1)
  ap.Set(max_nms_boxes=8, nms_decoder_type=paf.NMSDecoderType.HEATMAP_NMS_DECODER,
         centerness_loss_weight=1.0, location_loss=paf.LaplaceKL.Params())
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    t2 = ap.Instantiate()
  pr = t2.ComputePredictions(t2.theta, gb)
  mm, _ = t2.ComputeLoss(t2.theta, pr, gb)
  mm.loss[0].backward()
  with cluster_factory.SetEval(True):
    o2 = t2.Decode(gb)
  assert pr.residuals.shape == (2, 256, 17) and o2.per_class_predicted_bboxes.shape == (2, 2, 8, 7)
  # camera fusion
  gb.images.image = torch.rand(2, 32, 48, 3)
  gb.pillar_points_projected = torch.rand(2, 64, 8, 2) * torch.tensor([48., 32.])
  fp = deep_fusion.MultiModalFeaturizer.Params(aligner='deep_fusion', image_channels=24)
  fp.lidar_featurizer.grid_size = (16, 16, 1)
  fp.name = 'mm'
  mmf = fp.Instantiate()
  assert mmf.FProp(mmf.theta, gb).shape == (2, 16, 16, 64)


def test_waymo_input(tmp_path):
  w = ops.host().TFRecordWriter(str(tmp_path / 'w.tfrecord'))
  rng = np.random.RandomState(0)
  for i in range(4):
    feats = {'pose': np.eye(4, dtype=np.float32).reshape(-1), 'run_segment': [b'seg_%d' % i],
             'run_start_offset': np.asarray([i * 100]), 'time_of_day': [b'Day'],
             'location': [b'location_sf'], 'weather': [b'sunny']}
    for l in wig.LIDAR_NAMES:
      for r in ('ri1', 'ri2'):
        pts = np.concatenate([rng.uniform(-40, 40, (200, 2)), rng.uniform(-1, 2, (200, 1)),
                              rng.rand(200, 2), rng.choice([-1.0, 1.0], (200, 1), p=[0.9, 0.1])],
                             1).astype(np.float32)
        feats['laser_%s_%s' % (l, r)] = pts.reshape(-1)
    boxes = np.concatenate([rng.uniform(-30, 30, (5, 2)), np.zeros((5, 1)),
                            np.tile([4.7, 2.1, 1.7], (5, 1)), rng.uniform(-3, 3, (5, 1))],
                           1).astype(np.float32)
    feats.update({'labels': np.array([1, 1, 2, 4, 1]), 'label_ids': [b'a', b'b', b'c', b'd', b'e'],
                  'detection_difficulties': np.array([1, 2, 1, 2, 1]),
                  'single_frame_detection_difficulties': np.array([1, 2, 1, 2, 1]),
                  'tracking_difficulties': np.ones(5, np.int64), 'bboxes_3d': boxes.reshape(-1),
                  'bboxes_3d_num_points': np.array([10, 3, 50, 7, 0]),
                  'label_metadata': rng.randn(20).astype(np.float32)})
    w.write(tf_example.MakeExample(feats))
  w.close()
  p = _Small(wig.WaymoSparseLaser.Params(), 'tfrecord:' + str(tmp_path / 'w.tfrecord'))
  p.preprocessors.pad_lasers.max_num_points = 2048
  p.preprocessors.viz_copy.pad_lasers.max_num_points = 2048
  p.preprocessors.select_centers.num_cell_centers = 32
  p.preprocessors.gather_features.num_points_per_cell = 16
  p.extractors.labels.max_num_objects = 16
  b = p.Instantiate().GetPreprocessedInputBatch()
  assert b.cell_feature.shape == (2, 32, 16, 3) and b.labels.speed.shape == (2, 16, 2)
  assert b.metadata.pose.shape == (2, 4, 4)
  # no-label-zone points were removed: fewer real points than the 2000 written
  assert int((b.lasers.points_padding[0] < 0.5).sum()) < 2000


def test_registered_car_params():
  import lingvo_b200.models.car.params.params  # noqa: F401
  names = [n for n in model_registry.GetAllRegisteredClasses() if n.startswith('car.')]
  assert len(names) >= 18
  for n in names:
    for ds in ('Train', 'Dev'):
      try:
        mp = model_registry.GetParams(n, ds)
      except NotImplementedError:
        continue
      assert mp.task.name
  mp = model_registry.GetParams('car.waymo_deepfusion.DeepFusionCenterPointPed', 'Train')
  assert 'images' in mp.input.extractors and mp.task.location_loss.cls is paf.LaplaceKL
  assert os.environ.get('LINGVO_B200_KITTI', '/tmp/kitti/') in model_registry.GetParams(
      'car.kitti.StarNetCarModel0701', 'Train').input.file_pattern


def test_breakdown_metrics_bins_and_tables():
  from lingvo_b200.core.nested_map import NestedMap
  from lingvo_b200.models.car import breakdown_metric as bm
  from lingvo_b200.models.car import kitti_metadata
  meta = kitti_metadata.KITTIMetadata()
  dist = bm.ByName('distance').Params().Set(metadata=meta).Instantiate()
  w = meta.DistanceBinWidth()
  boxes = np.array([[w * 0.5, 0, 0, 1, 1, 1, 0.0], [w * 2.2, 0, 0, 1, 1, 1, 2.0],
                    [1e6, 0, 0, 1, 1, 1, 0.0]], np.float32)
  bins = dist.Discretize(boxes)
  assert bins.tolist() == [0, 2, dist.NumBinsOfHistogram() - 1]              # far boxes clip
  dist.AccumulateHistogram(NestedMap(bboxes=boxes, labels=np.array([1, 1, 4])))
  assert dist._histogram[0, 1] == 1 and dist._histogram[2, 1] == 1 and dist._histogram[-1, 4] == 1
  assert len(dist.BinLabels()) == dist.NumBinsOfHistogram()
  rot = bm.ByRotation.Params().Set(metadata=meta).Instantiate()
  rb = rot.Discretize(np.array([[0, 0, 0, 1, 1, 1, 0.01], [0, 0, 0, 1, 1, 1, meta.MaximumRotation() - 0.01]]))
  assert rb[0] == 0 and rb[1] == rot.NumBinsOfHistogram() - 1
  pts = bm.ByNumPoints.Params().Set(metadata=meta).Instantiate()
  pb = pts.Discretize([1, 10, 10 ** 9])
  assert pb[0] == 0 and pb[0] <= pb[1] <= pb[2] == pts.NumBinsOfHistogram() - 1
  pts.AccumulateCumulative(NestedMap(num_points=np.array([1, 10])))
  assert pts._values.sum() == 11
  diff = bm.ByDifficulty.Params().Set(metadata=meta).Instantiate()
  assert diff.BinLabels()[-1] == 'default' and diff.NumBinsOfHistogram() == 4
  # ComputeMetrics fills AP / recall tables from whatever the AP metric returns per bin
  n_eval, pr_pts = len(meta.EvalClassIndices()), meta.NumberOfPrecisionRecallPoints()
  calls = []
  def Fn(difficulty=None):
    calls.append(difficulty)
    if difficulty == 'hard':
      return None                                                     # no ground truth in this bin
    pr = np.zeros((n_eval, pr_pts, 2), np.float32)
    pr[..., 0] = 0.9
    pr[..., 1] = np.linspace(1.0, 0.0, pr_pts)[None, :]
    return np.full(n_eval, 0.5, np.float32), pr
  diff.ComputeMetrics(Fn)
  assert calls == ['hard', 'moderate', 'easy', None]
  scalars, images = diff.GenerateSummaries('kitti')
  assert any(k.endswith('_default') and v == 0.5 for k, v in scalars.items())
  assert not any('_hard' in k for k in scalars)
  assert isinstance(images, list)
  assert np.allclose(diff._max_recall[1:], 1.0)
  with pytest.raises(Exception):
    bm.ByName('nope')
