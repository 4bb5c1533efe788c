"""Car data tools: KITTI parsers / exporter / detection export / crops, Waymo converter."""

import math
import os
import pickle
import struct
import zlib

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.models.car.tools import create_kitti_crop_dataset
from lingvo_b200.models.car.tools import export_kitti_detection
from lingvo_b200.models.car.tools import kitti_data
from lingvo_b200.models.car.tools import kitti_exporter
from lingvo_b200.models.car.waymo import export_to_submission_format
from lingvo_b200.models.car.waymo.tools import waymo_proto_to_tfe as w2t
from lingvo_b200.utils import protowire as pw
from lingvo_b200.utils import tf_example

CALIB = """P0: 1 0 0 0 0 1 0 0 0 0 1 0
P2: 700 0 600 0 0 700 180 0 0 0 1 0
R0_rect: 1 0 0 0 1 0 0 0 1
Tr_velo_to_cam: 0 -1 0 0 0 0 -1 0 1 0 0 0
Tr_imu_to_velo: 1 0 0 0 0 1 0 0 0 0 1 0
"""
LABEL = ("Car 0.00 0 -1.5 600 150 700 250 1.5 1.6 3.9 1.0 1.6 20.0 -1.57\n"
         "DontCare -1 -1 -10 500 100 520 120 -1 -1 -1 -1000 -1000 -1000 -10\n")


def _Png(w=8, h=4):
  """Minimal valid PNG header + IHDR (enough for the exporter's size probe)."""
  ihdr = struct.pack('>IIBBBBB', w, h, 8, 2, 0, 0, 0)
  return b'\x89PNG\r\n\x1a\n' + struct.pack('>I', 13) + b'IHDR' + ihdr + b'\0\0\0\0'


def _MakeKitti(root):
  for sub in ('velodyne', 'calib', 'image_2', 'label_2'):
    os.makedirs(os.path.join(root, 'training', sub), exist_ok=True)
  rng = np.random.RandomState(0)
  pts = np.concatenate([rng.uniform(0, 40, (200, 1)), rng.uniform(-10, 10, (200, 1)),
                        rng.uniform(-2, 0, (200, 1)), rng.rand(200, 1)], 1).astype(np.float32)
  pts[:30, :3] = [20.0, -1.0, -0.85] + rng.uniform(-0.4, 0.4, (30, 3))
  pts.tofile(os.path.join(root, 'training', 'velodyne', '000001.bin'))
  open(os.path.join(root, 'training', 'calib', '000001.txt'), 'w').write(CALIB)
  open(os.path.join(root, 'training', 'label_2', '000001.txt'), 'w').write(LABEL)
  open(os.path.join(root, 'training', 'image_2', '000001.png'), 'wb').write(_Png())
  open(os.path.join(root, 'train.txt'), 'w').write('000001\n')


def test_kitti_tools_roundtrip(tmp_path):
  root = str(tmp_path / 'kitti')
  _MakeKitti(root)
  calib = kitti_data.LoadCalibrationFile(os.path.join(root, 'training/calib/000001.txt'))
  objs = kitti_data.AnnotateKITTIObjectsWithBBox3D(
      kitti_data.LoadLabelFile(os.path.join(root, 'training/label_2/000001.txt')), calib)
  assert [o['has_3d_info'] for o in objs] == [True, False]
  box = objs[0]['bbox3d']
  # camera (x right, y down, z fwd) = (1, 1.6, 20) bottom centre → velodyne (20, −1, −1.6 + h/2)
  np.testing.assert_allclose(box[:3], [20.0, -1.0, -0.85], atol=1e-6)
  np.testing.assert_allclose(box[3:6], [3.9, 1.6, 1.5], atol=1e-6)
  loc, dims, rot = kitti_data.BBox3DToKITTIObject(box, kitti_data.VeloToCameraTransformation(calib))
  np.testing.assert_allclose(loc, [1.0, 1.6, 20.0], atol=1e-5)
  assert abs(rot - (-1.57)) < 1e-6 and dims == [1.5, 1.6, 3.9]
  out = str(tmp_path / 'kitti.tfrecord')
  n = kitti_exporter._ExportObjectDatasetToTFRecord(root, 'train', os.path.join(root, 'train.txt'),
                                                    out, 1)
  assert n == 1
  y = ops.host().sequential_record_yielder('tfrecord:%s-00000-of-00001' % out, 1)
  rec = y.next()[0]
  ex = tf_example.ParseExample(rec)
  assert ex['image/width'][0] == 8 and list(ex['object/label']) == [b'Car', b'DontCare']
  crops = create_kitti_crop_dataset.CropObjects(rec)
  assert len(crops) == 1
  crop = tf_example.ParseExample(crops[0])
  assert crop['num_points'][0] >= 25 and list(crop['label']) == [b'Car']
  det = dict(source_id=b'000001', bboxes=np.asarray([box], np.float32), scores=np.asarray([0.9]),
             class_ids=np.asarray([1]))
  cal = export_kitti_detection.LoadCalibData(os.path.join(root, 'training/calib/000001.txt'))
  lines = export_kitti_detection.ExtractNpContent(det, cal)
  path = export_kitti_detection.ExportKITTIDetection(str(tmp_path / 'sub'), '000001', lines,
                                                     ['Background', 'Car'])
  cols = open(path).read().split()
  assert cols[0] == 'Car' and len(cols) == 16 and abs(float(cols[13]) - 20.0) < 1e-3


def _Frame():
  """A tiny synthetic waymo Frame proto."""
  def transform(m):
    return pw.f_packed_double(1, np.asarray(m, np.float64).reshape(-1))
  ri = np.zeros((4, 8, 4), np.float32)
  ri[..., 0] = 10.0                       # 10 m everywhere
  ri[0, 0, 0] = 0.0                       # one invalid pixel
  ri[..., 3] = -1.0
  matrix = pw.f_bytes(1, ri.tobytes()) + pw.f_bytes(2, pw.f_packed_varint(1, ri.shape))
  range_image = pw.f_bytes(w2t.RANGE_IMAGE['range_image_compressed'], zlib.compress(matrix))
  laser = pw.f_varint(w2t.LASER['name'], 1) + pw.f_bytes(w2t.LASER['ri_return1'], range_image)
  lcal = (pw.f_varint(w2t.LASER_CALIB['name'], 1) +
          pw.f_packed_double(w2t.LASER_CALIB['beam_inclinations'], np.linspace(-0.3, 0.04, 4)) +
          pw.f_bytes(w2t.LASER_CALIB['extrinsic'], transform(np.eye(4))))
  stats = (pw.f_bytes(w2t.STATS['time_of_day'], b'Day') + pw.f_bytes(w2t.STATS['location'], b'sf') +
           pw.f_bytes(w2t.STATS['weather'], b'sunny'))
  ctx = (pw.f_bytes(w2t.CONTEXT['name'], b'segment-123') +
         pw.f_bytes(w2t.CONTEXT['laser_calibrations'], lcal) + pw.f_bytes(w2t.CONTEXT['stats'], stats))
  box = b''.join(pw.f_double(w2t.BOX[k], v) for k, v in dict(
      center_x=5.0, center_y=1.0, center_z=0.5, width=2.0, length=4.5, height=1.6,
      heading=0.3).items())
  label = (pw.f_bytes(w2t.LABEL['box'], box) + pw.f_varint(w2t.LABEL['type'], 1) +
           pw.f_bytes(w2t.LABEL['id'], b'obj0') + pw.f_varint(w2t.LABEL['num_lidar_points_in_box'], 3))
  return (pw.f_bytes(w2t.FRAME['context'], ctx) + pw.f_varint(w2t.FRAME['timestamp_micros'], 1234) +
          pw.f_bytes(w2t.FRAME['pose'], transform(np.eye(4))) +
          pw.f_bytes(w2t.FRAME['lasers'], laser) + pw.f_bytes(w2t.FRAME['laser_labels'], label))


def test_waymo_frame_to_tfe_and_submission(tmp_path):
  ex = tf_example.ParseExample(w2t.FrameToTFE().process(_Frame()))
  assert list(ex['run_segment']) == [b'segment-123'] and ex['run_start_offset'][0] == 1234
  pts = ex['laser_TOP_ri1'].reshape(-1, 6)
  assert pts.shape == (31, 6)                                  # 32 pixels − 1 invalid
  np.testing.assert_allclose(np.linalg.norm(pts[:, :3], axis=1), 10.0, rtol=1e-4)
  np.testing.assert_allclose(ex['bboxes_3d'], [5.0, 1.0, 0.5, 4.5, 2.0, 1.6, 0.3], rtol=1e-6)
  assert ex['single_frame_detection_difficulties'].tolist() == [2]     # ≤ 5 points → LEVEL_2
  assert ex['TOP_ri1_shape'].tolist() == [4, 8, 4]
  dump = [('k', dict(frame_id='segment-123_1234', bboxes=np.asarray([[5, 1, .5, 4.5, 2, 1.6, .3]]),
                     scores=np.asarray([0.8]), class_ids=np.asarray([1])))]
  path = str(tmp_path / 'dec.pkl')
  pickle.dump(dump, open(path, 'wb'))
  blob = export_to_submission_format.convert_detections(path)
  objs = pw.parse_dict(blob)[1]
  o = pw.parse_dict(objs[0])
  assert o[4] == [b'segment-123'] and o[5] == [1234] and abs(pw.as_float(o[2][0]) - 0.8) < 1e-6
  box = pw.parse_dict(pw.parse_dict(o[1][0])[1][0])
  assert abs(pw.as_double(box[5][0]) - 4.5) < 1e-9 and abs(pw.as_double(box[7][0]) - 0.3) < 1e-9
  assert math.isclose(pw.as_double(box[4][0]), 2.0)
