"""Decoder outputs → Waymo submission objects (ref
`lingvo/tasks/car/waymo/export_to_submission_format.py`).

  python -m lingvo_b200.models.car.waymo.export_to_submission_format \
      --decoder_path=/logdir/decoder_test/decoder_out_000010000 --output_path=/tmp/preds.bin

Writes a serialized `waymo.open_dataset.Objects` message (`metrics.proto`): every
detection becomes an `Object { context_name, frame_timestamp_micros, score, object {
box, type } }`, encoded with the in-repo protobuf wire writer.
"""

import argparse
import pickle
import sys

import numpy as np

from lingvo_b200.utils import protowire as pw

# metrics.proto / label.proto field numbers
OBJECTS_OBJECTS = 1
OBJECT = dict(object=1, score=2, context_name=4, frame_timestamp_micros=5)
LABEL = dict(box=1, type=3)
BOX = dict(center_x=1, center_y=2, center_z=3, width=4, length=5, height=6, heading=7)


def _Box(b):
  x, y, z, dx, dy, dz, phi = [float(v) for v in b]
  vals = dict(center_x=x, center_y=y, center_z=z, length=dx, width=dy, height=dz, heading=phi)
  return b''.join(pw.f_double(BOX[k], v) for k, v in vals.items())


def _Object(context_name, timestamp, box, score, cls):
  label = pw.f_bytes(LABEL['box'], _Box(box)) + pw.f_varint(LABEL['type'], int(cls))
  return (pw.f_bytes(OBJECT['object'], label) + pw.f_float(OBJECT['score'], float(score)) +
          pw.f_bytes(OBJECT['context_name'], context_name.encode()) +
          pw.f_varint(OBJECT['frame_timestamp_micros'], int(timestamp)))


def convert_detections(table_path):  # pylint: disable=invalid-name
  """Decoder dump → serialized `Objects` (ref :40)."""
  with open(table_path, 'rb') as f:
    dump = pickle.load(f)   # noqa: S301  (our own decoder output)
  out = []
  for item in dump:
    d = item[1] if isinstance(item, (tuple, list)) else item
    frame = str(d['frame_id'].decode() if isinstance(d['frame_id'], bytes) else d['frame_id'])
    seg, ts = frame.rsplit('_', 1)
    for box, score, cls in zip(np.asarray(d['bboxes']), np.asarray(d['scores']),
                               np.asarray(d['class_ids'])):
      if score > 0:
        out.append(pw.f_bytes(OBJECTS_OBJECTS, _Object(seg, ts, box, score, cls)))
  return b''.join(out)


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--decoder_path', required=True)
  ap.add_argument('--output_path', required=True)
  a = ap.parse_args(argv)
  data = convert_detections(a.decoder_path)
  with open(a.output_path, 'wb') as f:
    f.write(data)
  print('wrote %d bytes to %s' % (len(data), a.output_path))
  return 0


if __name__ == '__main__':
  sys.exit(main())
