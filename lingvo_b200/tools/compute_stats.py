"""Per-dimension mean/std of a float feature over tf.Example records
(ref `lingvo/tools/compute_stats.py`, used to normalise ASR features)."""
import numpy as np
from absl import app
from absl import flags

from lingvo_b200 import ops
from lingvo_b200.utils import tf_example

flags.DEFINE_string('input_filepattern', '', 'tfrecord glob.')
flags.DEFINE_string('feature_name', 'frames', 'Flattened float feature.')
flags.DEFINE_integer('frame_size', 80, 'Feature dims per frame.')
FLAGS = flags.FLAGS


def ComputeStats(pattern, feature_name, frame_size):
  if ':' not in pattern.split('/')[0]:
    pattern = 'tfrecord:' + pattern
  y = ops.host().sequential_record_yielder(pattern, repeat_count=1)
  n = 0
  s = np.zeros(frame_size, np.float64)
  ss = np.zeros(frame_size, np.float64)
  while True:
    rec = y.next()
    if rec is None:
      break
    x = tf_example.ParseExample(rec[0])[feature_name].reshape(-1, frame_size).astype(np.float64)
    n += x.shape[0]
    s += x.sum(0)
    ss += (x * x).sum(0)
  mean = s / max(n, 1)
  std = np.sqrt(np.maximum(ss / max(n, 1) - mean ** 2, 1e-12))
  return mean, std, n


def main(argv):
  del argv
  mean, std, n = ComputeStats(FLAGS.input_filepattern, FLAGS.feature_name, FLAGS.frame_size)
  print('frames: %d\nmean: %s\nstddev: %s' % (n, mean.tolist(), std.tolist()))


if __name__ == '__main__':
  app.run(main)
