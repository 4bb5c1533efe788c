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


class StatsCollector:
  """Streams parsed tf.Examples (`{feature: array}` dicts): per-dimension mean / std of a
  float feature plus the distribution of sequence lengths, from which bucket upper bounds
  for the bucketing batcher are proposed (ref :29)."""

  def __init__(self, feature_name='frames', frame_size=80, num_buckets=8):
    self._feature_name, self._frame_size, self._num_buckets = feature_name, frame_size, num_buckets
    self._num_examples = 0
    self._lengths = []
    self._num_frames = 0
    self._mean_acc = np.zeros(frame_size, np.float64)
    self._var_acc = np.zeros(frame_size, np.float64)

  def Accumulate(self, example):
    self._num_examples += 1
    v = np.asarray(example[self._feature_name])
    num_frames = v.size // self._frame_size
    if v.dtype.kind == 'f':
      frames = v.reshape(-1, self._frame_size).astype(np.float64)
      self._num_frames += frames.shape[0]
      self._mean_acc += frames.sum(0)
      self._var_acc += (frames * frames).sum(0)
    elif v.dtype.kind not in 'iu':
      raise ValueError('Only float / int64 lists are supported: %s' % v.dtype)
    self._lengths.append(num_frames)

  def MeanVar(self):
    """→ (mean, stddev) per feature dimension."""
    n = max(self._num_frames, 1)
    mu = self._mean_acc / n
    return mu, np.sqrt(np.maximum(self._var_acc / n - mu * mu, 0.0))

  def LengthBuckets(self):
    """→ (equal-population bucket upper limits, {loss fraction: last-bucket candidate})."""
    lengths = sorted(self._lengths)
    n = len(lengths)
    idx = (n * (np.arange(self._num_buckets - 1) + 1)) // self._num_buckets
    buckets = [lengths[i] for i in idx] + [lengths[-1]]
    alternatives = {loss: lengths[min(int(n * (1.0 - loss)), n - 1)]
                    for loss in (0.001, 0.01, 0.02)}
    return buckets, alternatives

  def Print(self):
    print('== Total number of examples: %u' % self._num_examples)
    buckets, alt = self.LengthBuckets()
    print('== Buckets.\nbucket upper limits: %s\nOther candidates for last bucket:' % buckets)
    for loss, length in sorted(alt.items()):
      print('  %4.1f%% loss: %u' % (loss * 100.0, length))
    mean, std = self.MeanVar()
    print('== Mean/variance.\nmean = %s\nvar = %s' % (mean.tolist(), std.tolist()))


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
