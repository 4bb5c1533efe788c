"""Counts records matching a file pattern (ref `lingvo/tools/count_records.py`)."""
from absl import app
from absl import flags

from lingvo_b200 import ops

flags.DEFINE_string('input', '', 'type:glob, e.g. tfrecord:/data/train-*')
FLAGS = flags.FLAGS


def CountRecords(pattern):
  if ':' not in pattern.split('/')[0]:
    pattern = 'tfrecord:' + pattern
  y = ops.host().sequential_record_yielder(pattern, repeat_count=1)
  n = 0
  while y.next() is not None:
    n += 1
  return n


def main(argv):
  del argv
  print('%d records' % CountRecords(FLAGS.input))


if __name__ == '__main__':
  app.run(main)
