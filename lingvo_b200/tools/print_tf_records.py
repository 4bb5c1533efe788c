"""Prints records of a TFRecord file (ref `lingvo/tools/print_tf_records.py`).

  python -m lingvo_b200.tools.print_tf_records --input_filepattern='tfrecord:/x/*' \\
      [--skip_first_n=0] [--print_only_n=10] [--bytes_as_utf8]
"""
import sys

from absl import app
from absl import flags

from lingvo_b200 import ops
from lingvo_b200.utils import tf_example

flags.DEFINE_string('input_filepattern', '', 'type:glob of the records to print.')
flags.DEFINE_integer('skip_first_n', 0, 'Records to skip.')
flags.DEFINE_integer('print_only_n', -1, 'Max records to print (-1: all).')
flags.DEFINE_bool('bytes_as_utf8', False, 'Decode bytes features as UTF-8.')
FLAGS = flags.FLAGS


def main(argv):
  del argv
  pat = FLAGS.input_filepattern
  if ':' not in pat.split('/')[0]:
    pat = 'tfrecord:' + pat
  y = ops.host().sequential_record_yielder(pat, repeat_count=1)
  n = 0
  while True:
    rec = y.next()
    if rec is None:
      break
    n += 1
    if n <= FLAGS.skip_first_n:
      continue
    if 0 <= FLAGS.print_only_n < n - FLAGS.skip_first_n:
      break
    try:
      ex = tf_example.ParseExample(rec[0])
      print('--- record %d' % n)
      for k in sorted(ex):
        v = ex[k]
        if v.dtype == object and FLAGS.bytes_as_utf8:
          v = [x.decode('utf-8', 'replace') for x in v]
        print('%s: %s' % (k, v if len(v) <= 32 else '%s ... (%d values)' % (v[:32], len(v))))
    except Exception:  # pylint: disable=broad-except
      print('--- record %d (%d raw bytes): %r' % (n, len(rec[0]), rec[0][:120]))
  return 0


if __name__ == '__main__':
  app.run(main)
