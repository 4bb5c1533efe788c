"""Encodes parallel text into NMT tf.Example records with a word-piece model
(ref `lingvo/tools/wpm_encode_file.py`)."""
import numpy as np
from absl import app
from absl import flags

from lingvo_b200 import ops
from lingvo_b200.core import wpm_encoder
from lingvo_b200.utils import tf_example

flags.DEFINE_string('source_filepath', '', 'Source sentences, one per line.')
flags.DEFINE_string('target_filepath', '', 'Target sentences, one per line.')
flags.DEFINE_string('wpm_filepath', '', 'Word-piece vocabulary.')
flags.DEFINE_string('output_filepath', '', 'Output TFRecord file.')
flags.DEFINE_integer('num_shards', -1, 'Total shards of the job (-1: single).')
flags.DEFINE_integer('shard_id', -1, 'This worker\'s shard.')
flags.DEFINE_integer('max_len', 0, 'Drop pairs longer than this (0: keep all).')
FLAGS = flags.FLAGS


def _MakeExample(enc, src, tgt):
  s_ids, _ = enc.Encode(src)
  t_ids, _ = enc.Encode(tgt)
  s_ids = s_ids + [enc.sentence_end_id]
  return tf_example.MakeExample({
      'source_id': np.asarray(s_ids, np.int64),
      'source_padding': np.zeros(len(s_ids), np.float32),
      'target_id': np.asarray([enc.sentence_start_id] + t_ids, np.int64),
      'target_padding': np.zeros(len(t_ids) + 1, np.float32),
      'target_label': np.asarray(t_ids + [enc.sentence_end_id], np.int64),
      'target_weight': np.ones(len(t_ids) + 1, np.float32)}), max(len(s_ids), len(t_ids) + 1)


def main(argv):
  del argv
  enc = wpm_encoder.WpmEncoder(FLAGS.wpm_filepath)
  out = FLAGS.output_filepath
  if FLAGS.num_shards > 0:
    out = '%s-%05d-of-%05d' % (out, FLAGS.shard_id, FLAGS.num_shards)
  w = ops.host().TFRecordWriter(out)
  n = 0
  with open(FLAGS.source_filepath, encoding='utf-8') as fs, \
      open(FLAGS.target_filepath, encoding='utf-8') as ft:
    for i, (s, t) in enumerate(zip(fs, ft)):
      if FLAGS.num_shards > 0 and i % FLAGS.num_shards != FLAGS.shard_id:
        continue
      ex, length = _MakeExample(enc, s.strip(), t.strip())
      if FLAGS.max_len and length > FLAGS.max_len:
        continue
      w.write(ex)
      n += 1
  w.close()
  print('wrote %d examples to %s' % (n, out))


if __name__ == '__main__':
  app.run(main)
