"""Builds ASR training records from a tarball/directory of WAV files + transcripts
(ref `lingvo/tools/create_asr_features.py`).

  python -m lingvo_b200.tools.create_asr_features --input_dir=… --transcripts=trans.txt \\
      --output_template=/out/train.tfrecords-%5.5d-of-%5.5d --num_output_shards=10

`transcripts`: lines of `<utt-id> <text>`; audio is `<input_dir>/<utt-id>.wav`.
"""
import os

from absl import app
from absl import flags

from lingvo_b200 import ops
from lingvo_b200.tools import audio_lib
from lingvo_b200.utils import tf_example

flags.DEFINE_string('input_dir', '', 'Directory with <utt-id>.wav files.')
flags.DEFINE_string('transcripts', '', 'File with "<utt-id> <text>" lines.')
flags.DEFINE_string('output_template', '', 'e.g. /out/train.tfrecords-%5.5d-of-%5.5d')
flags.DEFINE_integer('num_output_shards', 1, 'Number of output shards.')
flags.DEFINE_integer('num_mel_bins', 80, 'Mel bins.')
FLAGS = flags.FLAGS


def WriteFeatures(utterances, output_template, n_shards=1, num_mel_bins=80):
  """`utterances`: iterable of (utt id, transcript, wav bytes) → sharded TFRecords of
  {uttid, transcript, frames}; returns the number written."""
  writers = [ops.host().TFRecordWriter(output_template % (i, n_shards))
             for i in range(n_shards)]
  n = 0
  for uttid, text, wav in utterances:
    feats = audio_lib.ExtractLogMelFeatures(wav, num_mel_bins)
    writers[n % n_shards].write(tf_example.MakeExample({
        'uttid': [uttid.encode()], 'transcript': [text.lower().encode()],
        'frames': feats.reshape(-1)}))
    n += 1
  for w in writers:
    w.close()
  return n


def _FromDir(input_dir, transcripts):
  with open(transcripts, encoding='utf-8') as f:
    for line in f:
      parts = line.strip().split(' ', 1)
      if len(parts) != 2:
        continue
      path = os.path.join(input_dir, parts[0] + '.wav')
      if os.path.exists(path):
        with open(path, 'rb') as wf:
          yield parts[0], parts[1], wf.read()


def main(argv):
  del argv
  n = WriteFeatures(_FromDir(FLAGS.input_dir, FLAGS.transcripts), FLAGS.output_template,
                    FLAGS.num_output_shards, FLAGS.num_mel_bins)
  print('wrote %d utterances to %d shards' % (n, FLAGS.num_output_shards))


if __name__ == '__main__':
  app.run(main)
