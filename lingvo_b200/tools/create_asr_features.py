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


def main(argv):
  del argv
  n_shards = FLAGS.num_output_shards
  writers = [ops.host().TFRecordWriter(FLAGS.output_template % (i, n_shards))
             for i in range(n_shards)]
  n = 0
  with open(FLAGS.transcripts, encoding='utf-8') as f:
    for line in f:
      parts = line.strip().split(' ', 1)
      if len(parts) != 2:
        continue
      uttid, text = parts
      path = os.path.join(FLAGS.input_dir, uttid + '.wav')
      if not os.path.exists(path):
        continue
      with open(path, 'rb') as wf:
        feats = audio_lib.ExtractLogMelFeatures(wf.read(), FLAGS.num_mel_bins)
      writers[n % n_shards].write(tf_example.MakeExample({
          'uttid': [uttid.encode()], 'transcript': [text.lower().encode()],
          'frames': feats.reshape(-1)}))
      n += 1
  for w in writers:
    w.close()
  print('wrote %d utterances to %d shards' % (n, n_shards))


if __name__ == '__main__':
  app.run(main)
