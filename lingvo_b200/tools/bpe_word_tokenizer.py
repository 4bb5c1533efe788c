"""Applies BPE merge rules to a word list (ref `lingvo/tools/bpe_word_tokenizer.py`):
reads words from --input, writes `word<TAB>piece piece …` lines."""
from absl import app
from absl import flags

from lingvo_b200 import ops

flags.DEFINE_string('input', '', 'One word per line.')
flags.DEFINE_string('codes_file', '', 'BPE merge rules.')
flags.DEFINE_string('vocab_file', '', 'BPE vocabulary.')
flags.DEFINE_string('output', '', 'Output path.')
FLAGS = flags.FLAGS


def main(argv):
  del argv
  bpe = ops.host().BpeTokenizer(FLAGS.codes_file, FLAGS.vocab_file)
  with open(FLAGS.input, encoding='utf-8') as fi, open(FLAGS.output, 'w', encoding='utf-8') as fo:
    for line in fi:
      w = line.strip()
      if w:
        fo.write('%s\t%s\n' % (w, ' '.join(bpe.encode_word(w))))


if __name__ == '__main__':
  app.run(main)
