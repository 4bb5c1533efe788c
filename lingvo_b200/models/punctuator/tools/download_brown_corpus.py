"""Prepares the Brown corpus for the punctuator codelab (ref
`lingvo/tasks/punctuator/tools/download_brown_corpus.py`).

  python -m lingvo_b200.models.punctuator.tools.download_brown_corpus --outdir=/tmp/punctuator_data

Fetches `brown.zip` from the NLTK data mirror (or takes `--zip` from disk), strips the
part-of-speech tags (`word/tag` tokens), re-attaches punctuation to the preceding word,
and writes shuffled `train.txt` (80 %) / `test.txt` (20 %), one sentence per line — the
format `PunctuatorInput` reads.
"""

from __future__ import annotations

import argparse
import io
import os
import random
import sys
import urllib.request
import zipfile

URL = 'https://raw.githubusercontent.com/nltk/nltk_data/gh-pages/packages/corpora/brown.zip'
_NO_SPACE_BEFORE = {'.', ',', '!', '?', ';', ':', "''", ')', "'"}
_NO_SPACE_AFTER = {'``', '(', '`'}


def DetagSentence(line):
  """`The/at jury/nn said/vbd ./.` → `The jury said.`"""
  words = []
  for tok in line.split():
    word = tok.rsplit('/', 1)[0] if '/' in tok else tok
    if not word:
      continue
    words.append(word)
  out = ''
  prev = None
  for w in words:
    if out and w not in _NO_SPACE_BEFORE and prev not in _NO_SPACE_AFTER:
      out += ' '
    out += {'``': '"', "''": '"'}.get(w, w)
    prev = w
  return out.strip()


def ReadSentences(zf):
  """Sentences of all `brown/c???` files in the archive."""
  sents = []
  for name in sorted(zf.namelist()):
    base = os.path.basename(name)
    if len(base) != 4 or not base.startswith('c'):
      continue
    for line in io.TextIOWrapper(zf.open(name), encoding='utf-8', errors='replace'):
      s = DetagSentence(line)
      if len(s.split()) >= 2:
        sents.append(s)
  return sents


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--outdir', default='/tmp/punctuator_data')
  ap.add_argument('--zip', default='', help='use a local brown.zip instead of downloading')
  ap.add_argument('--seed', type=int, default=1234)
  a = ap.parse_args(argv)
  os.makedirs(a.outdir, exist_ok=True)
  path = a.zip
  if not path:
    path = os.path.join(a.outdir, 'brown.zip')
    if not os.path.exists(path):
      urllib.request.urlretrieve(URL, path)   # noqa: S310
  with zipfile.ZipFile(path) as zf:
    sents = ReadSentences(zf)
  random.Random(a.seed).shuffle(sents)
  n_train = int(len(sents) * 0.8)
  for fname, part in (('train.txt', sents[:n_train]), ('test.txt', sents[n_train:])):
    with open(os.path.join(a.outdir, fname), 'w', encoding='utf-8') as f:
      f.write('\n'.join(part) + '\n')
  print('%d train / %d test sentences in %s' % (n_train, len(sents) - n_train, a.outdir))
  return 0


if __name__ == '__main__':
  sys.exit(main())
