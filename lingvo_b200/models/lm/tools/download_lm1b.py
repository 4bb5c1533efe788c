"""Fetches and unpacks the One Billion Word benchmark (ref
`lingvo/tasks/lm/tools/download_lm1b.py`).

  python -m lingvo_b200.models.lm.tools.download_lm1b --outdir=/tmp/lm1b

Downloads `1-billion-word-language-modeling-benchmark-r13output.tar.gz` (or uses
`--tarball` if already on disk — e.g. on an air-gapped box), extracts the tokenised
training / held-out shards and writes `vocab.txt` (word count ≥ `--min_count`, most
frequent first, with `<S> </S> <UNK>` on top) next to them.
"""

from __future__ import annotations

import argparse
import collections
import glob
import os
import sys
import tarfile
import urllib.request

URL = ('http://www.statmt.org/lm-benchmark/'
       '1-billion-word-language-modeling-benchmark-r13output.tar.gz')
ROOT = '1-billion-word-language-modeling-benchmark-r13output'


def Download(url, path):
  if os.path.exists(path):
    return path
  tmp = path + '.part'
  with urllib.request.urlopen(url) as r, open(tmp, 'wb') as f:   # noqa: S310
    while True:
      chunk = r.read(1 << 20)
      if not chunk:
        break
      f.write(chunk)
  os.replace(tmp, path)
  return path


def Extract(tarball, outdir):
  with tarfile.open(tarball) as tf:
    members = [m for m in tf.getmembers() if 'tokenized.shuffled' in m.name]
    tf.extractall(outdir, members=members)   # noqa: S202
  return os.path.join(outdir, ROOT)


def BuildVocab(corpus_dir, out_path, min_count=3):
  counts = collections.Counter()
  pattern = os.path.join(corpus_dir, 'training-monolingual.tokenized.shuffled', 'news.en-*')
  for f in sorted(glob.glob(pattern)):
    with open(f, encoding='utf-8') as fh:
      for line in fh:
        counts.update(line.split())
  with open(out_path, 'w', encoding='utf-8') as fh:
    for w in ('<S>', '</S>', '<UNK>'):
      fh.write(w + '\n')
    for w, c in counts.most_common():
      if c < min_count:
        break
      fh.write(w + '\n')
  return len(counts)


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--outdir', default='/tmp/lm1b')
  ap.add_argument('--tarball', default='')
  ap.add_argument('--min_count', type=int, default=3)
  a = ap.parse_args(argv)
  os.makedirs(a.outdir, exist_ok=True)
  tarball = a.tarball or Download(URL, os.path.join(a.outdir, os.path.basename(URL)))
  corpus = Extract(tarball, a.outdir)
  n = BuildVocab(corpus, os.path.join(corpus, 'vocab.txt'), a.min_count)
  print('extracted to %s (%d distinct words)' % (corpus, n))
  return 0


if __name__ == '__main__':
  sys.exit(main())
