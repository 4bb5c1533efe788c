"""WMT data preparation (ref `lingvo/tasks/mt/tools/wmt14.*.sh`, `wmtm16.*.sh`,
`wmt14_lib.sh`): download → unpack → tokenise → word-piece encode → TFRecords, as one
restartable Python driver instead of nine shell stages.

  python -m lingvo_b200.models.mt.tools.wmt_get_data --dataset=wmt14 --root=/tmp/wmt14 \
      [--steps=download,unpack,tokenize,encode] [--wpm_vocab=/path/wpm-ende.voc]

Every step writes a `<step>.done` marker; finished steps are skipped on re-run. Downloads
need network access (`--mirror` may point at a local directory holding the tarballs).
"""

from __future__ import annotations

import argparse
import glob
import gzip
import os
import re
import shutil
import sys
import tarfile
import urllib.request

import numpy as np

DATASETS = {
    'wmt14': dict(
        train=['http://www.statmt.org/wmt13/training-parallel-europarl-v7.tgz',
               'http://www.statmt.org/wmt13/training-parallel-commoncrawl.tgz',
               'http://www.statmt.org/wmt14/training-parallel-nc-v9.tgz'],
        devtest=['http://www.statmt.org/wmt14/dev.tgz', 'http://www.statmt.org/wmt14/test-full.tgz'],
        train_pairs=[('training/europarl-v7.de-en.en', 'training/europarl-v7.de-en.de'),
                     ('commoncrawl.de-en.en', 'commoncrawl.de-en.de'),
                     ('training/news-commentary-v9.de-en.en',
                      'training/news-commentary-v9.de-en.de')],
        dev_pairs=[('dev/newstest2013.en', 'dev/newstest2013.de')],
        test_pairs=[('test-full/newstest2014-deen-src.en.sgm',
                     'test-full/newstest2014-deen-ref.de.sgm')],
        vocab='wpm-ende.voc', shards=36),
    'wmtm16': dict(
        train=['http://www.quest.dcs.shef.ac.uk/wmt16_files_mmt/training.tar.gz'],
        devtest=['http://www.quest.dcs.shef.ac.uk/wmt16_files_mmt/validation.tar.gz',
                 'http://www.quest.dcs.shef.ac.uk/wmt16_files_mmt/mmt16_task1_test.tar.gz'],
        train_pairs=[('train.en', 'train.de')], dev_pairs=[('val.en', 'val.de')],
        test_pairs=[('test.en', 'test.de')], vocab='wpm-ende-2k.voc', shards=1),
}


def _Done(root, step):
  return os.path.join(root, step + '.done')


def _Fetch(url, dst, mirror):
  if os.path.exists(dst):
    return
  name = os.path.basename(url)
  if mirror and os.path.exists(os.path.join(mirror, name)):
    shutil.copy(os.path.join(mirror, name), dst)
    return
  tmp = dst + '.part'
  with urllib.request.urlopen(url) as r, open(tmp, 'wb') as f:   # noqa: S310
    shutil.copyfileobj(r, f, 1 << 20)
  os.replace(tmp, dst)


def Download(root, spec, mirror=None):
  os.makedirs(os.path.join(root, 'raw'), exist_ok=True)
  for url in spec['train'] + spec['devtest']:
    _Fetch(url, os.path.join(root, 'raw', os.path.basename(url)), mirror)


def Unpack(root, spec):
  del spec
  out = os.path.join(root, 'unpacked')
  os.makedirs(out, exist_ok=True)
  for f in sorted(glob.glob(os.path.join(root, 'raw', '*'))):
    if f.endswith(('.tgz', '.tar.gz')):
      with tarfile.open(f) as t:
        t.extractall(out)   # noqa: S202
    elif f.endswith('.gz'):
      with gzip.open(f, 'rb') as src, open(os.path.join(out, os.path.basename(f)[:-3]), 'wb') as d:
        shutil.copyfileobj(src, d)


_SGM = re.compile(r'<seg id="\d+">(.*)</seg>')
_PUNCT = re.compile(r'([.,!?;:()"„“”»«])')


def _ReadLines(path):
  with open(path, encoding='utf-8', errors='replace') as f:
    if path.endswith('.sgm'):
      return [m.group(1).strip() for m in (_SGM.search(l) for l in f) if m]
    return [l.rstrip('\n') for l in f]


def TokenizeLine(line):
  """Moses-style light tokenisation: split punctuation, squeeze blanks, normalise quotes."""
  line = line.replace('“', '"').replace('”', '"').replace('„', '"').replace('’', "'")
  return ' '.join(_PUNCT.sub(r' \1 ', line).split())


def Tokenize(root, spec):
  out = os.path.join(root, 'tok')
  os.makedirs(out, exist_ok=True)
  for split in ('train', 'dev', 'test'):
    src_all, tgt_all = [], []
    for s, t in spec[split + '_pairs']:
      sp, tp = (os.path.join(root, 'unpacked', x) for x in (s, t))
      if not (os.path.exists(sp) and os.path.exists(tp)):
        print('skipping missing pair', s, t)
        continue
      a, b = _ReadLines(sp), _ReadLines(tp)
      n = min(len(a), len(b))
      src_all += [TokenizeLine(x) for x in a[:n]]
      tgt_all += [TokenizeLine(x) for x in b[:n]]
    keep = [(a, b) for a, b in zip(src_all, tgt_all) if a and b and
            len(a.split()) <= 200 and len(b.split()) <= 200]
    with open(os.path.join(out, split + '.tsv'), 'w', encoding='utf-8') as f:
      for a, b in keep:
        f.write('%s\t%s\n' % (a, b))
    print('%s: %d sentence pairs' % (split, len(keep)))


def Encode(root, spec, wpm_vocab, max_len=200):
  """Word-piece encode the TSVs into `NmtInput` tf.Example records."""
  from lingvo_b200 import ops  # pylint: disable=g-import-not-at-top
  from lingvo_b200.core import wpm_encoder  # pylint: disable=g-import-not-at-top
  from lingvo_b200.utils import tf_example  # pylint: disable=g-import-not-at-top
  enc = wpm_encoder.WpmEncoder(wpm_vocab)
  out = os.path.join(root, 'wpm')
  os.makedirs(out, exist_ok=True)
  for split in ('train', 'dev', 'test'):
    tsv = os.path.join(root, 'tok', split + '.tsv')
    if not os.path.exists(tsv):
      continue
    shards = spec['shards'] if split == 'train' else 1
    name = lambda i: os.path.join(out, '%s.tfrecords-%05d-of-%05d' % (split, i, shards)) \
        if shards > 1 else os.path.join(out, split + '.tfrecords')
    writers = [ops.host().TFRecordWriter(name(i)) for i in range(shards)]
    with open(tsv, encoding='utf-8') as f:
      for n, line in enumerate(f):
        src, tgt = line.rstrip('\n').split('\t')
        s = list(enc.EncodeToStringAndIds(src)[1])[:max_len - 1] + [enc.sentence_end_id]
        t = list(enc.EncodeToStringAndIds(tgt)[1])[:max_len - 1]
        ex = tf_example.MakeExample({
            'source_id': np.asarray(s), 'source_padding': np.zeros(len(s), np.float32),
            'target_id': np.asarray([enc.sentence_start_id] + t),
            'target_padding': np.zeros(len(t) + 1, np.float32),
            'target_label': np.asarray(t + [enc.sentence_end_id]),
            'target_weight': np.ones(len(t) + 1, np.float32)})
        writers[n % shards].write(ex)
    for w in writers:
      w.close()


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--dataset', default='wmt14', choices=sorted(DATASETS))
  ap.add_argument('--root', required=True)
  ap.add_argument('--steps', default='download,unpack,tokenize,encode')
  ap.add_argument('--mirror', default='')
  ap.add_argument('--wpm_vocab', default='')
  a = ap.parse_args(argv)
  spec = DATASETS[a.dataset]
  os.makedirs(a.root, exist_ok=True)
  for step in a.steps.split(','):
    if os.path.exists(_Done(a.root, step)):
      print('[skip] %s' % step)
      continue
    print('[run ] %s' % step)
    if step == 'download':
      Download(a.root, spec, a.mirror)
    elif step == 'unpack':
      Unpack(a.root, spec)
    elif step == 'tokenize':
      Tokenize(a.root, spec)
    elif step == 'encode':
      vocab = a.wpm_vocab or os.path.join(a.root, spec['vocab'])
      Encode(a.root, spec, vocab)
    else:
      raise ValueError('unknown step ' + step)
    open(_Done(a.root, step), 'w').close()
  return 0


if __name__ == '__main__':
  sys.exit(main())
