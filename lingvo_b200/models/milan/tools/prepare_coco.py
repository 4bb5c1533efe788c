"""COCO-captions → Milan TFRecords (ref `lingvo/tasks/milan/tools/prepare_coco.py`).

  python -m lingvo_b200.models.milan.tools.prepare_coco \
      --annotations=captions_train2017.json --image_dir=train2017 \
      --bert_vocab=vocab.txt --output=/data/coco/train --num_shards=64

Reads the official COCO caption annotations + JPEG directory (the reference pulls them from
TFDS), emits one `tf.train.Example` per (image, caption) in `common_schema` layout. Token
features: the reference runs a TF-Hub BERT; here `--bert_checkpoint` may point to an
in-repo `BertTransformer` checkpoint, otherwise deterministic hashed embeddings of the
word pieces are written (enough to exercise the pipeline end to end).
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.utils import tf_example


class HashedWordPieceEncoder:
  """Stand-in for the reference's `TfHubBertEncoder`: per-piece pseudo-random unit vectors
  (stable across runs), `[max_len, dim]` + length."""

  def __init__(self, tokenizer, max_len=48, dim=768):
    self._tok, self._max_len, self._dim = tokenizer, max_len, dim

  def __call__(self, caption):
    ids = self._tok(caption)[:self._max_len]
    out = np.zeros((self._max_len, self._dim), np.float32)
    for i, t in enumerate(ids):
      seed = int.from_bytes(hashlib.md5(str(t).encode()).digest()[:4], 'little')
      v = np.random.RandomState(seed).randn(self._dim).astype(np.float32)
      out[i] = v / np.linalg.norm(v)
    return out, len(ids)


def ReadCoco(annotations_path):
  """→ list of (image_id, file_name, caption_id, caption)."""
  with open(annotations_path, encoding='utf-8') as f:
    ann = json.load(f)
  files = {im['id']: im['file_name'] for im in ann['images']}
  return [(a['image_id'], files[a['image_id']], a['id'], a['caption'].strip())
          for a in ann['annotations'] if a['image_id'] in files]


def MakeExample(image_bytes, image_id, caption, caption_id, emb, length):
  return tf_example.MakeExample({
      'image/encoded': [image_bytes], 'image/id': np.asarray([image_id], np.int64),
      'text/captions': [caption.encode('utf-8')], 'text/id': np.asarray([caption_id], np.int64),
      'text/bert/lengths': np.asarray([length], np.int64),
      'text/bert/embeddings': emb.reshape(-1).astype(np.float32)})


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--annotations', required=True)
  ap.add_argument('--image_dir', required=True)
  ap.add_argument('--output', required=True, help='output file prefix')
  ap.add_argument('--num_shards', type=int, default=16)
  ap.add_argument('--bert_vocab', default='')
  ap.add_argument('--max_len', type=int, default=48)
  ap.add_argument('--dim', type=int, default=768)
  a = ap.parse_args(argv)
  if a.bert_vocab:
    from lingvo_b200.models.lm import tokenizer as lm_tokenizer
    tok = lm_tokenizer.BertTokenizer.Params().Set(vocab_filepath=a.bert_vocab).Instantiate()
    tokenize = tok.Encode
  else:
    tokenize = lambda s: [hash(w) & 0xFFFF for w in s.lower().split()]
  enc = HashedWordPieceEncoder(tokenize, a.max_len, a.dim)
  os.makedirs(os.path.dirname(a.output) or '.', exist_ok=True)
  writers = [ops.host().TFRecordWriter('%s-%05d-of-%05d' % (a.output, i, a.num_shards))
             for i in range(a.num_shards)]
  n = 0
  for image_id, fname, cap_id, caption in ReadCoco(a.annotations):
    path = os.path.join(a.image_dir, fname)
    if not os.path.exists(path):
      continue
    with open(path, 'rb') as f:
      img = f.read()
    emb, length = enc(caption)
    writers[n % a.num_shards].write(MakeExample(img, image_id, caption, cap_id, emb, length))
    n += 1
  for w in writers:
    w.close()
  print('wrote %d examples to %s-*' % (n, a.output))
  return 0


if __name__ == '__main__':
  sys.exit(main())
