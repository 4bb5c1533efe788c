"""Native host library (`_H.so`) tests: records, yielders, batcher, tokenizers,
packing, MASS, best_step."""

import collections
import os

import numpy as np
import pytest

from lingvo_b200 import ops
from lingvo_b200.core import generic_input
from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.ops import host_ops
from lingvo_b200.utils import tfrecord


@pytest.fixture(scope='module')
def shards(tmp_path_factory):
  d = tmp_path_factory.mktemp('recs')
  h = ops.host()
  n = 0
  for s in range(3):
    w = h.TFRecordWriter(str(d / ('data-%05d' % s)))
    for _ in range(40):
      w.write(('rec%04d' % n).encode() + b'x' * (n % 17))
      n += 1
    w.close()
  return str(d / 'data-*'), n


def test_crc_and_python_reader_interop(shards):
  pattern, n = shards
  h = ops.host()
  assert h.crc32c(b'123456789') == 0xE3069283
  # files written by the native writer are readable by the pure-python reader
  files = h.glob_files(pattern)
  assert len(files) == 3
  recs = list(tfrecord.ReadRecords(files[0])) if hasattr(tfrecord, 'ReadRecords') else \
      list(tfrecord.TFRecordReader(files[0]))
  assert len(recs) == 40 and recs[0].startswith(b'rec0000')


def test_basic_yielder_epochs_and_shuffle(shards):
  pattern, n = shards
  h = ops.host()
  y = h.basic_record_yielder('tfrecord:' + pattern, seed=7, bufsize=32, parallelism=2,
                             num_epochs=2)
  seen = []
  while True:
    r = y.next()
    if r is None:
      break
    seen.append(r[0][:7])
  assert len(seen) == 2 * n
  first, second = seen[:n], seen[n:]
  assert sorted(first) == sorted(second) == sorted(('rec%04d' % i).encode() for i in range(n))
  assert first != sorted(first)           # shuffled


def test_sequential_and_mix_yielders(shards):
  pattern, n = shards
  h = ops.host()
  y = h.sequential_record_yielder('tfrecord:' + pattern, repeat_count=1)
  recs = []
  while True:
    r = y.next()
    if r is None:
      break
    recs.append(r[0][:7])
  assert recs == [('rec%04d' % i).encode() for i in range(n)]
  mix = h.weighted_mix_record_yielder(
      [h.basic_record_yielder('iota:1000000', seed=1),
       h.basic_record_yielder('text:' + __file__, seed=2)], [0.8, 0.2], seed=3)
  src = collections.Counter(mix.next()[1] for _ in range(500))
  assert 330 < src[0] < 470
  mix.close()


def test_generic_input_buckets_and_pads(shards):
  pattern, n = shards

  def proc(record):
    k = len(record)
    if k % 5 == 0:
      return None                                        # filtered
    ids = np.frombuffer(record, np.uint8).astype(np.int32)
    return NestedMap(ids=ids, length=np.int32(k)), k

  gi = generic_input.GenericInput(
      proc, file_pattern='tfrecord:' + pattern, bucket_upper_bound=[12, 18, 22],
      bucket_batch_limit=[8, 4, 2], file_random_seed=1, file_buffer_size=16,
      repeat_count=1, num_threads=3)
  total = 0
  for batch, keys in gi:
    m = batch.ids.shape[0]
    total += m
    assert batch.ids.shape[1] == keys.max()
    bucket = 0 if keys.max() <= 12 else (1 if keys.max() <= 18 else 2)
    assert m <= [8, 4, 2][bucket]
    for row, k in zip(batch.ids, keys):
      assert (row[k:] == 0).all() and row[k - 1] != 0
    assert (batch.length == keys).all()
  expect = sum(1 for i in range(n) if (7 + i % 17) % 5 != 0 and (7 + i % 17) <= 22)
  assert total == expect
  assert gi.records_skipped == sum(1 for i in range(n)
                                   if (7 + i % 17) % 5 != 0 and (7 + i % 17) > 22)
  gi.Close()


def test_ascii_and_vocab_and_bpe_tokenizers(tmp_path):
  h = ops.host()
  ids = h.ascii_to_ids('Hello <noise> w0rld!')
  assert ids[:5] == [12, 9, 16, 16, 19] and 4 in ids
  assert h.ascii_to_string(ids) == 'hello <noise> w0rld!'
  tok = tokenizers.AsciiTokenizer.Params().Instantiate()
  i, l, p = tok.StringsToIds(['ab', 'hello'], 6)
  assert i[0].tolist() == [1, 5, 6, 2, 2, 2] and l[0].tolist() == [5, 6, 2, 2, 2, 2]
  assert p[0].tolist() == [0, 0, 0, 1, 1, 1] and p[1].tolist() == [0] * 6
  assert tok.IdsToStrings(l, [2, 5]) == ['ab', 'hello']
  vocab = tmp_path / 'vocab.txt'
  vocab.write_text('<unk>\t0\n<s>\t1\n</s>\t2\nthe\t3\ncat\t4\n')
  vt = tokenizers.VocabFileTokenizer.Params().Set(
      token_vocab_filepath=str(vocab)).Instantiate()
  i, l, p = vt.StringsToIds(['the dog cat'], 5)
  assert l[0].tolist() == [3, 0, 4, 2, 2]
  codes = tmp_path / 'codes.txt'
  codes.write_text('l o\nlo w</w>\ne r</w>\n')
  bvocab = tmp_path / 'bpe_vocab.txt'
  bvocab.write_text('<unk>\nlow</w>\nlo\nw\ner</w>\nn\ne\n')
  bt = h.BpeTokenizer(str(codes), str(bvocab))
  assert bt.encode_word('low') == ['low</w>']
  assert bt.encode_word('lower') == ['lo', 'w', 'er</w>']
  assert bt.to_string(bt.to_ids('low lower')) == 'low lower'


def test_pack_sequences_and_apply():
  src = [3, 2, 4, 1, 9]
  tgt = [2, 2, 1, 1, 1]
  out = host_ops.PackSequences(src, tgt, packed_batch_size=3, packed_src_seq_len=6,
                               packed_tgt_seq_len=5, seed=1)
  s_seg, s_pos, s_idx, t_seg, t_pos, t_idx = out
  assert s_seg.shape == (3, 6) and t_seg.shape == (3, 5)
  # input 4 (len 9) does not fit and is dropped; every other input appears exactly once
  used = set()
  for r in range(3):
    for seg in set(s_seg[r]) - {0}:
      items = set(s_idx[r][s_seg[r] == seg])
      assert len(items) == 1
      i = items.pop()
      assert i not in used
      used.add(i)
      assert (s_seg[r] == seg).sum() == src[i]
      assert s_pos[r][s_seg[r] == seg].tolist() == list(range(src[i]))
      assert (t_seg[r] == seg).sum() == tgt[i]
  assert used == {0, 1, 2, 3}
  data = np.arange(5 * 9).reshape(5, 9)
  packed = host_ops.ApplyPacking(data, -1, s_seg, s_idx)
  for r in range(3):
    for c in range(6):
      if s_seg[r, c]:
        assert packed[r, c] == data[s_idx[r, c], s_pos[r, c]]
      else:
        assert packed[r, c] == -1


def test_pack_single_sequence():
  g = host_ops.PackSingleSequence([5, 4, 3, 8, 2, 20], 10, False)
  assert g[5] == -1
  loads = collections.Counter()
  for gi, ln in zip(g, [5, 4, 3, 8, 2, 20]):
    if gi >= 0:
      loads[gi] += ln
  assert all(v <= 10 for v in loads.values()) and len(loads) == 3
  seq = host_ops.PackSingleSequence([5, 4, 3, 8, 2], 10, True)
  assert seq.tolist() == [0, 0, 1, 2, 2]


def test_mass_and_best_step(tmp_path):
  ids = np.arange(4, 24, dtype=np.int32).reshape(2, 10)
  src, tgt, lab, w = host_ops.Mass(ids, np.ones((2, 10), np.float32), [10, 6], mask_id=3,
                                   mask_ratio=0.5, vocab_size=100, seed=5)
  assert (lab == ids).all()
  assert w[0].sum() == 5 and w[1].sum() == 3 and w[1, 6:].sum() == 0
  masked = w[0] > 0
  assert (src[0][~masked] == ids[0][~masked]).all()
  assert (tgt[0][~masked] == 3).all()
  hist = tmp_path / 'h.txt'
  hist.write_text('100 3.0\n200 2.5\n300 2.6\n400 2.49\n')
  assert host_ops.BestStep(str(hist), 0.0, True) == (400, 400)
  assert host_ops.BestStep(str(hist), 0.05, True) == (200, 400)
