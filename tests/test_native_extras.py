"""Native host ops added for inventory parity: MLPerf sub-words, n-gram vocab, static maps,
apply_packing, 2-D AP, point sampling, preconditioner captain."""

import os

import numpy as np
import pytest
import torch

from lingvo_b200 import ops
from lingvo_b200.ops import host_ops


def test_mlperf_subword_decode(tmp_path):
  toks = ['<pad>_', '<EOS>_', 'Hello_', 'wor', 'ld_', ',_', '!_', 'über', 'all_', '日本_', '語_']
  f = tmp_path / 'vocab.subwords'
  f.write_text('\n'.join("'%s'" % t for t in toks) + '\n')
  ids = np.array([[2, 3, 4, 5, 2, 6, 0], [7, 8, 9, 10, 0, 0, 0]], np.int32)
  out = host_ops.MlPerfSubwordIdToString(ids, [6, 4], str(f))
  # blanks only between two tokens that both start alphanumerically
  assert out[0] == 'Hello world,Hello!'
  assert out[1] == 'überall 日本 語'
  with pytest.raises(IndexError):
    host_ops.MlPerfSubwordIdToString(np.array([[99]], np.int32), [1], str(f))
  # the tokenizer layer on top of the native decoder (ref core/ml_perf_tokenizer.py)
  from lingvo_b200.core import ml_perf_tokenizer
  tok = ml_perf_tokenizer.MlPerfTokenizer.Params().Set(vocab_filepath=str(f)).Instantiate()
  assert tok.IdsToStrings(torch.as_tensor(ids), torch.tensor([6, 4])) == out
  with pytest.raises(NotImplementedError):
    tok.StringsToIds(['x'], 4)


def test_ngram_id_to_token_and_token_in_vocab(tmp_path):
  f = tmp_path / 'ngrams.txt'
  f.write_text('<unk>\n<s>\n</s>\nth\ne\nqu\nick\n')
  out = host_ops.NgramIdToToken(np.array([[3, 4, 0], [5, 6, 4]], np.int32), [2, 3], str(f))
  assert out == ['the', 'quicke']
  out = host_ops.NgramIdToToken(np.array([[3, 4, 0]], np.int32), [3], str(f), ngram_separator='|')
  assert out == ['th|e|<unk>']
  assert host_ops.TokenInVocab('qu', str(f)) and not host_ops.TokenInVocab('zz', str(f))
  np.testing.assert_array_equal(host_ops.TokenInVocab(['e', 'x'], str(f)), [True, False])


def test_str_to_vocab_tokens(tmp_path):
  f = tmp_path / 'v.txt'
  f.write_text('<unk>\n<s>\n</s>\nhello\nworld\n')
  ids, tgt, pad = host_ops.StrToVocabTokens(['hello world', 'world foo hello'], str(f), maxlen=5,
                                            load_token_ids_from_vocab=False)
  np.testing.assert_array_equal(tgt[0], [3, 4, 2, 2, 2])
  np.testing.assert_array_equal(ids[0, :3], [1, 3, 4])
  np.testing.assert_array_equal(pad[0], [0, 0, 0, 1, 1])
  np.testing.assert_array_equal(tgt[1, :4], [4, 0, 3, 2])
  assert pad[1].sum() == 1


def test_static_maps():
  m = host_ops.StaticMap(['a', 'b', 'c'])
  np.testing.assert_array_equal(m.Lookup([['c', 'a'], ['zz', b'b']]), [[2, 0], [-1, 1]])
  m = host_ops.StaticMap([10, 20], ['x', 'y'], unk='?')
  assert m.Lookup([20, 5, 10]).tolist() == ['y', '?', 'x']
  m = host_ops.StaticMap([7, 8], [70, 80], unk=0)
  np.testing.assert_array_equal(m.Lookup(np.array([8, 9, 7])), [80, 0, 70])
  with pytest.raises(ValueError):
    host_ops.StaticMap(['a', 'a'])


def test_apply_packing_native_matches_manual():
  src = np.arange(4 * 5 * 2, dtype=np.float32).reshape(4, 5, 2)
  lens = np.array([2, 3, 5, 1], np.int32)
  seg, pos, idx, *_ = host_ops.PackSequences(lens, lens, 3, 5, 5, seed=1)
  out = host_ops.ApplyPacking(src, -1.0, seg, idx)
  assert out.shape == (3, 5, 2)
  for r in range(3):
    for c in range(5):
      if seg[r, c] == 0:
        assert (out[r, c] == -1).all()
      else:
        np.testing.assert_array_equal(out[r, c], src[idx[r, c], pos[r, c]])
  ints = host_ops.ApplyPacking(np.arange(20, dtype=np.int64).reshape(4, 5), 0, seg, idx)
  assert ints.dtype == np.int64 and ints.shape == (3, 5)


def test_average_precision_2d():
  from lingvo_b200.models.car import ops as car_ops
  gt = np.array([[0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 10, 10]], np.float32)
  gt_img = np.array([0, 0, 1], np.int32)
  pd = np.array([[0, 0, 10, 9], [21, 21, 30, 30], [50, 50, 60, 60], [0, 0, 10, 10]], np.float32)
  pd_img = np.array([0, 0, 0, 1], np.int32)
  score = np.array([0.9, 0.8, 0.7, 0.6], np.float32)
  ap, pr, sh = car_ops.average_precision2d(0.5, gt, gt_img, np.zeros(3, np.int32), pd, pd_img,
                                           np.zeros(4, np.int32), score, num_recall_points=4)
  np.testing.assert_array_equal(sh[:, 1].numpy(), [1, 1, 0, 1])
  # TP TP FP TP → interpolated precision 1, 1, .75 at recalls 1/3, 2/3, 1
  assert abs(ap - (1 / 3 + 1 / 3 + 0.75 / 3)) < 1e-5
  assert pr.shape == (4, 2)


def _Scene(n=400, seed=0):
  rng = np.random.RandomState(seed)
  pts = rng.uniform(-10, 10, size=(2, n, 4)).astype(np.float32)
  pad = np.zeros((2, n), np.float32)
  pad[1, n // 2:] = 1.0
  return pts, pad


@pytest.mark.parametrize('algo', ['auto', 'hash'])
def test_sample_points_closest_matches_bruteforce(algo):
  from lingvo_b200.models.car import ops as car_ops
  pts, pad = _Scene()
  c, cp, idx, ip = car_ops.sample_points(pts, pad, 16, 8, max_distance=4.0, random_seed=3,
                                         neighbor_algorithm=algo)
  assert c.shape == (2, 16) and idx.shape == (2, 16, 8)
  assert cp.sum() == 0
  for b in range(2):
    valid = np.nonzero(pad[b] < 0.5)[0]
    assert len(set(c[b].tolist())) == 16 and set(c[b].tolist()) <= set(valid.tolist())
    for j in range(16):
      d = np.linalg.norm(pts[b, valid, :3] - pts[b, c[b, j], :3], axis=1)
      order = valid[np.argsort(d, kind='stable')]
      within = int((d <= 4.0).sum())
      k = min(8, within)
      got = idx[b, j, :k].numpy()
      np.testing.assert_allclose(
          np.sort(np.linalg.norm(pts[b, got, :3] - pts[b, c[b, j], :3], axis=1)),
          np.sort(d)[:k], rtol=1e-5)
      assert ip[b, j, :k].sum() == 0 and ip[b, j, k:].sum() == 8 - k
      assert order[0] == c[b, j]            # the centre is its own nearest neighbour


def test_sample_points_farthest_spreads_and_respects_z_and_seeds():
  from lingvo_b200.models.car import ops as car_ops
  pts, pad = _Scene(seed=1)
  c, cp, idx, ip = car_ops.sample_points(pts, pad, 8, 4, center_z_min=0.0, center_z_max=5.0,
                                         random_seed=1, num_seeded_points=10,
                                         neighbor_sampler='uniform', max_distance=6.0)
  for b in range(2):
    z = pts[b, c[b], 2]
    assert (z >= 0).all() and (z <= 5).all()
    assert (c[b] >= 10).all() and (idx[b][ip[b] < 0.5] >= 10).all()
    d = np.linalg.norm(pts[b, idx[b], :3] - pts[b, c[b], None, :3], axis=-1)
    assert (d[ip[b].numpy() < 0.5] <= 6.0 + 1e-5).all()
  # farthest-point centres are better spread than uniform ones
  def MinSep(sel):
    x = pts[0, sel[0], :3]
    dm = np.linalg.norm(x[:, None] - x[None], axis=-1) + np.eye(len(x)) * 1e9
    return dm.min()
  cf, *_ = car_ops.sample_points(pts, pad, 32, 1, random_seed=5)
  cu, *_ = car_ops.sample_points(pts, pad, 32, 1, random_seed=5, center_selector='uniform')
  assert MinSep(cf) > MinSep(cu)
  # determinism for a fixed seed, padding when the scene is too small
  c2, *_ = car_ops.sample_points(pts, pad, 32, 1, random_seed=5)
  assert torch.equal(cf, c2)
  small_pad = np.ones((1, 50), np.float32); small_pad[0, :3] = 0
  c3, cp3, _, ip3 = car_ops.sample_points(pts[:1, :50], small_pad, 5, 2, random_seed=0)
  assert cp3[0].tolist() == [0, 0, 0, 1, 1] and ip3[0, 3:].sum() == 4


def test_preconditioner_captain_cpu_async_and_sync():
  from lingvo_b200.core import preconditioner_captain as pc
  cap = pc.PreconditionerCaptain(num_compute_threads=2)
  a = torch.randn(6, 6, dtype=torch.float64)
  stat = (a @ a.t() + 0.5 * torch.eye(6, dtype=torch.float64)).float()
  t, ok = cap.GetPreconditioner('w/L')
  assert t is None and not ok
  cap.InsertGradientStatistics('w/L', stat, 4, global_step=7)
  cap.WaitAll()
  t, ok = cap.GetPreconditioner('w/L')
  assert ok and cap.StatisticsStep('w/L') == 7
  # (A^{-1/4})^4 A ≈ I
  prod = torch.linalg.matrix_power(t.double(), 4) @ stat.double()
  assert (prod - torch.eye(6, dtype=torch.float64)).abs().max() < 5e-2
  cap.InsertGradientStatistics('w/R', stat, 2, global_step=8, sync=True)
  t2, ok2 = cap.GetPreconditioner('w/R')
  assert ok2 and ((t2.double() @ t2.double() @ stat.double()) - torch.eye(6, dtype=torch.float64)).abs().max() < 5e-2
  outs, oks = pc.GetPreconditioners([(3, 3)], ['missing'])
  assert not oks[0] and outs[0].shape == (3, 3)


def test_inverse_root_no_sync_matches_reference_iteration():
  from lingvo_b200.core import matrix_functions as mf
  a = torch.randn(12, 12)
  stat = a @ a.t() + 0.1 * torch.eye(12)
  x = mf.inlined_matrix_inverse_pth_root(stat, 4)
  y = mf.inverse_pth_root_no_sync(stat, 4, iter_count=100)
  torch.testing.assert_close(x, y, atol=2e-3, rtol=2e-2)


def test_shampoo_async_preconditioning_trains():
  from lingvo_b200.core import optimizer, py_utils
  torch.manual_seed(0)
  w = torch.nn.Parameter(torch.randn(8, 6))
  w.var_name = 'layer/w/var'
  target = torch.randn(8, 6)
  opt = optimizer.DistributedShampoo.Params().Set(
      name='sh', momentum=0.0, start_preconditioning_steps=2, preconditioning_compute_steps=2,
      async_preconditioning=True).Instantiate()
  losses = []
  for _ in range(60):
    loss = (w - target).square().sum()
    g, = torch.autograd.grad(loss, w)
    opt.Apply(0.3, [py_utils.VarGrad(w, g)])
    losses.append(float(loss))
  assert losses[-1] < 0.2 * losses[0]


def _WriteGz(path, recs):
  import gzip, struct
  h = ops.host()
  with gzip.open(path, 'wb') as f:
    for r in recs:
      hdr = struct.pack('<Q', len(r))
      f.write(hdr + struct.pack('<I', h.masked_crc32c(hdr)) + r + struct.pack('<I', h.masked_crc32c(r)))


def _Drain(y):
  out = []
  while True:
    r = y.next()
    if r is None:
      return out
    out.append(r[0])


def test_gzip_tfrecords_shard_specs_and_indirect_filesets(tmp_path):
  h = ops.host()
  _WriteGz(str(tmp_path / 'z.gz'), [b'alpha', b'beta' * 1000, b''])
  y = h.sequential_record_yielder('tfrecord_gzip:' + str(tmp_path / 'z.gz'), 1)
  assert _Drain(y) == [b'alpha', b'beta' * 1000, b'']
  # sharded spec: data@2 → data-?????-of-00002
  for i in range(2):
    w = h.TFRecordWriter(str(tmp_path / ('data-%05d-of-00002' % i)))
    w.write(b'r%d' % i); w.close()
  y = h.sequential_record_yielder('tfrecord:' + str(tmp_path / 'data@2'), 1)
  assert sorted(_Drain(y)) == [b'r0', b'r1']
  y = h.sequential_record_yielder('tfrecord:' + str(tmp_path / 'data@*'), 1)
  assert len(_Drain(y)) == 2
  # text_indirect: a VersionedFileSet text proto next to the data
  (tmp_path / 'v1.txt').write_text('one\ntwo\n')
  (tmp_path / 'v0.txt').write_text('old\n')
  (tmp_path / 'ckpt').write_text(
      'current {\n  file_pattern: "v1.txt"\n  create_timestamp: 12.5\n}\n'
      'history { file_pattern: "v0.txt" }\n')
  y = h.sequential_record_yielder('text_indirect:' + str(tmp_path / 'ckpt'), 1)
  assert _Drain(y) == [b'one', b'two']


def test_record_batcher_wait_stats(tmp_path):
  h = ops.host()
  y = h.sequential_record_yielder('iota:50', 1)
  def Proc(rec, source_id):
    v = int(rec)
    return v % 7, [np.array([v], np.int64)]
  b = h.RecordBatcher(y, Proc, [10], [8], num_threads=2)
  n = 0
  while True:
    try:
      out = b.get_next()
    except StopIteration:
      break
    n += out[1][0].shape[0] if isinstance(out[1], (list, tuple)) else 1
  st = b.stats()
  assert st['records_processed'] == 50 and st['records_skipped'] == 0
  assert st['consumer_wait_s'] >= 0 and st['producer_wait_s'] >= 0 and isinstance(st['hint'], str)
  b.close()


def test_hypothesis_proto_and_hyp_ops():
  from lingvo_b200.core import hyps as hyps_lib
  h = hyps_lib.Hypothesis(beam_id=3, ids=[5, 9, 2], scores=[-0.5, -1.25, -0.125],
                          atten_vecs=[[0.5, 0.5], [1.0, 0.0], [0.25, 0.75]], normalized_score=-0.75)
  back = hyps_lib.Hypothesis.FromString(h.SerializeToString())
  assert back == h
  # hand-encoded bytes: field 1 varint 3; field 2 packed [5, 9, 2]
  assert h.SerializeToString().startswith(b'\x08\x03\x12\x03\x05\x09\x02')
  # two beams × two hyps, three steps; slot 1 terminates at step 2 with path 0 ← 1 ← 1
  hyps = np.array([[11, 12, 13, 14], [21, 22, 23, 24], [31, 32, 33, 34]], np.int32)
  prev = np.array([[0, 1, 2, 3], [1, 0, 2, 3], [0, 1, 2, 3]], np.int32)
  done = np.zeros((3, 4), bool); done[2, 1] = True
  scores = -np.arange(12, dtype=np.float32).reshape(3, 4)
  att = np.random.RandomState(0).rand(3, 4, 2).astype(np.float32)
  eos_scores = np.full((3, 4), -9.0, np.float32)
  eos_att = np.random.RandomState(1).rand(3, 4, 2).astype(np.float32)
  out = hyps_lib.HypsFromBeamSearchOuts(hyps, prev, done, scores, att, eos_scores, eos_att,
                                        eos_id=2, num_hyps_per_beam=2)
  assert out.shape == (3, 4) and out[0, 0] == b''
  got = hyps_lib.Hypothesis.FromString(out[2, 1])
  # step 1 read slot 1 (token 22, parent 0) → step 0 read slot 0 (token 11)
  assert got.ids == [11, 22, 2] and got.beam_id == 1
  np.testing.assert_allclose(got.scores, [scores[0, 0], scores[1, 1], -9.0])
  np.testing.assert_allclose(got.atten_vecs[0], att[0, 0]); np.testing.assert_allclose(got.atten_vecs[2], eos_att[2, 1])
  ids, lens, sc = hyps_lib.UnpackHyp([out[2, 1], h.SerializeToString(), b''], max_seq_length=4)
  np.testing.assert_array_equal(ids, [[11, 22, 2, 0], [5, 9, 2, 0], [0, 0, 0, 0]])
  np.testing.assert_array_equal(lens, [3, 3, 0]); assert sc[1] == -0.75


def test_generate_proto_def(tmp_path):
  from lingvo_b200.tools import generate_proto_def
  paths = generate_proto_def.Generate(str(tmp_path))
  assert any(p.endswith('hyps.proto') for p in paths)
  text = open([p for p in paths if p.endswith('hyps.proto')][0]).read()
  assert 'repeated int32 ids = 2 [packed = true];' in text


def test_record_pipeline_is_tsan_clean():
  """Builds the yielders with -fsanitize=thread and runs the concurrency stress (≈10 s)."""
  import importlib.util
  import shutil
  if shutil.which('g++') is None:
    pytest.skip('no g++')
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'tsan_host.py')
  spec = importlib.util.spec_from_file_location('tsan_host', path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  ok, out = mod.Run('thread')
  assert ok, out[-3000:]
