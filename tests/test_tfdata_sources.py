"""The dataset-backed DataSource family (ref `lingvo/core/datasource.py:351-900`,
`datasource_test.py`)."""

import numpy as np
import pytest

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import datasource as ds
from lingvo_b200.core.nested_map import NestedMap


class _Gen(base_input_generator.BaseInputGenerator):
  """Input generator providing the hooks the sources call by name."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('repeat_steps', None, '')
    p.Define('repeat_with_sentinel', None, '')
    p.Define('n', 10, '')
    return p

  def LoadDataset(self, start=0):
    n = self.params.n
    return ds.Dataset.FromElements(
        [NestedMap(x=np.arange(i % 5 + 1, dtype=np.float32) + 10 * i, id=np.int32(i))
         for i in range(start, start + n)])

  def GetSequenceLength(self, example):
    return len(example.x)

  def _InputShape(self, key):
    return (None,) if key == 'x' else ()

  def _InputPaddingValue(self, key, spec):
    return -1 if key == 'x' else 0

  def Double(self, dataset, by=2):
    return dataset.map(lambda e: NestedMap(x=e.x * by, id=e.id))


def _Wire(source_params, **gen_kw):
  gen = _Gen.Params().Set(name='gen', batch_size=1, **gen_kw).Instantiate()
  src = source_params.Instantiate()
  src.SetInputGenerator(gen)
  return src


def _Drain(src, limit=1000):
  out = []
  try:
    for _ in range(limit):
      out.append(src.GetNext())
  except StopIteration:
    pass
  return out


def test_dataset_algebra():
  d = ds.Dataset.FromElements(range(10))
  assert list(d.map(lambda x: x * 2).filter(lambda x: x % 3 == 0)) == [0, 6, 12, 18]
  assert list(d.take(3).repeat(2)) == [0, 1, 2, 0, 1, 2]
  assert list(d.shard(3, 1)) == [1, 4, 7]
  assert list(d.take(2).concatenate(d.take(1))) == [0, 1, 0]
  sh = d.shuffle(4, seed=1)
  a, b = list(sh), list(sh)
  assert sorted(a) == list(range(10)) and a != list(range(10)) and a != b   # reshuffled
  assert list(d.prefetch(2)) == list(range(10))
  mix = ds.Dataset.SampleFrom([ds.Dataset.FromElements([0] * 50),
                               ds.Dataset.FromElements([1] * 50)], [0.9, 0.1], seed=0)
  got = list(mix)
  assert len(got) == 100 and sum(got[:30]) < 10                           # weights respected
  boom = ds.Dataset.FromGenerator(lambda: (1 / (2 - i) for i in range(5))).prefetch(1)
  with pytest.raises(ZeroDivisionError):
    list(boom)


def test_fn_input_shuffles_repeats_and_respects_eval():
  with pytest.raises(ValueError, match='shuffle_buffer_size'):
    ds.TFDatasetFnInput.Params().Set(name='s').Instantiate()
  src = _Wire(ds.TFDatasetFnInput.Params().Set(name='s', shuffle_buffer_size=4, random_seed=3))
  ids = [int(src.GetNext().id) for _ in range(25)]                        # repeats in training
  assert sorted(ids[:10]) == list(range(10)) and ids[:10] != list(range(10))
  assert sorted(ids[10:20]) == list(range(10)) and ids[10:20] != ids[:10]
  with cluster_factory.ForTestingWorker(do_eval=True):
    with cluster_factory.Cluster(cluster_factory.Current().params.Copy().Set(
        require_sequential_input_order=True, do_eval=True)):
      ev = _Wire(ds.TFDatasetFnInput.Params().Set(name='e', kwargs=dict(start=100)))
      out = _Drain(ev)
      assert [int(e.id) for e in out] == list(range(100, 110))            # one ordered epoch
      ev.Reset()
      assert int(ev.GetNext().id) == 100


def test_custom_transform_and_adaptor_over_plain_sources():
  base = ds.TFDatasetFnInput.Params().Set(name='s', shuffle_buffer_size=1)
  src = _Wire(ds.CustomTFDatasetTransform.Params().Set(name='t', sub=base, fn='Double',
                                                       kwargs=dict(by=3)))
  e = src.GetNext()
  assert np.allclose(e.x, 3 * (np.arange(int(e.id) % 5 + 1) + 10 * int(e.id)))
  # a non-dataset DataSource is adapted automatically
  it = ds.IteratorDataSource.Params().Set(
      name='it', iter_fn=lambda: iter([NestedMap(x=np.ones(2, np.float32), id=np.int32(7))]),
      repeat=False)
  ad = _Wire(ds.CustomTFDatasetTransform.Params().Set(name='t2', sub=it, fn='Double'))
  assert isinstance(ad.sub, ds.TFDatasetAdaptor)
  assert np.allclose(ad.GetNext().x, 2.0)
  with pytest.raises(StopIteration):
    ad.GetNext()


def test_batch_by_sequence_length_pads_to_bucket_boundaries():
  with cluster_factory.Cluster(cluster_factory.Current().params.Copy().Set(
      require_sequential_input_order=True, do_eval=True)):
    src = _Wire(ds.TFDatasetBatchBySequenceLength.Params().Set(
        name='b', sub=ds.TFDatasetFnInput.Params().Set(name='s'),
        bucket_upper_bound=[2, 4], bucket_batch_limit=[3, 2]), n=10)
    batches = _Drain(src)
  # lengths cycle 1..5: length-5 examples exceed the last bucket and are dropped
  seen = sorted(int(i) for b in batches for i in b.id)
  assert seen == [0, 1, 2, 3, 5, 6, 7, 8]
  for b in batches:
    assert b.x.shape[1] in (2, 4) and b.x.shape[0] == b.id.shape[0] == b.bucket_keys.shape[0]
    assert b.x.shape[0] <= (3 if b.x.shape[1] == 2 else 2)
    for row, n in zip(b.x, b.bucket_keys):
      assert (row[n:] == -1).all() and (row[:n] >= 0).all()              # custom padding value
  assert sum(b.x.shape[0] for b in batches) == 8                          # remainder flushed


def test_repeatable_transform_steps_and_sentinel():
  base = ds.TFDatasetBatchBySequenceLength.Params().Set(
      name='b', sub=ds.TFDatasetFnInput.Params().Set(name='s', shuffle_buffer_size=1),
      bucket_upper_bound=[5], bucket_batch_limit=[2])
  with cluster_factory.Cluster(cluster_factory.Current().params.Copy().Set(do_eval=True)):
    rep = _Wire(ds.RepeatableTFDatasetTransform.Params().Set(name='r', sub=base),
                repeat_steps=2)
    ids = [tuple(int(i) for i in rep.GetNext().id) for _ in range(6)]
    assert ids[0] != ids[1] and ids[:2] == ids[2:4] == ids[4:6]          # first 2 batches forever
    sen = _Wire(ds.RepeatableTFDatasetTransform.Params().Set(name='r2', sub=base),
                repeat_with_sentinel=True)
    epoch = []
    with pytest.raises(ds.RepeatSentinelError, match='REPEAT_SENTINEL_'):
      for _ in range(100):
        epoch.append(sen.GetNext())
    assert len(epoch) == 5                                                # 10 examples / 2
    again = [sen.GetNext() for _ in range(5)]                             # pipeline continues
    assert [tuple(b.id) for b in again] == [tuple(b.id) for b in epoch]


def test_mixer_tags_sources_and_broadcasts_structures():
  a = ds.IteratorDataSource.Params().Set(
      name='a', iter_fn=lambda: (NestedMap(x=np.float32([i])) for i in range(100)), repeat=False)
  b = ds.IteratorDataSource.Params().Set(
      name='b', iter_fn=lambda: (NestedMap(x=np.float32([i]), extra=np.zeros((2, 3), np.int64) + i)
                                 for i in range(100)), repeat=False)
  mix = _Wire(ds.TFDatasetMixer.Params().Set(name='m', sub=[a, b], weights=[0.5, 0.5],
                                             broadcast_dataset_structures=True, random_seed=5))
  out = _Drain(mix)
  assert len(out) == 200
  ids = [int(e.source_id) for e in out]
  assert 60 < sum(ids[:150]) < 90
  for e in out:
    assert e.extra.dtype == np.int64
    if int(e.source_id) == 0:
      assert e.extra.shape == (1, 1) and not e.extra.any()               # broadcast placeholder
    else:
      assert e.extra.shape == (2, 3)
  solo = _Wire(ds.TFDatasetMixer.Params().Set(name='m1', sub=[a]))
  assert int(solo.GetNext().source_id) == 0


def test_prefetch_and_data_service_cover_each_element_once():
  base = ds.IteratorDataSource.Params().Set(
      name='a', iter_fn=lambda: (NestedMap(id=np.int32(i)) for i in range(57)), repeat=False)
  pre = _Wire(ds.TFDatasetPrefetch.Params().Set(name='p', sub=base, buffer_size=3))
  assert [int(e.id) for e in _Drain(pre)] == list(range(57))
  svc_p = ds.TFDataServiceSource.Params().Set(
      name='svc', num_workers=4, buffer_size=5,
      sub=ds.TFDatasetFnInput.Params().Set(name='s', shuffle_buffer_size=1))
  with cluster_factory.Cluster(cluster_factory.Current().params.Copy().Set(do_eval=True)):
    svc = _Wire(svc_p, n=41)
    got = sorted(int(e.id) for e in _Drain(svc))
  assert got == list(range(41))                                           # exactly once each


def _my_dataset(begin=0, end=10, scale=1.0, batch=1):
  return ds.Dataset.FromElements({'value': np.float32(i * scale), 'b': np.int32(batch)}
                                 for i in range(begin, end))


MyInput = base_input_generator.DefineTFDataInput('MyInput', _my_dataset,
                                                 map_args={'batch': 'batch_size'})


def test_define_tfdata_input_generates_params_from_the_signature():
  p = MyInput.Params()
  assert p.args.begin == 0 and p.args.end == 10 and p.args.scale == 1.0
  assert 'batch' not in p.args                       # mapped from p.batch_size instead
  assert MyInput.__module__ == __name__
  p.name = 'my'
  p.batch_size = 7
  p.args.begin, p.args.end, p.args.scale = 2, 5, 0.5
  with cluster_factory.Cluster(cluster_factory.Current().params.Copy().Set(
      do_eval=True, require_sequential_input_order=True)):
    ig = p.Instantiate()
    assert isinstance(ig, MyInput)
    vals = []
    with pytest.raises(StopIteration):
      while True:
        b = ig.GetPreprocessedInputBatch()
        assert isinstance(b, NestedMap) and int(b.b) == 7
        vals.append(float(b.value))
  assert vals == [1.0, 1.5, 2.0]
  # text round trip keeps the pipeline arguments
  from lingvo_b200.core import hyperparams
  q = MyInput.Params()
  q.FromText(p.ToText())
  assert q.args.end == 5 and q.args.scale == 0.5
  del hyperparams
