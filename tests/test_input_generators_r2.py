"""File-pattern → datasource conversion, Dataset-backed sequence inputs and tf.Example inputs
(ref lingvo/core/base_input_generator_test.py)."""
import numpy as np
import pytest
import torch

from lingvo_b200.core import base_input_generator as big
from lingvo_b200.core import datasource
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.utils import tf_example
from lingvo_b200.utils import tfrecord


class _FakeFiles(big.BaseInputGeneratorFromFiles):
  """Every 'file pattern' is a number; batches carry it so the mixing can be observed."""

  def _DataSourceFromFilePattern(self, file_pattern, input_source_weights=None, **kw):
    pats = [file_pattern] if isinstance(file_pattern, str) else list(file_pattern)
    w = np.asarray(input_source_weights or [1.0] * len(pats), np.float64)
    rng = np.random.RandomState(0)

    def Next():
      ids = rng.choice(len(pats), size=4, p=w / w.sum())
      return NestedMap(value=torch.tensor([float(pats[i].split(':')[-1]) for i in ids]),
                       source_id=torch.tensor(ids, dtype=torch.int32))
    return Next


def test_file_pattern_forms_map_to_datasources():
  p = _FakeFiles.Params()
  p.file_pattern = 'tfrecord:1'
  ds = big.FilePatternToDataSource(p)
  assert ds.cls is datasource.SimpleDataSource and ds.file_pattern == 'tfrecord:1'
  p.file_pattern = ['tfrecord:1', 'tfrecord:2']
  assert big.FilePatternToDataSource(p).file_pattern == ['tfrecord:1', 'tfrecord:2']
  p.file_pattern = [('tfrecord:1', 0.3), ('tfrecord:2', 0.7)]
  p.use_within_batch_mixing = True
  ds = big.FilePatternToDataSource(p)
  assert ds.cls is datasource.SimpleDataSource and ds.weights == [0.3, 0.7]
  p.use_within_batch_mixing = False
  p.file_pattern = [('tfrecord:1', 0.3, 'a.*'), ('tfrecord:2', 0.7)]
  ds = big.FilePatternToDataSource(p)
  assert ds.cls is datasource.CrossBatchMixingDataSource and ds.weights == [0.3, 0.7]
  assert ds.bprop_variable_filters == ['a.*', '']
  assert [s.source_id_offset for s in ds.sub] == [0, 0]          # legacy: all zero
  p.all_zero_source_id_without_within_batch_mixing = False
  assert [s.source_id_offset for s in big.FilePatternToDataSource(p).sub] == [0, 1]
  with pytest.raises(ValueError):
    big.FilePatternToDataSource(p.Copy().Set(file_pattern=['tfrecord:1', ('tfrecord:2', 1.0)]))
  with pytest.raises(ValueError):
    big.FilePatternToDataSource(p.Copy().Set(file_pattern=3))


def test_partitioned_batch_mixing():
  p = _FakeFiles.Params().Set(
      file_pattern=[('t:1', 1.0), ('t:2', 3.0), ('t:3', 2.0), ('t:4', 2.0), ('t:5', 1.0)],
      batch_mixing_partition_boundaries=[2, 3],
      all_zero_source_id_without_within_batch_mixing=False)
  subs, weights = big.PartitionFilePatternsIntoDataSources(p)
  assert [s.file_pattern for s in subs] == [['t:1', 't:2'], ['t:3'], ['t:4', 't:5']]
  assert [s.weights for s in subs] == [[1.0, 3.0], [2.0], [2.0, 1.0]]
  assert weights == [4.0, 2.0, 3.0] and [s.source_id_offset for s in subs] == [0, 2, 3]
  for bad in ([0, 2], [3, 2], [2, 5]):
    with pytest.raises(ValueError):
      big.PartitionFilePatternsIntoDataSources(p.Copy().Set(
          batch_mixing_partition_boundaries=bad))
  with pytest.raises(ValueError):
    big.PartitionFilePatternsIntoDataSources(p.Copy().Set(
        file_pattern=[('t:1', 1.0, 'f'), ('t:2', 1.0)], batch_mixing_partition_boundaries=[1]))
  gen = p.Instantiate()
  assert isinstance(gen.datasource, datasource.CrossBatchMixingDataSource)
  seen = set()
  for _ in range(60):
    b = gen.GetPreprocessedInputBatch()
    vals = set(b.value.tolist())
    part = int(b.source_selected[0])
    assert vals <= [{1.0, 2.0}, {3.0}, {4.0, 5.0}][part]
    # source ids are global pattern indices thanks to the offsets
    assert torch.equal(b.source_id.float() + 1.0, b.value)
    seen |= vals
  assert seen == {1.0, 2.0, 3.0, 4.0, 5.0}
  meta = gen.datasource.GetMeta()
  assert 'bprop_variable_filters' not in meta


class _SeqInput(big.TFDataSequenceInputGenerator):

  def LoadDataset(self, file_pattern):
    base = int(file_pattern.split(':')[-1])
    return datasource.Dataset.FromElements(
        [NestedMap(n=np.int32(base + i)) for i in range(12)])

  def ProcessDataset(self, dataset):
    def Fn(ex):
      n = int(ex.n) % 7 + 1
      return NestedMap(ids=np.full([n], int(ex.n), np.int32), ids_paddings=np.zeros([n], np.float32))
    return dataset.map(Fn)

  def GetSequenceLength(self, example):
    return len(example.ids)

  def _InputShape(self, key):
    if key in ('ids', 'ids_paddings'):
      return (None,)
    return super()._InputShape(key)


def test_tfdata_sequence_input_generator_buckets_and_pads():
  p = _SeqInput.Params().Set(file_pattern='mem:100', bucket_upper_bound=[4, 8],
                             bucket_batch_limit=[3, 2], file_buffer_size=1, tokenizer=None)
  gen = p.Instantiate()
  assert gen.params.resettable
  seen = []
  for _ in range(4):
    b = gen.GetPreprocessedInputBatch()
    t = b.ids.shape[1]
    assert t in (4, 8) and b.ids.shape[0] == (3 if t == 4 else 2)
    lens = (1 - np.asarray(b.ids_paddings)).sum(1)
    assert (lens <= t).all() and (lens > (0 if t == 4 else 4)).all()
    # padded positions: ids 0, paddings 1
    assert (np.asarray(b.ids)[np.asarray(b.ids_paddings) > 0] == 0).all()
    seen += np.asarray(b.ids)[:, 0].tolist()
  gen.Reset()
  again = np.asarray(gen.GetPreprocessedInputBatch().ids)[:, 0].tolist()
  assert again == seen[:len(again)]
  # two patterns need within-batch mixing
  with pytest.raises(ValueError):
    _SeqInput.Params().Set(file_pattern='mem:0,mem:1', tokenizer=None).Instantiate()
  mix = _SeqInput.Params().Set(
      file_pattern=[('mem:0', 1.0), ('mem:1000', 1.0)], use_within_batch_mixing=True,
      bucket_upper_bound=[8], bucket_batch_limit=[4], file_buffer_size=1,
      tokenizer=None).Instantiate()
  firsts = []
  for _ in range(4):
    firsts += np.asarray(mix.GetPreprocessedInputBatch().ids)[:, 0].tolist()
  assert any(v >= 1000 for v in firsts) and any(v < 1000 for v in firsts)


class _ExampleInput(big.BaseDataExampleInputGenerator):

  def GetFeatureSpec(self):
    return {'x': (np.float32, [2]), 'y': (np.int64, [])}

  def _PreprocessInputBatch(self, batch):
    batch.x2 = batch.x * 2
    return batch


def test_example_input_generator(tmp_path):
  n = 0
  for f in range(3):
    w = tfrecord.TFRecordWriter(str(tmp_path / ('data-%d.tfrecord' % f)))
    for _ in range(5):
      w.write(tf_example.MakeExample({'x': np.asarray([n, n + 0.5], np.float32),
                                      'y': np.asarray([n], np.int64)}))
      n += 1
    w.close()
  p = _ExampleInput.Params().Set(
      input_files=str(tmp_path / 'data-*.tfrecord'), dataset_type=tfrecord.ReadRecords,
      batch_size=4, randomize_order=False, num_epochs=1, parallel_readers=2)
  gen = p.Instantiate()
  ys = []
  for b in gen:
    assert b.x.shape == (4, 2) and b.y.shape == (4,) and b.x.dtype == torch.float32
    assert torch.equal(b.x2, b.x * 2) and torch.equal(b.x[:, 0], b.y.float())
    ys += b.y.tolist()
  # 15 records → 3 full batches; 2 readers interleave files 0 and 1 first
  assert len(ys) == 12 and ys[:4] == [0, 5, 1, 6]
  shuf = p.Copy().Set(randomize_order=True, random_seed=3, num_epochs=2).Instantiate()
  ys2 = [y for b in shuf for y in b.y.tolist()]
  assert len(ys2) == 28 and sorted(set(ys2)) == list(range(15)) and ys2[:12] != ys
  few = p.Copy().Set(num_examples=8, num_epochs=-1).Instantiate()
  got = [few.GetPreprocessedInputBatch().y.tolist() for _ in range(5)]
  assert {y for g in got for y in g} == {0, 5, 1, 6, 2, 7, 3, 8}   # take(8), repeated forever
