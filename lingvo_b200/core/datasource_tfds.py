"""TFDS-style named datasets (ref `lingvo/core/datasource_tfds.py`).

TensorFlow Datasets is not available offline; `NamedDataSource` resolves a dataset name
to local files under `data_dir` (`<data_dir>/<dataset>/<split>*`) and reads them through
the native record yielders, which is the role `TFDSInput` plays for the reference."""
import os

from lingvo_b200.core import datasource


class NamedDataSource(datasource.SimpleDataSource):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dataset', None, 'Dataset name (directory under data_dir).')
    p.Define('split', 'train', 'Split prefix.')
    p.Define('data_dir', os.environ.get('LINGVO_B200_DATA', '/tmp/lingvo_b200_data'), 'Root.')
    p.Define('load_fn', '', 'Kept for parity.')
    p.Define('shuffle_buffer_size', 10000, 'Kept for parity.')
    return p

  def _Patterns(self):
    p = self.params
    ftype = p.file_type or 'tfrecord'
    return '%s:%s' % (ftype, os.path.join(p.data_dir, p.dataset, p.split + '*')), None


TFDSInput = NamedDataSource
