"""Image summaries / sequence + attention plots (ref lingvo/core/summary_utils_test.py)."""
import numpy as np
import torch

from lingvo_b200.core import cluster_factory
from lingvo_b200.core import plot
from lingvo_b200.core import summary_utils
from lingvo_b200.utils import tfevents


def test_png_roundtrip_gray_rgb_rgba():
  rng = np.random.RandomState(0)
  for shape in [(5, 7), (4, 6, 3), (3, 3, 4)]:
    a = rng.randint(0, 256, shape).astype(np.uint8)
    png = tfevents.EncodePng(a)
    assert png[:8] == b'\x89PNG\r\n\x1a\n'
    np.testing.assert_array_equal(tfevents.DecodePng(png), a)
  f = tfevents.DecodePng(tfevents.EncodePng(np.array([[0.0, 0.5, 1.0, 2.0]])))
  assert f.tolist() == [[0, 128, 255, 255]]


def test_sequence_length_and_prepare():
  pad = torch.tensor([[0., 0, 0, 1], [0, 1, 1, 1]])
  assert summary_utils.SequenceLength(pad).tolist() == [3, 1]
  x = torch.randn(2, 4, 3, 2)
  t, lens = summary_utils.PrepareSequenceForPlot(x, pad, 'feat')
  assert t.shape == (2, 6, 4) and lens.tolist() == [3, 1]
  assert summary_utils.GetTensorName(x, 'n', 2) == 'n_2'


def test_attention_summary_writes_trimmed_images(tmp_path):
  b, tl, sl = 2, 5, 6
  probs = torch.softmax(torch.randn(b, tl, sl), -1)
  src_pad = torch.zeros(b, sl); src_pad[1, 4:] = 1
  tgt_pad = torch.zeros(b, tl); tgt_pad[1, 3:] = 1
  with cluster_factory.SetEval() if hasattr(cluster_factory, 'SetEval') else _Null():
    with summary_utils.SummaryCollector() as c:
      if not summary_utils._ShouldAddSummary():
        import pytest
        pytest.skip('summaries are off in this cluster mode')
      summary_utils.AddAttentionSummaryBatchMajor('dec', [probs, probs], [src_pad], tgt_pad)
      # time-major entry point produces the same thing
      summary_utils.AddAttentionSummary('tm', [probs.transpose(0, 1)], src_pad.t(), tgt_pad.t())
      feats, lens = summary_utils.PrepareSequenceForPlot(torch.rand(b, tl, 4), tgt_pad, 'f')
      summary_utils.PlotSequenceFeatures([(feats, lens)], 'features')
      counter = summary_utils.StatsCounter('examples_seen')
      assert int(counter.IncBy(5)) == 5 and int(counter.IncBy(2)) == 7
      assert int(counter.Value()) == 7
  assert 'dec/Attention/average_normalized_entropy/1' in c.scalars
  assert float(c.scalars['examples_seen']) == 5.0
  assert len(c.images['dec/Attention']) == 2 and len(c.images['tm/Attention']) == 2
  w = tfevents.EventFileWriter(str(tmp_path))
  c.WriteTo(w, 3)
  w.close()
  imgs = {tag: img for _, tag, img in tfevents.ReadImages(w.path)}
  assert 'dec/Attention/image/0' in imgs and 'features/image/1' in imgs
  if not plot._HAS_MPL:
    # heat-map fallback: 4 px per cell; example 1 is trimmed to 3 × 4, two matrices stacked
    assert imgs['dec/Attention/image/0'].shape == ((tl * 2 + 1) * 4, sl * 4)
    assert imgs['dec/Attention/image/1'].shape == ((3 * 2 + 1) * 4, 4 * 4)
    got = 1.0 - imgs['tm/Attention/image/1'][::4, ::4] / 255.0
    np.testing.assert_allclose(got, probs[1, :3, :4].numpy(), atol=1 / 255.0 + 1e-6)
    assert imgs['features/image/1'].shape == (4 * 4, 3 * 4)


class _Null:
  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False


def test_multi_curve_data_and_figure_summary():
  t1 = np.arange(8, dtype=np.float64).reshape(2, 4)
  pad = np.array([[0, 0, 0, 1], [0, 0, 1, 1]], np.float64)
  data, lens, labels = plot.MultiCurveData([t1, None, t1 * 2], pad, ['a', 'skip', 'b'])
  assert data.shape == (2, 2, 4) and lens.tolist() == [3, 2] and labels == ['a', 'b']
  assert data[0, 0].tolist() == [0, 1, 2, 0] and data[1, 1].tolist() == [8, 10, 0, 0]
  fig = plot.MatplotlibFigureSummary('curves')
  plot.AddMultiCurveSubplot(fig, [t1, t1 * 2], [pad, pad], ['a', 'b'], xlabels=['x0', 'x1'],
                            title='t')
  pngs = fig.Finalize()
  assert (pngs is None) == (not plot._HAS_MPL)
  val = tfevents.ImageValue('m/image', np.zeros((2, 3, 3), np.uint8))
  assert b'm/image' in val


def test_image_summary_of_tensor_batches(tmp_path):
  with summary_utils.SummaryCollector() as c:
    if summary_utils._ShouldAddSummary():
      summary_utils.image('img', torch.rand(5, 4, 6, 3))
      summary_utils.image_v2('one', torch.rand(4, 6))
  if not c.images:
    return
  w = tfevents.EventFileWriter(str(tmp_path))
  c.WriteTo(w, 1)
  w.close()
  tags = sorted(tag for _, tag, _ in tfevents.ReadImages(w.path))
  assert tags == ['img/image/0', 'img/image/1', 'img/image/2', 'one/image']
