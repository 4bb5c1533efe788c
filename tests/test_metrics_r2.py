"""Metrics added in round 2, checked against sklearn / scipy (ref lingvo/core/metrics_test.py)."""
import numpy as np
import pytest
import torch

from lingvo_b200.core import metrics


def _Data(n=200, seed=0):
  rng = np.random.RandomState(seed)
  y = (rng.rand(n) > 0.6).astype(int)
  s = np.clip(0.35 * y + rng.rand(n) * 0.7, 0, 1).round(2)      # ties on purpose
  w = rng.rand(n) + 0.1
  return y.tolist(), s.tolist(), w.tolist()


def test_precision_recall_operating_points_match_sklearn():
  sk = pytest.importorskip('sklearn.metrics')
  y, s, w = _Data()
  p, r, t = sk.precision_recall_curve(y, s, sample_weight=w)

  def SkPatR(rec):
    last = 0.0
    for pp, rr, _ in zip(p, r, t):
      if rr >= rec:
        last = pp
    return last

  def SkRatP(prec):
    for pp, rr, _ in zip(p, r, t):
      if pp >= prec:
        return rr
    return 0.0

  for th in (0.2, 0.5, 0.8, 0.95):
    m = metrics.PrecisionAtRecall(th)
    m.Update(y, s, w)
    assert m.value == pytest.approx(SkPatR(th), abs=1e-9)
    m2 = metrics.RecallAtPrecision(th)
    m2.Update(y, s, w)
    assert m2.value == pytest.approx(SkRatP(th), abs=1e-9)
  auc = metrics.AUCMetric('pr')
  auc.Update(y, s, w)
  assert auc.value == pytest.approx(sk.average_precision_score(y, s, sample_weight=w), abs=1e-9)
  empty = metrics.PrecisionAtRecall(0.5)
  empty.Update([1, 1], [0.2, 0.3])
  assert empty.value == 0.0


def test_correlation_modes_match_scipy():
  st = pytest.importorskip('scipy.stats')
  rng = np.random.RandomState(1)
  t = rng.randint(0, 6, 40).astype(float)
  p = t + rng.randn(40)
  p[::5] = p[1::5][:8]                               # some ties in pred too
  for mode, fn in (('pearson', st.pearsonr), ('spearman', st.spearmanr),
                   ('kendalltau', st.kendalltau)):
    m = metrics.CorrelationMetric(mode)
    m.Update(t.tolist(), p.tolist())
    assert m.value == pytest.approx(fn(t, p)[0], abs=1e-9), mode
  keyed = metrics.AverageKeyedCorrelationMetric('spearman')
  keyed.Update('a', t[:20].tolist(), p[:20].tolist())
  keyed.Update('b', t[20:].tolist(), p[20:].tolist())
  keyed.Update('single', [1.0], [2.0])               # undefined → skipped
  keyed.Update('const', [1.0, 1.0], [2.0, 3.0])      # undefined → skipped
  want = np.mean([st.spearmanr(t[:20], p[:20])[0], st.spearmanr(t[20:], p[20:])[0]])
  assert keyed.value == pytest.approx(want, abs=1e-9)
  assert metrics.AverageKeyedCorrelationMetric().value == 0.0


def test_group_pair_auc_streams_contiguous_groups():
  sk = pytest.importorskip('sklearn.metrics')
  gids = [0, 0, 0, 1, 1, 1, 1, 0, 0]
  tgt = [1.0, 0.0, 2.0, 0.5, 0.5, 1.0, 0.0, 3.0, 1.0]
  lg = [0.2, 0.1, 0.9, -1.0, 0.3, 0.2, 0.4, 0.0, 1.0]
  m = metrics.GroupPairAUCMetric()
  m.UpdateRaw(gids[:5], tgt[:5], lg[:5])
  m.UpdateRaw(gids[5:], tgt[5:], lg[5:])
  labels, probs = [], []
  for s, e in [(0, 3), (3, 5), (5, 7), (7, 9)]:      # the second update starts a new chunk
    for i in range(s, e):
      for j in range(i + 1, e):
        if tgt[i] != tgt[j]:
          labels.append(int(tgt[i] > tgt[j]))
          probs.append(1 / (1 + np.exp(-(lg[i] - lg[j]))))
  assert m.value == pytest.approx(sk.roc_auc_score(labels, probs), abs=1e-9)
  m2 = metrics.GroupPairAUCMetric()
  m2.UpdateRaw(gids, tgt, lg, weight=[0.3] * 9, ignore_ids=[0, 0, 1, 0, 0, 0, 0, 0, 0])
  assert all(abs(w - 0.6) < 1e-9 for w in m2._weight) and len(m2._label) < len(labels) + 3


def test_sampling_metric_params_and_summary_reset():
  class Texts(metrics.SamplingMetric):
    def _CreateSummary(self, name):
      return (name, sorted(self.samples))

  m = Texts(Texts.Params().Set(num_samples=3))
  for i in range(10):
    m.Update(i)
  assert len(m.samples) == 3
  name, kept = m.Summary('x')
  assert name == 'x' and len(kept) == 3
  assert m.Summary('x')[1] == kept and len(m.samples) == 0      # cached; sampler restarted
  m.Update(42)
  assert m.Summary('x')[1] == [42]
  legacy = metrics.SamplingMetric(4)
  legacy.Update('a', 1, k=2)
  assert legacy.samples == [(('a', 1), {'k': 2})]
  assert isinstance(metrics.ConfigurableMetric.Params().Instantiate(), metrics.ConfigurableMetric)


def test_device_variable_metrics_fixed_slots():
  m = metrics.TpuVariableMetrics(max_metrics=4)
  assert len(m.variables) == 8
  m.AccumulateStepMetrics({'loss': (torch.tensor(2.0), torch.tensor(3.0)), 'acc': (0.5, 1.0)})
  m.AccumulateStepMetrics({'loss': (torch.tensor(4.0), torch.tensor(1.0)), 'acc': (1.0, 1.0)})
  out = m.FinalizeMetricsWithStructure({'loss': None, 'acc': None})
  assert float(out['loss'][0]) == pytest.approx(2.5) and float(out['loss'][1]) == 4.0
  assert float(out['acc'][0]) == pytest.approx(0.75)
  m.ResetState()
  assert float(sum(m.variables)) == 0.0
  with pytest.raises(AssertionError):
    m.AccumulateStepMetrics({str(i): (1.0, 1.0) for i in range(5)})


def test_tpu_eval_metrics_loop_carried_form_and_average_metric_setters():
  import torch
  from lingvo_b200.core import metrics
  m = metrics.TpuEvalMetrics(max_metrics=4)
  assert len(m.initial_values) == 8
  carried = m.initial_values
  for step in range(3):
    step_out = m.PackStepMetricsForAccumulation(
        {'loss': (torch.tensor(float(step)), torch.tensor(2.0)),
         'acc': (torch.tensor(0.5), torch.tensor(1.0))}, carried)
    assert len(step_out) == 8
    carried = [a + b for a, b in zip(carried, step_out[:4])] + list(carried[4:])
  final = m.FinalizeMetrics(carried)
  assert len(final) == 4
  m.PackMetricsValues(final)
  assert m.metrics['acc'] == (0.5, 3.0)
  assert m.metrics['loss'] == (1.0, 6.0)
  avg = m.ToAverageMetrics()
  assert avg['loss'].value == 1.0 and avg['loss'].total_weight == 6.0
  a = metrics.AverageMetric()
  a.SetTotalValue(6.0)
  a.SetTotalWeight(4.0)
  assert a.GetTotalValue() == 6.0 and a.GetTotalWeight() == 4.0 and a.value == 1.5
  import pytest
  with pytest.raises(AssertionError):
    metrics.TpuEvalMetrics(max_metrics=1).PackStepMetricsForAccumulation(
        {'a': (torch.tensor(1.), torch.tensor(1.)), 'b': (torch.tensor(1.), torch.tensor(1.))},
        [])


def test_auc_metric_summary_has_scalar_and_curve_image(tmp_path):
  from lingvo_b200.core import metrics
  from lingvo_b200.utils import tfevents
  for mode in ('roc', 'pr'):
    m = metrics.AUCMetric(mode=mode)
    m.Update([1, 0, 1, 0, 1], [0.9, 0.8, 0.7, 0.3, 0.2])
    xs, ys, labels = m.Curve()
    assert len(xs) == len(ys) and len(labels) == 2
    summ = m.Summary('auc_' + mode)
    w = tfevents.EventFileWriter(str(tmp_path / mode))
    w.add_summary_bytes(summ, 1)
    w.close()
    assert isinstance(summ, bytes) and b'\x89PNG' in summ
    imgs = list(tfevents.ReadImages(w.path))
    assert len(imgs) == 1 and imgs[0][1] == 'auc_' + mode and imgs[0][2].shape[0] >= 64
    assert (imgs[0][2] == 0).any()                     # the curve was drawn
