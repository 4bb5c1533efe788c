"""Point-set recipes (GIN, set abstraction, PointConv) and the pillars backbone
(ref tasks/car/builder_lib_test.py, pillars_test.py)."""
import pytest
import torch

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import builder_lib
from lingvo_b200.models.car import pillars


def _Points(b=2, p=7, f=4, seed=0):
  g = torch.Generator().manual_seed(seed)
  pad = torch.zeros(b, p)
  pad[1, 5:] = 1
  return NestedMap(points=torch.randn(b, p, 3, generator=g),
                   features=torch.randn(b, p, f, generator=g), padding=pad)


def _B():
  return builder_lib.ModelBuilderBase.Params().Instantiate() if hasattr(
      builder_lib.ModelBuilderBase, 'Params') and hasattr(
          builder_lib.ModelBuilderBase.Params(), 'Instantiate') else builder_lib.ModelBuilderBase()


def _Make(layer_p):
  return layer_p.Set(name=layer_p.name or 'l').Instantiate()


def test_padded_pooling_ignores_padded_points():
  b = _B()
  inp = _Points()
  mean = _Make(b._PaddedMean('mean')).FPropDefaultTheta(inp)
  total = _Make(b._PaddedSum('sum')).FPropDefaultTheta(inp)
  torch.testing.assert_close(mean[0], inp.features[0].mean(0))
  torch.testing.assert_close(mean[1], inp.features[1, :5].mean(0))
  torch.testing.assert_close(total[1], inp.features[1, :5].sum(0))
  allpad = inp.copy(); allpad.padding = torch.ones(2, 7)
  assert _Make(b._PaddedMean('mean')).FPropDefaultTheta(allpad).abs().sum() == 0


def test_features_mlp_keeps_the_points_tensor_shape():
  b = _B()
  layer = _Make(b._FeaturesMLP('mlp', [4, 8, 6], use_bn=False))
  out = layer.FPropDefaultTheta(_Points())
  assert sorted(out.keys()) == ['features', 'padding', 'points'] and out.features.shape == (2, 7, 6)
  cat = _Make(b._ConcatPointsToFeatures('cat')).FPropDefaultTheta(_Points())
  assert cat.features.shape == (2, 7, 7)
  torch.testing.assert_close(cat.features[..., :3], cat.points)


@pytest.mark.parametrize('combine', ['add', 'concat', 'cond_fc'])
def test_gin_concatenates_readouts_of_every_depth(combine):
  b = _B()
  dims = [[4, 4], [8 if combine == 'concat' else 4, 6]]
  if combine == 'concat':
    dims = [[8, 4], [8, 6]]
    inp = _Points(f=4)
    # with concat the first MLP sees [f ‖ aggregate]: 2·4
  else:
    inp = _Points(f=4)
  gin = _Make(b._GIN('gin', dims, b._PaddedMean('agg'), b._PaddedMax('readout'),
                     combine_method=combine, use_bn=False))
  out = gin.FPropDefaultTheta(inp)
  assert out.shape == (2, 4 + 4 + 6)                            # readout(f0) ‖ f1 ‖ f2
  torch.testing.assert_close(out[0, :4], inp.features[0].max(0).values)
  torch.testing.assert_close(out[1, :4], inp.features[1, :5].max(0).values)
  # padded points never influence the result
  inp2 = inp.DeepCopy()
  inp2.features[1, 5:] = 100.0
  torch.testing.assert_close(gin.FPropDefaultTheta(inp2), out)
  with pytest.raises(ValueError):
    b._GIN('bad', [[4, 4], [5, 6]], b._PaddedMean('a'), b._PaddedMax('r'))
  with pytest.raises(ValueError):
    b._GIN('bad', dims, b._PaddedMean('a'), b._PaddedMax('r'), combine_method='mul')


def test_cond_fc_applies_a_per_example_matrix():
  b = _B()
  layer = _Make(b._CondFC('cfc', idims=4, adims=3, odims=5, use_bn=False, activation_fn='NONE'))
  feats, agg = torch.randn(2, 7, 4), torch.randn(2, 1, 3)
  out = layer.FPropDefaultTheta(feats, agg)
  assert out.shape == (2, 7, 5)
  # linear in the features for a fixed aggregate
  out2 = layer.FPropDefaultTheta(feats * 2, agg)
  bias = layer.FPropDefaultTheta(torch.zeros_like(feats), agg)
  torch.testing.assert_close(out2 - bias, 2 * (out - bias), atol=1e-5, rtol=1e-5)
  # and different aggregates give different maps
  assert (layer.FPropDefaultTheta(feats, agg + 1.0) - out).abs().max() > 1e-4


def test_set_abstraction_and_pointconv():
  b = _B()
  inp = _Points(b=2, p=12, f=4, seed=3)
  inp.padding = torch.zeros(2, 12)
  extract = b._Seq('fx', b._FeaturesMLP('mlp', [7, 8], use_bn=False), b._PaddedMax('max'))
  sa = _Make(b._SetAbstraction('sa', extract, num_samples=5, group_size=4, ball_radius=2.0))
  out = sa.FPropDefaultTheta(inp)
  assert out.points.shape == (2, 5, 3) and out.features.shape == (2, 5, 8)
  assert out.padding.shape == (2, 5)
  pc = _Make(b._PointConvParametricConv('pc', [3, 8, 2], num_in_channels=4, num_out_channels=6))
  y = pc.FPropDefaultTheta(_Points(f=4))
  assert y.shape == (2, 6)
  with pytest.raises(ValueError):
    b._PointConvParametricConv('pc', [4, 8], 4, 6)


def test_pillars_backbone_runs_each_block_once_and_sparse_to_dense():
  pb = pillars.Builder.Params().Instantiate() if hasattr(
      pillars.Builder.Params(), 'Instantiate') else pillars.Builder()
  bb = _Make(pb.Backbone(idims=8, dims=(64, 128, 256), repeats=(2, 1, 1), up_dims=16))
  x = torch.randn(2, 16, 16, 8)
  y = bb.FPropDefaultTheta(x)
  assert y.shape == (2, 8, 8, 48)                        # 1/2 resolution, 3 · up_dims
  names = [v.var_name for v in bb.vars.Flatten()]
  assert sum('topdown/b0/c3x3' in n and n.endswith('/w/var') for n in names) == 1
  contract = _Make(pb.Contract(idims=8, repeats=(1, 1, 1)))
  outs = contract.FPropDefaultTheta(x)
  assert [tuple(o.shape) for o in outs] == [(2, 2, 2, 256), (2, 4, 4, 128), (2, 8, 8, 64)]
  # SparseToDense: features land in their cells, duplicates add up
  loc = torch.tensor([[[0, 0, 0], [1, 2, 0], [1, 2, 0]]])
  feats = torch.tensor([[[1.0, 2.0], [3.0, 4.0], [10.0, 20.0]]])
  grid = pillars.SparseToDense((2, 3, 1), loc, feats)
  assert grid.shape == (1, 2, 3, 2)
  assert grid[0, 0, 0].tolist() == [1, 2] and grid[0, 1, 2].tolist() == [13, 24]
  assert float(grid.sum()) == 40.0
  mlp = _Make(pb.ScalePillarsFeaturizer('f', 4, 5))
  assert mlp.FPropDefaultTheta(_Points()).shape == (2, 7, 5)
