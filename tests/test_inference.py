"""Inference export + Predictor round trip (CPU)."""

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import inference_graph_exporter
from lingvo_b200.core import predictor
from lingvo_b200.core import tokenizers
from lingvo_b200.models.lm import input_generator as lm_inp
from lingvo_b200.models.lm import layers as lm_layers
from lingvo_b200.models.lm import model as lm_model


def _Cfg(tmp_path):
  text = tmp_path / 't.txt'
  text.write_text('the cat sat\n' * 50)
  inp = lm_inp.LmInput.Params().Set(name='inp', file_pattern='text:' + str(text),
                                    bucket_upper_bound=[40], bucket_batch_limit=[4],
                                    file_buffer_size=8, file_parallelism=1)
  inp.tokenizer = tokenizers.AsciiTokenizer.Params()
  task = lm_model.LanguageModel.Params().Set(name='lm', input=inp)
  task.lm = lm_layers.RnnLm.CommonParams(vocab_size=76, emb_dim=8, num_layers=1, rnn_dims=8)
  cfg = base_model.SingleTaskModel.Params(task)
  return cfg


def test_export_and_predict(tmp_path):
  cfg = _Cfg(tmp_path)
  model = cfg.Instantiate()
  out_dir = str(tmp_path / 'export')
  graph = inference_graph_exporter.InferenceGraphExporter.Export(
      cfg, export_path=out_dir, model=model)
  assert 'default' in graph.subgraphs
  assert graph.subgraphs['default']['feeds'] == ['ids', 'paddings']
  loaded = inference_graph_exporter.LoadInferenceGraph(out_dir)
  assert loaded.subgraphs == graph.subgraphs
  pred = predictor.Predictor(out_dir, device='cpu', model_cfg=cfg.Copy())
  ids = torch.randint(3, 30, (2, 7))
  pads = torch.zeros(2, 7)
  got = pred.Run(['log_pplx_per_token'], ids=ids, paddings=pads)[0]
  want = model.tasks[0]._InferenceDefault(ids, pads).log_pplx_per_token
  torch.testing.assert_close(got, want.detach(), atol=1e-5, rtol=1e-5)
  # signature introspection (reference predictor.py:157-222)
  assert pred.feed_keys == ['ids', 'paddings'] == pred.subgraph_feed_keys('default')
  assert 'log_pplx_per_token' in pred.fetch_keys
  assert pred.fetch_shapes.log_pplx_per_token in (None, [2, 7], list(got.shape))
  assert set(pred.subgraph_feed_shapes('default').keys()) == {'ids', 'paddings'}
  import pytest
  with pytest.raises(KeyError):
    pred.subgraph_fetch_keys('no_such_subgraph')


def test_inference_graph_proto_round_trip(tmp_path):
  """`inference_graph.pb` is a well-formed `tensorflow.lingvo.InferenceGraph` message:
  subgraph feeds / fetches (+ dtype / shape metadata), hyper-parameters, asset list."""
  import os
  from lingvo_b200.utils import protowire as pw
  cfg = _Cfg(tmp_path)
  out_dir = str(tmp_path / 'export')
  inference_graph_exporter.InferenceGraphExporter.Export(cfg, export_path=out_dir,
                                                         model=cfg.Instantiate())
  pb = os.path.join(out_dir, 'inference_graph.pb')
  assert os.path.exists(pb)
  buf = open(pb, 'rb').read()
  fields = pw.parse_dict(buf)
  assert set(fields) >= {1, 5, 7, 10}                              # graph_def, subgraphs, …
  graph, hyper, assets = inference_graph_exporter.InferenceGraph.FromProto(buf)
  assert graph.subgraphs['default']['feeds'] == ['ids', 'paddings']
  assert hyper == cfg.ToText() and 'weights.pt' in assets
  via_pb = inference_graph_exporter.LoadInferenceGraph(pb)
  assert via_pb.subgraphs['default']['feeds'] == ['ids', 'paddings']
  # metadata round trip
  g = inference_graph_exporter.InferenceGraph(subgraphs={
      'enc': {'feeds': ['x'], 'fetches': ['y', 'z'],
              'feeds_meta': {'x': {'dtype': 'bfloat16', 'shape': [None, 128], 'layout': 'BD'}},
              'fetches_meta': {'y': {'dtype': 'int64', 'shape': [4]}}}})
  g2, _, _ = inference_graph_exporter.InferenceGraph.FromProto(g.ToProto())
  assert g2.subgraphs['enc']['fetches'] == ['y', 'z']
  assert g2.subgraphs['enc']['feeds_meta']['x'] == {'dtype': 'bfloat16',
                                                   'shape': [None, 128], 'layout': 'BD'}
  assert g2.subgraphs['enc']['fetches_meta']['y'] == {'dtype': 'int64', 'shape': [4]}
