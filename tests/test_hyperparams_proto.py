"""`Params.ToProto / FromProto / ToProtoText` (reference hyperparams.py:529,611; wire format
of `hyperparams.proto`)."""

import collections
import enum

import numpy as np
import torch

from lingvo_b200.core import hyperparams as hp
from lingvo_b200.utils import protowire as pw


class Color(enum.Enum):
  RED = 1
  BLUE = 2


Point = collections.namedtuple('Point', ['x', 'y'])


def _Tree():
  inner = hp.Params()
  inner.Define('alpha', 0.5, '')
  inner.Define('name', 'in"ner\nline', '')
  p = hp.Params()
  p.Define('an_int', -7, '')
  p.Define('a_big_int', 1 << 40, '')
  p.Define('a_bool', True, '')
  p.Define('a_float', 3.25, '')
  p.Define('a_precise_float', 0.1, '')
  p.Define('nothing', None, '')
  p.Define('a_list', [1, 'two', [3.0, None]], '')
  p.Define('a_tuple', (1, (2, 3)), '')
  p.Define('a_dict', {'k': 1, 'sub': inner.Copy()}, '')
  p.Define('a_dtype', torch.bfloat16, '')
  p.Define('np_dtype', np.dtype('int16'), '')
  p.Define('an_enum', Color.BLUE, '')
  p.Define('a_point', Point(1, 2.5), '')
  p.Define('a_class', collections.OrderedDict, '')
  p.Define('child', inner, '')
  p.Define('children', [inner.Copy().Set(alpha=1.0), inner.Copy().Set(alpha=2.0)], '')
  return p


def test_proto_round_trip_preserves_every_value_kind():
  p = _Tree()
  q = hp.Params.FromProto(p.ToProto())
  assert q.an_int == -7 and q.a_big_int == 1 << 40 and q.a_bool is True
  assert q.a_float == 3.25 and q.a_precise_float == 0.1 and q.nothing is None
  assert q.a_list == [1, 'two', [3.0, None]] and q.a_tuple == (1, (2, 3))
  assert q.a_dict['k'] == 1 and q.a_dict['sub'].alpha == 0.5
  assert q.a_dtype == torch.bfloat16 and q.np_dtype == np.dtype('int16')
  assert q.an_enum is Color.BLUE and q.a_point == Point(1, 2.5)
  assert q.a_class is collections.OrderedDict
  assert q.child.name == 'in"ner\nline'
  assert [c.alpha for c in q.children] == [1.0, 2.0]
  assert q.ToText() == p.ToText()


def test_proto_bytes_follow_the_hyperparam_schema():
  p = hp.Params()
  p.Define('x', 5, '')
  msg = pw.parse_dict(p.ToProto())
  assert list(msg) == [1]                      # Hyperparam.items (map<string, HyperparamValue>)
  entry = pw.parse_dict(msg[1][0])
  assert entry[1][0] == b'x'                   # map key
  value = pw.parse_dict(entry[2][0])
  assert value == {9: [5]}                     # HyperparamValue.int_val = 9


def test_instantiable_params_come_back_bound_to_their_class():
  from lingvo_b200.core import layers
  lp = layers.FCLayer.Params().Set(name='fc', input_dim=4, output_dim=3, activation='TANH')
  back = hp.Params.FromProto(lp.ToProto())
  assert isinstance(back, hp.InstantiableParams) and back.cls is layers.FCLayer
  layer = back.Instantiate()
  assert layer.FPropDefaultTheta(torch.zeros(2, 4)).shape == (2, 3)


def test_proto_text_is_valid_textformat_shape():
  text = _Tree().ToProtoText()
  assert text.count('{') == text.count('}')
  assert 'key: "an_enum"' in text and 'enum_val {' in text and 'name: "BLUE"' in text
  assert 'dtype_val: "bfloat16"' in text and 'int_val: -7' in text
  assert 'string_val: "in\\"ner\\nline"' in text
