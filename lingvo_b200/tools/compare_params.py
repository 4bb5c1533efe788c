"""Diffs the hyper-parameters of two registered models
(ref `lingvo/tools/compare_params.py`).

  python -m lingvo_b200.tools.compare_params --model1=lm.x.A --model2=lm.x.B [--dataset=Train]
"""
from absl import app
from absl import flags

from lingvo_b200 import model_registry

flags.DEFINE_string('model1', '', 'First model name.')
flags.DEFINE_string('model2', '', 'Second model name.')
flags.DEFINE_string('dataset', 'Train', 'Dataset whose params are compared.')
FLAGS = flags.FLAGS


def _Lines(name, dataset):
  p = model_registry.GetParams(name, dataset)
  out = {}
  for line in p.ToText().splitlines():
    if ' : ' in line:
      k, v = line.split(' : ', 1)
      out[k] = v
  return out


def CompareParams(model1, model2, dataset='Train'):
  a, b = _Lines(model1, dataset), _Lines(model2, dataset)
  diffs = []
  for k in sorted(set(a) | set(b)):
    if a.get(k) != b.get(k):
      diffs.append((k, a.get(k, '<missing>'), b.get(k, '<missing>')))
  return diffs


# -- text-level API (ref :22-118): works on `Params.ToText()` dumps, e.g. the `params.txt` a
#    trainer writes into its log dir, so two *runs* can be compared without their code --------
def _hyperparams_text_to_dict(cfg_text):   # pylint: disable=invalid-name
  out = {}
  for line in cfg_text.split('\n'):
    if not line:
      continue
    vals = line.split(' : ')
    if len(vals) != 2:
      raise ValueError(line)
    out[vals[0]] = vals[1]
  return out


def hyperparams_text_diff(cfg1_text, cfg2_text):   # pylint: disable=invalid-name
  """→ (keys only in cfg1, keys only in cfg2, {common key: (value1, value2)} where they
  differ); `.cls` values are compared by class name only (module paths may move)."""
  d1, d2 = _hyperparams_text_to_dict(cfg1_text), _hyperparams_text_to_dict(cfg2_text)
  only1, only2 = sorted(set(d1) - set(d2)), sorted(set(d2) - set(d1))
  cls_name = lambda v: v[v.rindex('/') + 1:] if '/' in v else v
  diff = {}
  for k in set(d1) & set(d2):
    v1, v2 = d1[k], d2[k]
    if k.endswith('.cls'):
      v1, v2 = cls_name(v1), cls_name(v2)
    if v1 != v2:
      diff[k] = (v1, v2)
  return only1, only2, diff


def print_hyperparams_text_diff(path1, path2, cfg1_not_cfg2, cfg2_not_cfg1,   # pylint: disable=invalid-name
                                cfg1_and_cfg2_diff):
  if cfg1_not_cfg2:
    print('\n\nKeys in %s but not %s:' % (path1, path2))
    for k in cfg1_not_cfg2:
      print('  %s' % k)
  if cfg2_not_cfg1:
    print('\n\nKeys in %s but not %s:' % (path2, path1))
    for k in cfg2_not_cfg1:
      print('  %s' % k)
  if cfg1_and_cfg2_diff:
    print('\n\nKeys with differences and their values: \n\n')
    for k, (v1, v2) in sorted(cfg1_and_cfg2_diff.items()):
      print('%s:\n      [%s]\n  vs. [%s]' % (k, v1, v2))
    print('\n\n')


def get_model_params_as_text(model_path, dataset='Train'):   # pylint: disable=invalid-name
  """`model_path`: a registered model name, or a file holding a `ToText()` dump."""
  try:
    return model_registry.GetParams(model_path, dataset).ToText()
  except LookupError:
    with open(model_path) as f:
      return f.read()


def main(argv):
  del argv
  for k, x, y in CompareParams(FLAGS.model1, FLAGS.model2, FLAGS.dataset):
    print('%s:\n  < %s\n  > %s' % (k, x, y))


if __name__ == '__main__':
  app.run(main)
