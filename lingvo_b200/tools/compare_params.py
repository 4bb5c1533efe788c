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


def main(argv):
  del argv
  for k, x, y in CompareParams(FLAGS.model1, FLAGS.model2, FLAGS.dataset):
    print('%s:\n  < %s\n  > %s' % (k, x, y))


if __name__ == '__main__':
  app.run(main)
