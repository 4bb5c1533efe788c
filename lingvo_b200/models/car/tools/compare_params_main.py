"""Diffs the hyper-parameters of two registered car models (ref
`lingvo/tasks/car/tools/compare_params_main.py`).

  python -m lingvo_b200.models.car.tools.compare_params_main \
      car.kitti.StarNetCarModel0701 car.kitti.StarNetPedCycModel0704 [Train]
"""

import sys

from lingvo_b200.tools import compare_params


def main(argv=None):
  argv = list(sys.argv if argv is None else argv)
  if len(argv) < 3:
    print(__doc__)
    return 1
  import lingvo_b200.models.car.params.params  # noqa: F401  pylint: disable=g-import-not-at-top
  dataset = argv[3] if len(argv) > 3 else 'Train'
  for k, x, y in compare_params.CompareParams(argv[1], argv[2], dataset):
    print('%s:\n  < %s\n  > %s' % (k, x, y))
  return 0


if __name__ == '__main__':
  sys.exit(main())
