import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: test needs a CUDA (B200) device')
  config.addinivalue_line('markers', 'slow: long-running test')


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='needs a GPU')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
