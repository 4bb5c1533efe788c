"""Test data paths (ref `lingvo/core/test_helper.py`)."""
import os


def test_src_dir_path(relative_path):  # pylint: disable=invalid-name
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  return os.path.join(root, relative_path)
