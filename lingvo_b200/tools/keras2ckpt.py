"""Converts a Keras/HDF5-style or PyTorch state dict into a lingvo_b200 checkpoint
(ref `lingvo/tools/keras2ckpt.py`, which converts Keras weights to a TF checkpoint).

  python -m lingvo_b200.tools.keras2ckpt --input=weights.pt --output=/ckpt/ckpt-00000000 \\
      [--name_map=map.txt]      # lines "src_name dst_var_name"
"""
import numpy as np
import torch
from absl import app
from absl import flags

from lingvo_b200.utils import tensor_bundle

flags.DEFINE_string('input', '', '.pt / .npz file with named tensors.')
flags.DEFINE_string('output', '', 'Checkpoint prefix to write.')
flags.DEFINE_string('name_map', '', 'Optional "src dst" renaming file.')
FLAGS = flags.FLAGS


def Convert(inp, out, name_map=None):
  if inp.endswith('.npz'):
    with np.load(inp) as f:
      tensors = {k: f[k] for k in f.files}
  else:
    sd = torch.load(inp, map_location='cpu')
    sd = sd.get('state_dict', sd) if isinstance(sd, dict) else sd
    tensors = {k: v.detach().cpu().numpy() for k, v in sd.items() if isinstance(v, torch.Tensor)}
  name_map = name_map or {}
  renamed = {name_map.get(k, k): v for k, v in tensors.items()}
  w = tensor_bundle.BundleWriter(out)
  for name in sorted(renamed):
    w.Add(name, renamed[name])
  w.Finish()
  return sorted(renamed)


def main(argv):
  del argv
  nm = {}
  if FLAGS.name_map:
    with open(FLAGS.name_map) as f:
      nm = dict(l.split()[:2] for l in f if l.strip())
  names = Convert(FLAGS.input, FLAGS.output, nm)
  print('wrote %d tensors to %s' % (len(names), FLAGS.output))


if __name__ == '__main__':
  app.run(main)
