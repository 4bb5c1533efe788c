"""Cluster launcher (ref `lingvo/tools/gke_launch.py`, which builds GKE pod specs for TPU
trainers + GPU evalers + tensorboard).

On a B200 node the unit of deployment is `torchrun` with one process per GPU; this
tool prints / executes the launch command for trainer, evaler and decoder jobs that
share a `--logdir`.

  python -m lingvo_b200.tools.gke_launch --model=lm.x.Y --logdir=/logs/run1 \\
      --gpus=8 [--nnodes=1 --node_rank=0 --master_addr=…] [--dry_run] up|evaler|decoder
"""
import shlex
import subprocess
import sys

from absl import app
from absl import flags

flags.DEFINE_string('model', '', 'Registered model name.')
flags.DEFINE_string('logdir', '', 'Shared log directory.')
flags.DEFINE_integer('gpus', 8, 'GPUs (processes) per node.')
flags.DEFINE_integer('nnodes', 1, 'Number of nodes.')
flags.DEFINE_integer('node_rank', 0, 'Rank of this node.')
flags.DEFINE_string('master_addr', '127.0.0.1', 'Rendezvous address.')
flags.DEFINE_integer('master_port', 29500, 'Rendezvous port.')
flags.DEFINE_bool('dry_run', False, 'Only print the command.')
FLAGS = flags.FLAGS


def BuildCommand(action):
  base = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=%d' % FLAGS.nnodes,
          '--node-rank=%d' % FLAGS.node_rank, '--master-addr', FLAGS.master_addr,
          '--master-port', str(FLAGS.master_port)]
  if action == 'up':
    return base + ['--nproc-per-node', str(FLAGS.gpus), '-m', 'lingvo_b200.trainer',
                   '--model=' + FLAGS.model, '--logdir=' + FLAGS.logdir, '--mode=sync',
                   '--job=controller,trainer_client', '--worker_gpus=%d' % FLAGS.gpus]
  job = {'evaler': 'evaler_dev', 'decoder': 'decoder_dev'}[action]
  return [sys.executable, '-m', 'lingvo_b200.trainer', '--model=' + FLAGS.model,
          '--logdir=' + FLAGS.logdir, '--job=' + job]


def main(argv):
  action = argv[1] if len(argv) > 1 else 'up'
  cmd = BuildCommand(action)
  print(' '.join(shlex.quote(c) for c in cmd))
  if not FLAGS.dry_run:
    return subprocess.call(cmd)
  return 0


if __name__ == '__main__':
  app.run(main)
