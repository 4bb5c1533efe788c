r"""Trainer binary.

  python -m lingvo_b200.trainer --run_locally=cpu --mode=sync \
      --model=image.mnist.LeNet5 --logdir=/tmp/lenet5

Reference `lingvo/trainer.py`: flags (:54-205), `RunnerManager` (:224) with
`MaybeConfigRunLocally` (:628), `MaybeConfigRunDistributed` (:329),
`UpdateClusterParamsFromFlags` (:450), `_CreateRunner` (:506),
`StartRunners` (:575), inspect modes (`inspect_params`, `inspect_model`,
`inspect_evaler`, `inspect_decoder`), `write_inference_graph` (:764).

Distributed launch is `torchrun` (one process per GPU): RANK / WORLD_SIZE /
LOCAL_RANK are read from the environment and `--cluster_spec` is accepted
only for flag parity.
"""

from __future__ import annotations

import logging
import os
import re
import sys
import threading
import time

import torch

from lingvo_b200 import base_trial
from lingvo_b200 import datasets
from lingvo_b200 import executor
from lingvo_b200 import flags
from lingvo_b200 import model_imports
from lingvo_b200 import model_registry
from lingvo_b200 import runners
from lingvo_b200 import trainer_utils  # pylint: disable=unused-import
from lingvo_b200.core import base_model
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import py_utils
from lingvo_b200.core import summary_utils

flags.DEFINE_bool('interactive', False, 'Kept for parity.')
flags.DEFINE_string('run_locally', '', 'cpu|gpu|tpu: run everything locally.')
flags.DEFINE_string('mode', 'async', 'sync|async|shell|inspect_params|'
                    'inspect_model|inspect_evaler|inspect_decoder|'
                    'write_inference_graph.')
flags.DEFINE_string('cluster_spec', '', 'Kept for parity (use torchrun).')
flags.DEFINE_string('inspect_model_part_regex', None, 'Filter inspect_model.')
flags.DEFINE_integer('inspect_model_topn', 0, 'Top-n vars for inspect_model.')
flags.DEFINE_string('controller_job', '/job:controller', 'Job name.')
flags.DEFINE_integer('controller_gpus', 0, 'Number of controller GPUs.')
flags.DEFINE_integer('worker_replicas', 1, 'Number of replicas.')
flags.DEFINE_integer('worker_gpus', 0, 'Number of gpus to use per replica.')
flags.DEFINE_integer('worker_split_size', 1, 'Devices for one split.')
flags.DEFINE_string('ps_job', '/job:ps', 'Kept for parity.')
flags.DEFINE_integer('ps_replicas', 1, 'Kept for parity.')
flags.DEFINE_integer('ps_gpus', 0, 'Kept for parity.')
flags.DEFINE_string('input_job', '/job:input', 'Job name.')
flags.DEFINE_integer('input_replicas', 0, 'Number of replicas.')
flags.DEFINE_string('input_targets', '', 'Input server targets.')
flags.DEFINE_string('tf_data_service_address', '', 'Input service address.')
flags.DEFINE_string('inference_graph_filename', None, 'Output filename.')
flags.DEFINE_string('inference_graph_device', None, 'cpu|gpu.')
flags.DEFINE_integer('inference_graph_random_seed', None, 'Seed.')
flags.DEFINE_string('inspect_params_dataset_name', None, 'Dataset to inspect.')
flags.DEFINE_string('inference_dataset_name', 'Test', 'Dataset for export.')
flags.DEFINE_bool('evaler_in_same_address_as_controller', False, 'Kept.')
flags.DEFINE_string('vizier_reporting_job', 'evaler', 'Kept for parity.')
flags.DEFINE_bool('add_summary', None, 'Overrides cluster.add_summary.')
flags.DEFINE_bool('use_eager', True, 'Always eager here; kept for parity.')
flags.DEFINE_bool('checkpoint_in_trainer_tpu', False, 'Kept for parity.')
flags.DEFINE_bool('pdb_on_exception', False, 'Post-mortem debugger.')
flags.DEFINE_string('use_cuda_graph', None, "auto|on|off: overrides train.use_cuda_graph "
                    '(whole-step CUDA-graph replay in the trainer / executor).')

FLAGS = flags.FLAGS


class RunnerManager:
  """Creates and runs the jobs named by --job."""

  Controller = runners.Controller
  Trainer = runners.Trainer
  TrainerTpu = runners.TrainerTpu
  Evaler = runners.Evaler
  Decoder = runners.Decoder
  ExecutorTpu = executor.ExecutorTpu

  def __init__(self, model):
    self._model_name = model

  # ---------------------------------------------------------------- config --
  def MaybeConfigRunLocally(self):
    """Rewrites flags so all jobs run in this process (reference :628)."""
    if not FLAGS.run_locally:
      return
    if not FLAGS.mode:
      FLAGS.mode = 'sync'
    if not FLAGS.job:
      FLAGS.job = 'controller,trainer_client' if FLAGS.mode == 'sync' else (
          'controller,trainer')
    FLAGS.task = 0
    FLAGS.controller_job = '/job:localhost'
    FLAGS.worker_job = '/job:localhost'
    FLAGS.worker_replicas = 1
    if FLAGS.run_locally == 'gpu':
      if not FLAGS.worker_gpus:
        FLAGS.worker_gpus = 1
    else:
      FLAGS.worker_gpus = 0
    FLAGS.worker_split_size = max(1, FLAGS.worker_split_size)
    FLAGS.ps_job = '/job:localhost'
    FLAGS.ps_replicas = 1
    FLAGS.evaler_job = '/job:localhost'
    FLAGS.evaler_replicas = 1
    FLAGS.evaler_gpus = 1 if FLAGS.run_locally == 'gpu' else 0
    FLAGS.decoder_job = '/job:localhost'
    FLAGS.decoder_replicas = 1
    FLAGS.decoder_gpus = 1 if FLAGS.run_locally == 'gpu' else 0

  def MaybeConfigRunDistributed(self):
    """Initialises torch.distributed from torchrun's environment (:329)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
      return
    import torch.distributed as dist
    if not dist.is_initialized():
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
      if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
      dist.init_process_group(backend)
    FLAGS.task = int(os.environ.get('RANK', '0'))
    FLAGS.worker_replicas = world
    if torch.cuda.is_available() and not FLAGS.worker_gpus:
      FLAGS.worker_gpus = 1

  def UpdateClusterParamsFromFlags(self, cluster, job_name):
    """Fills cluster params from flags (reference :450-504)."""
    cluster.mode = FLAGS.mode if FLAGS.mode in ('sync', 'async') else 'sync'
    cluster.job = job_name
    cluster.task = FLAGS.task
    cluster.logdir = FLAGS.logdir
    cluster.controller.name = FLAGS.controller_job
    cluster.controller.gpus_per_replica = FLAGS.controller_gpus
    cluster.worker.name = FLAGS.worker_job
    cluster.worker.replicas = FLAGS.worker_replicas
    cluster.worker.gpus_per_replica = FLAGS.worker_gpus
    cluster.worker.devices_per_split = FLAGS.worker_split_size
    cluster.ps.name = FLAGS.ps_job
    cluster.ps.replicas = FLAGS.ps_replicas
    cluster.ps.gpus_per_replica = FLAGS.ps_gpus
    cluster.input.name = FLAGS.input_job
    cluster.input.replicas = FLAGS.input_replicas
    cluster.input.targets = FLAGS.input_targets
    cluster.evaler.name = FLAGS.evaler_job
    cluster.evaler.replicas = FLAGS.evaler_replicas
    cluster.evaler.gpus_per_replica = FLAGS.evaler_gpus
    cluster.decoder.name = FLAGS.decoder_job
    cluster.decoder.replicas = FLAGS.decoder_replicas
    cluster.decoder.gpus_per_replica = FLAGS.decoder_gpus
    cluster.tf_data_service_address = FLAGS.tf_data_service_address
    cluster.add_summary = FLAGS.add_summary

  def _ClusterJobName(self, job):
    if job in ('controller',):
      return 'controller'
    if job in ('trainer', 'trainer_client', 'worker'):
      return job
    if job.startswith('evaler'):
      return 'evaler'
    if job.startswith('decoder'):
      return 'decoder'
    if job.startswith('executor'):
      return 'executor_tpu'
    return job

  def GetParamsForDataset(self, job_name, dataset_name):
    """Model params for `dataset_name` with the cluster filled in (:301)."""
    try:
      cfg = model_registry.GetParams(self._model_name, dataset_name)
    except Exception as e:  # pylint: disable=broad-except
      # Dataset names are case-insensitive on the command line.
      cls = model_registry.GetClass(self._model_name)
      all_ds = datasets.GetDatasets(cls)
      match = [d for d in all_ds if d.lower() == dataset_name.lower()]
      if not match:
        raise
      cfg = model_registry.GetParams(self._model_name, match[0])
    self.UpdateClusterParamsFromFlags(cfg.cluster, self._ClusterJobName(job_name))
    task_trains = ([cfg.task.train] if 'task' in cfg else
                   [t.train for _, t in cfg.task_params.IterParams()])
    if FLAGS.use_cuda_graph:
      for tp in task_trains:
        tp.use_cuda_graph = FLAGS.use_cuda_graph
    if FLAGS['mode'].present and FLAGS.mode == 'async' and int(
        os.environ.get('WORLD_SIZE', '1')) > 1:
      # explicit `--mode=async` under torchrun: asynchronous data parallelism
      for tp in task_trains:
        tp.async_data_parallel = True
    if FLAGS.saver_max_to_keep is not None:
      cfg.train.save_max_to_keep = FLAGS.saver_max_to_keep
    if FLAGS.saver_keep_checkpoint_every_n_hours is not None:
      cfg.train.save_keep_checkpoint_every_n_hours = (
          FLAGS.saver_keep_checkpoint_every_n_hours)
    return cfg

  # --------------------------------------------------------------- runners --
  def _CreateRunner(self, job, model_task_name, logdir, tf_master, trial):
    evaler_prefix, decoder_prefix = 'evaler_', 'decoder_'
    trial = trial or base_trial.NoOpTrial()
    if job == 'controller':
      cfg = self.GetParamsForDataset('controller', 'Train')
      cfg.cluster.xla_device = 'cpu'
      return self.Controller(cfg, model_task_name, logdir, tf_master, trial)
    if job in ('trainer', 'trainer_client', 'worker'):
      cfg = self.GetParamsForDataset(job, 'Train')
      return self.Trainer(cfg, model_task_name, logdir, tf_master, trial)
    if job.startswith(evaler_prefix):
      ds = job[len(evaler_prefix):]
      cfg = self.GetParamsForDataset('evaler', ds.title())
      return self.Evaler(ds.lower(), cfg, model_task_name, logdir, tf_master,
                         trial)
    if job.startswith(decoder_prefix):
      ds = job[len(decoder_prefix):]
      cfg = self.GetParamsForDataset('decoder', ds.title())
      return self.Decoder(ds.lower(), cfg, model_task_name, logdir, tf_master,
                          trial)
    if job in ('executor_tpu', 'executor', 'host_driven_executor'):
      cluster_p = cluster_factory.Cluster.Params()
      self.UpdateClusterParamsFromFlags(cluster_p, 'executor_tpu')
      ps_dict, train_cfg = executor.GetExecutorParams(
          self._model_name, cluster_p, model_registry)
      return self.ExecutorTpu(train_cfg, ps_dict, model_task_name, logdir,
                              tf_master, trial)
    raise ValueError('job %s is not supported' % job)

  def CreateRunners(self, jobs, logdir, trial=None):
    runners_list = []
    for j in jobs:
      tf_master = FLAGS.tf_master
      runners_list.append(self._CreateRunner(j, FLAGS.model_task_name, logdir,
                                             tf_master, trial))
    return runners_list

  def StartRunners(self, runners_list):
    """One thread per runner; waits for all (reference :575-620)."""
    trainers = [r for r in runners_list if isinstance(r, runners.Trainer)]
    for r in runners_list:
      if isinstance(r, runners.Controller) and trainers:
        r._peer_done = lambda t=trainers: all(x.done() for x in t)  # pylint: disable=protected-access
    errors = []

    def run(r):
      try:
        r.Start()
      except BaseException as e:  # pylint: disable=broad-except
        errors.append(e)
        for other in runners_list:
          other.RequestStop()

    threads = []
    if len(runners_list) == 1:
      run(runners_list[0])
    else:
      for r in runners_list:
        t = threading.Thread(target=run, args=(r,), daemon=True,
                             name=type(r).__name__)
        t.start()
        threads.append(t)
      for t in threads:
        while t.is_alive():
          t.join(0.5)
    if errors:
      raise errors[0]

  def RunTrial(self, job, logdir, trial):
    self.StartRunners(self.CreateRunners([job], logdir, trial))

  # -------------------------------------------------------------- inspect --
  def InspectParams(self):
    """`--mode=inspect_params`: prints the params of a dataset (:687)."""
    FLAGS.mode = 'sync'
    cls = model_registry.GetClass(self._model_name)
    tf_dataset = FLAGS.inspect_params_dataset_name
    if tf_dataset:
      names = [tf_dataset]
    else:
      names = datasets.GetDatasets(cls)
    if not names:
      names = ['Train']
    cfg = self.GetParamsForDataset('controller', names[0])
    print(cfg.ToText())
    return cfg

  def InspectDatasets(self):
    cls = model_registry.GetClass(self._model_name)
    print(','.join([d.lower() for d in datasets.GetDatasets(cls)]))

  def InspectModel(self):
    """`--mode=inspect_model`: variable table (reference :702-735)."""
    FLAGS.mode = 'sync'
    p = self.GetParamsForDataset('controller', 'Train')
    c = cluster_factory.Cluster(p.cluster)
    with c, py_utils.StubVariablesScope('zeros'):
      model = p.Instantiate()
    text, _ = summary_utils.ModelAnalysis(model)
    if FLAGS.inspect_model_part_regex:
      keep = re.compile(FLAGS.inspect_model_part_regex)
      text = '\n'.join(l for l in text.split('\n') if keep.search(l))
    print(text)
    return text

  def InspectDecoder(self):
    self._InspectJobs('decoder_')

  def InspectEvaler(self):
    self._InspectJobs('evaler_')

  def _InspectJobs(self, prefix):
    cls = model_registry.GetClass(self._model_name)
    for ds in datasets.GetDatasets(cls):
      print(prefix + ds.lower())

  def WriteInferenceGraph(self, cfg=None, prune_graph=True):
    """`--mode=write_inference_graph` (reference :764-866)."""
    from lingvo_b200.core import inference_graph_exporter
    inference_graph_dir = os.path.join(FLAGS.logdir, 'inference_graphs')
    os.makedirs(inference_graph_dir, exist_ok=True)
    if not cfg:
      cfg = self.GetParamsForDataset('controller', FLAGS.inference_dataset_name)
    filename = FLAGS.inference_graph_filename or 'inference.pbtxt'
    path = os.path.join(inference_graph_dir, filename)
    inference_graph_exporter.InferenceGraphExporter.Export(
        model_cfg=cfg, model_task_name=FLAGS.model_task_name,
        device_options=inference_graph_exporter.InferenceDeviceOptions(
            device=FLAGS.inference_graph_device or 'cpu',
            retain_device_placement=False, var_options=None, gen_init_op=True,
            dtype_override=None, fprop_dtype_override=None),
        export_path=path, random_seed=FLAGS.inference_graph_random_seed)
    return path

  # ----------------------------------------------------------------- start --
  def Start(self):
    """Parses jobs and runs them (reference :879-922)."""
    if FLAGS.mode == 'inspect_params':
      self.InspectParams()
      return
    if FLAGS.mode == 'inspect_datasets':
      self.InspectDatasets()
      return
    if FLAGS.mode == 'inspect_model':
      self.InspectModel()
      return
    if FLAGS.mode == 'inspect_evaler':
      self.InspectEvaler()
      return
    if FLAGS.mode == 'inspect_decoder':
      self.InspectDecoder()
      return
    if FLAGS.mode == 'write_inference_graph':
      self.WriteInferenceGraph()
      return
    assert FLAGS.mode in ('sync', 'async', 'shell'), FLAGS.mode
    self.MaybeConfigRunLocally()
    self.MaybeConfigRunDistributed()
    if FLAGS.mode == 'shell':
      import code
      code.interact(local={'manager': self, 'FLAGS': FLAGS})
      return
    assert FLAGS.job, '--job is required'
    assert FLAGS.logdir, '--logdir is required'
    os.makedirs(FLAGS.logdir, exist_ok=True)
    jobs = [j for j in FLAGS.job.split(',') if j]
    self.StartRunners(self.CreateRunners(jobs, FLAGS.logdir))


def main(argv=None):
  logging.basicConfig(
      level=logging.INFO,
      format='I%(asctime)s %(threadName)s %(filename)s:%(lineno)d] %(message)s')
  FLAGS(sys.argv if argv is None else argv)
  assert FLAGS.model, '--model is required'
  model_imports.ImportParams(FLAGS.model)
  if FLAGS.pdb_on_exception:
    from lingvo_b200 import pdb_wrapper
    pdb_wrapper.InstallOnException()
  RunnerManager(FLAGS.model).Start()


def main_cli():
  """Console-script entry point (`lingvo_b200_trainer`, see setup.py / pip_package)."""
  main(sys.argv)


if __name__ == '__main__':
  main()
