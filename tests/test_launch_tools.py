"""Deployment tooling: Kubernetes launcher manifests and the wheel builder
(ref `lingvo/tools/gke_launch.py`, `pip_package/`)."""

import os
import zipfile

import yaml

from lingvo_b200.tools import gke_launch


def _Args(*extra):
  return ['--name=moe', '--model=lm.synthetic_packed_input.MoELm8E', '--image=reg/lb:1',
          '--logdir=/mnt/logs/moe', '--nodes=2', '--decoder=dev,test', '--evaler=dev',
          '--extra_envs=NCCL_DEBUG_SUBSYS=INIT,FOO=bar', *extra]


def test_print_writes_valid_manifests(capsys):
  calls = []
  rc, paths = gke_launch.Main(_Args('print'), run=lambda cmd: calls.append(cmd) or 0)
  assert rc == 0 and not calls                                  # print never calls kubectl
  assert sorted(paths) == ['decoder.yaml', 'evaler.yaml', 'tensorboard.yaml', 'trainer.yaml']
  svc, job = list(yaml.safe_load_all(open(paths['trainer.yaml'])))
  assert svc['kind'] == 'Service' and svc['spec']['clusterIP'] == 'None'
  assert job['kind'] == 'Job' and job['spec']['completionMode'] == 'Indexed'
  assert job['spec']['completions'] == job['spec']['parallelism'] == 2
  pod = job['spec']['template']['spec']
  c = pod['containers'][0]
  assert c['resources']['limits']['nvidia.com/gpu'] == 8 and pod['hostIPC'] is True
  assert {'name': 'dshm', 'mountPath': '/dev/shm'} in c['volumeMounts']
  assert pod['volumes'][0]['emptyDir']['medium'] == 'Memory'
  cmd = c['command'][-1]
  assert '--nnodes=2' in cmd and '--nproc-per-node=8' in cmd
  assert '--node-rank=$(JOB_COMPLETION_INDEX)' in cmd
  assert '--master-addr=moe-trainer-0.moe-trainer' in cmd
  assert '-m lingvo_b200.trainer --model=lm.synthetic_packed_input.MoELm8E' in cmd
  assert {'name': 'FOO', 'value': 'bar'} in c['env']
  decs = list(yaml.safe_load_all(open(paths['decoder.yaml'])))
  assert [d['metadata']['name'] for d in decs] == ['moe-decoder-dev', 'moe-decoder-test']
  dc = decs[1]['spec']['template']['spec']['containers'][0]
  assert '--job=decoder_test' in dc['command']
  assert dc['resources']['limits']['nvidia.com/gpu'] == 1
  assert dc['resources']['requests']['memory'] == '24G'
  tb_dep, tb_svc = list(yaml.safe_load_all(open(paths['tensorboard.yaml'])))
  assert tb_svc['spec']['type'] == 'LoadBalancer'
  assert '--logdir=/mnt/logs/moe' in tb_dep['spec']['template']['spec']['containers'][0]['command']
  assert 'kind: Job' in capsys.readouterr().out


def test_up_down_reload_and_single_target_call_kubectl():
  calls = []
  run = lambda cmd: calls.append(cmd) or 0
  gke_launch.Main(_Args('--cluster=b200-pool', 'reload', 'trainer'), run=run)
  verbs = [c[1] for c in calls]
  assert verbs == ['delete', 'create'] and all(c[0] == 'kubectl' for c in calls)
  assert all(c[-2:] == ['--context', 'b200-pool'] and c[3].endswith('trainer.yaml')
             for c in calls)
  calls.clear()
  gke_launch.Main(_Args('--nodes=1', 'up', 'tensorboard'), run=run)
  assert len(calls) == 1 and calls[0][1] == 'create'
  # single node: rendezvous on localhost
  cmd = gke_launch.TrainerCommand(gke_launch.ParseArgs(_Args('--nodes=1', 'print')))
  assert '--master-addr=127.0.0.1' in cmd


def test_build_pushes_a_timestamped_image():
  calls = []
  gke_launch.Main(['--name=x', '--model=m', '--image=reg/lb', '--logdir=/l', '--build=/src',
                   'up', 'tensorboard'], run=lambda cmd: calls.append(cmd) or 0)
  assert calls[0][:3] == ['docker', 'build', '-t'] and calls[0][3].startswith('reg/lb:20')
  assert 'BASE_IMAGE=nvidia/cuda:12.9.0-devel-ubuntu24.04' in calls[0]
  assert calls[1][:2] == ['docker', 'push'] and calls[2][0] == 'kubectl'


def test_wheel_builder_makes_an_installable_archive(tmp_path):
  import importlib.util
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location(
      'build_pip_pkg', os.path.join(root, 'pip_package', 'build_pip_pkg.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  whl = mod.BuildWheel(str(tmp_path), skip_native=True)
  assert whl.endswith('lingvo_b200-0.2.0-py3-none-any.whl')
  with zipfile.ZipFile(whl) as z:
    names = set(z.namelist())
    assert 'lingvo_b200/trainer.py' in names and 'lingvo_b200/ops/csrc/flash_attn.cu' in names
    assert 'lingvo_b200-0.2.0.dist-info/RECORD' in names
    record = z.read('lingvo_b200-0.2.0.dist-info/RECORD').decode().splitlines()
    assert len(record) == len(names)
    assert b'lingvo_b200.trainer:main_cli' in z.read(
        'lingvo_b200-0.2.0.dist-info/entry_points.txt')
    assert not any(n.endswith('.so') for n in names)              # --skip-native


def test_compare_params_text_diff_and_stats_collector(capsys, tmp_path):
  import numpy as np
  from lingvo_b200.tools import compare_params
  from lingvo_b200.tools import compute_stats
  a = 'x.cls : type/old.module/Layer\nx.dim : 4\nonly_a : 1\n'
  b = 'x.cls : type/new.module/Layer\nx.dim : 8\nonly_b : 2\n'
  only_a, only_b, diff = compare_params.hyperparams_text_diff(a, b)
  assert only_a == ['only_a'] and only_b == ['only_b'] and diff == {'x.dim': ('4', '8')}
  compare_params.print_hyperparams_text_diff('A', 'B', only_a, only_b, diff)
  out = capsys.readouterr().out
  assert 'Keys in A but not B' in out and 'x.dim:' in out and 'vs. [8]' in out
  f = tmp_path / 'params.txt'
  f.write_text(a)
  assert compare_params.get_model_params_as_text(str(f)) == a
  sc = compute_stats.StatsCollector('frames', frame_size=2, num_buckets=4)
  rng = np.random.RandomState(0)
  allf = []
  for n in range(1, 41):
    x = rng.randn(n, 2) * [1.0, 3.0] + [5.0, -2.0]
    allf.append(x)
    sc.Accumulate({'frames': x.reshape(-1).astype(np.float32)})
  mean, std = sc.MeanVar()
  allf = np.concatenate(allf)
  np.testing.assert_allclose(mean, allf.mean(0), rtol=1e-5)
  np.testing.assert_allclose(std, allf.std(0), rtol=1e-4)
  buckets, alt = sc.LengthBuckets()
  assert buckets == [11, 21, 31, 40] and alt[0.02] == 40
  sc.Print()
  assert 'bucket upper limits' in capsys.readouterr().out
