"""Round-2 program / executor breadth: HostDrivenTrainProgram (reference program.py:771),
ExperimentalDecodeProgram (:1807), MultiTaskProgramSchedule (:2319), program-state save."""

import os
import tempfile

import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import program
from lingvo_b200.core import program_utils
import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa: F401

MODEL = 'lm.synthetic_packed_input.DenseLmTiny'


def _Cfg(max_steps=6):
  cfg = model_registry.GetParams(MODEL, 'Train')
  cfg.train.max_steps = max_steps
  cfg.task.train.max_steps = max_steps
  return cfg


def test_host_driven_train_program_runs_callbacks_every_step():
  logdir = tempfile.mkdtemp()
  p = program.HostDrivenTrainProgram.Params().Set(
      name='train', task=_Cfg(), logdir=logdir, dataset_name='Train', steps_per_loop=3,
      metrics_every_n=2)
  prog = p.Instantiate()
  prog.BuildTpuSubgraph()
  seen = []
  prog.AddStepCallback(lambda step, m: seen.append((step, None if m is None else m['loss'])))
  done = prog.Run()
  assert not done and [s for s, _ in seen] == [1, 2, 3]
  assert seen[0][1] is None and seen[1][1] is not None         # host metrics every 2nd step
  assert not prog.engine.cuda_graph
  done = prog.Run()
  assert done and prog.global_step == 6


def test_trigger_scheduler_state_round_trips_through_schedule(tmp_path):
  cfg = _Cfg()
  ps = program.SimpleProgramScheduleForTask('Train', 2, ['Train'], 1)
  ps.task_dict = {'Train': cfg}
  ps.logdir = str(tmp_path)
  ps.eval_program_triggers = {'Train': (1, 2)}
  sched = ps.Instantiate()
  for pr in sched.Programs():
    pr.BuildTpuSubgraph()
  sched.Run()
  sched.Run()
  path = os.path.join(str(tmp_path), 'program_state.json')
  sched.SaveProgramState(path)
  fresh = ps.Instantiate()
  assert fresh.LoadProgramState(path)
  trig = list(fresh._triggers.values())[0]
  assert trig.count == 2 and trig.State()['interval'] == 2
  assert isinstance(trig, program_utils.TriggerScheduler)


def test_multi_task_program_schedule_runs_the_sampled_task(tmp_path):
  a, b = _Cfg(4), _Cfg(4)
  pa = program.SimpleProgramScheduleForTask('Train', 2, [], 0)
  pa.task_dict = {'Train': a}
  pb = program.SimpleProgramScheduleForTask('Train', 1, [], 0)
  pb.task_dict = {'Train': b}
  mp = program.MultiTaskProgramSchedule.Params().Set(
      program_schedule_dict={'a': pa, 'b': pb}, logdir=str(tmp_path))
  sched = mp.Instantiate()
  assert sorted(sched.schedules) == ['a', 'b'] and len(sched.Programs()) == 2
  for pr in sched.Programs():
    pr._task_name = None                       # two independent single-task models here
    pr.BuildTpuSubgraph()
  sched.Run('a')
  sched.Run('b')
  sched.Run('b')
  assert sched.steps_run == {'a': 1, 'b': 2}
  assert sched.schedules['a'].train_program.global_step == 2
  assert sched.schedules['b'].train_program.global_step == 2
  done, _, _ = sched.Run()                     # round-robin over both
  assert done is True or done is False
  assert sched.schedules['a'].train_program.global_step == 4


def test_mlperf_program_schedule_trains_and_decodes_in_one_program(tmp_path):
  from lingvo_b200.core import trainer_test_utils
  cls = trainer_test_utils.RegisterIdentityRegressionModel('MlperfIdentity', max_train_steps=4)
  cfg = model_registry.GetParams('test.test.MlperfIdentity', 'Train')
  assert cfg.input is not None and cls is not None
  ps = program.MLPerfProgramScheduleForTask('Train', 2, 'Train', 1)
  assert isinstance(ps.cls, type) and issubclass(ps.cls, program.BaseProgramSchedule)
  assert ps.dataset_names == ['Train', 'Train']
  ps.task_dict = {'Train': cfg}
  ps.logdir = str(tmp_path)
  ps.ml_perf.Set(benchmark_name='lm', steps_per_epoch=2, decoder_metric_name='diff',
                 decoder_metric_success_threshold=1e9, max_steps_to_train=4)
  sched = ps.Instantiate()
  progs = sched.Programs()
  assert len(progs) == 1 and isinstance(progs[0], program.MLPerfTrainDecodeProgram)
  progs[0].BuildTpuSubgraph()
  done, train_s, eval_s = sched.Run()
  assert not done and train_s > 0 and eval_s == 0.0
  assert sched.train_program.global_step == 2
  done, _, _ = sched.Run()
  assert done                                           # max_steps_to_train reached
  sched.Shutdown()
  import pytest
  with pytest.raises(ValueError):
    bad = ps.Copy()
    bad.task_dict = {'Dev': cfg}
    bad.Instantiate()
  with pytest.raises(TypeError):
    program.BaseProgramSchedule()                       # abstract
