"""Shared flag definitions (reference `lingvo/trainer_utils.py:19-59`)."""

from lingvo_b200 import flags

flags.DEFINE_string('model', None,
                    'Name of the model class to train. Must be a model '
                    'defined in the model_registry.')
flags.DEFINE_string('model_task_name', '',
                    'For multitask models: select the task to run.')
flags.DEFINE_string('logdir', '', 'Log directory.')
flags.DEFINE_string('job', '',
                    'trainer/controller/evaler_<ds>/decoder_<ds>/executor_tpu, '
                    'comma separated to run several in one process.')
flags.DEFINE_integer('task', 0, 'Task id within the job (rank).')
flags.DEFINE_string('tf_master', '', 'Kept for parity (rendezvous address).')
flags.DEFINE_string('worker_job', '/job:trainer', 'Job name.')
flags.DEFINE_list('additional_worker_jobs', [], 'Additional worker job names.')
flags.DEFINE_integer('worker_tpus', 0, 'Kept for parity.')
flags.DEFINE_integer('worker_num_tpu_hosts', 0, 'Kept for parity.')
flags.DEFINE_string('evaler_job', '/job:evaler', 'Job name.')
flags.DEFINE_integer('evaler_replicas', 0, 'Number of replicas.')
flags.DEFINE_integer('evaler_gpus', 0, 'Number of gpus to use per replica.')
flags.DEFINE_string('decoder_job', '/job:decoder', 'Job name.')
flags.DEFINE_integer('decoder_replicas', 0, 'Number of replicas.')
flags.DEFINE_integer('decoder_gpus', 0, 'Number of gpus to use per replica.')
flags.DEFINE_integer('saver_max_to_keep', None, 'Overrides save_max_to_keep.')
flags.DEFINE_float('saver_keep_checkpoint_every_n_hours', None,
                   'Overrides save_keep_checkpoint_every_n_hours.')
flags.DEFINE_integer('enqueue_max_steps', None, 'Kept for parity.')
flags.DEFINE_bool('run_functions_eagerly', False, 'Kept for parity.')
flags.DEFINE_bool('enable_tf_data_debug_mode', False, 'Kept for parity.')
