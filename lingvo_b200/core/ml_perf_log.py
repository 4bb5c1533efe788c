"""MLPerf logging (ref `lingvo/core/ml_perf_log.py`)."""
import inspect
import json
import os
import time

from absl import logging

PREFIX = ':::MLL'


def get_caller(stack_index=2, root_dir=None):  # pylint: disable=invalid-name
  caller = inspect.getframeinfo(inspect.stack()[stack_index][0])
  filename = caller.filename
  if root_dir is not None:
    filename = os.path.relpath(filename, root_dir)
  return '%s:%d' % (filename, caller.lineno)


def mlperf_format(key, value, now, stack_offset=0, metadata=None):  # pylint: disable=invalid-name
  meta = {'lineno': get_caller(3 + stack_offset)}
  meta.update(metadata or {})
  return '%s %s' % (PREFIX, json.dumps({'namespace': '', 'time_ms': int(now * 1e3),
                                        'event_type': 'POINT_IN_TIME', 'key': key,
                                        'value': value, 'metadata': meta}))


def mlperf_print(key, value, stack_offset=0, metadata=None):  # pylint: disable=invalid-name
  logging.info(mlperf_format(key, value, time.time(), stack_offset, metadata))
