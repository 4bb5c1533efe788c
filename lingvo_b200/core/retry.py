"""Retry decorator with exponential back-off (ref `lingvo/core/retry.py`)."""
import functools
import random
import time
import traceback

from absl import logging


def Retry(retry_value=Exception, max_retries=None, initial_delay_sec=1.0,
          delay_growth_factor=1.5, delay_growth_fuzz=0.1, max_delay_sec=60,
          initial_delay=None):
  """Retries the wrapped call on `retry_value` exceptions."""
  if initial_delay is not None:
    initial_delay_sec = initial_delay
  if max_retries is None:
    max_retries = 2 ** 30

  def _Decorator(func):

    @functools.wraps(func)
    def _Wrapper(*args, **kwargs):
      delay = initial_delay_sec
      for attempt in range(max_retries + 1):
        try:
          return func(*args, **kwargs)
        except retry_value as e:  # pylint: disable=broad-except
          if attempt >= max_retries:
            raise
          logging.warning('Retry %d of %s after %.2fs: %s\n%s', attempt + 1,
                          getattr(func, '__name__', func), delay, e,
                          traceback.format_exc(limit=3))
          time.sleep(delay)
          fuzz = 1.0 + delay_growth_fuzz * (2 * random.random() - 1)
          delay = min(delay * delay_growth_factor * fuzz, max_delay_sec)
    return _Wrapper
  return _Decorator
