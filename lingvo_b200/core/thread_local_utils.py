"""Thread-local containers (ref `lingvo/core/thread_local_utils.py`)."""
import threading


class ThreadLocalStack(threading.local):

  def __init__(self):
    super().__init__()
    self.stack = []


class ThreadLocalDict(threading.local):

  def __init__(self):
    super().__init__()
    self.dict = {}
