"""Post-mortem debugging (ref `lingvo/pdb_wrapper.py`): `--pdb_on_exception` drops into
pdb at the point of an uncaught exception (main thread and runner threads)."""
import pdb
import sys
import threading
import traceback


def post_mortem(*args):  # pylint: disable=invalid-name
  traceback.print_exc()
  pdb.post_mortem(*args)


def _ExceptHook(exc_type, exc, tb):
  traceback.print_exception(exc_type, exc, tb)
  if not isinstance(exc, (KeyboardInterrupt, SystemExit)):
    pdb.post_mortem(tb)


def InstallOnException():
  sys.excepthook = _ExceptHook
  if hasattr(threading, 'excepthook'):
    threading.excepthook = lambda a: _ExceptHook(a.exc_type, a.exc_value, a.exc_traceback)
