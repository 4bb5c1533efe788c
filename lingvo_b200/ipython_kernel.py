"""IPython kernel for interactive work with models (ref `lingvo/ipython_kernel.py`):
starts a kernel with `lingvo_b200` imported and all registered models loadable."""
import sys


def main(argv=None):
  try:
    from ipykernel import kernelapp  # pylint: disable=g-import-not-at-top
  except ImportError:
    print('ipykernel is not installed; falling back to an interactive console.')
    import code  # pylint: disable=g-import-not-at-top
    import lingvo_b200  # pylint: disable=g-import-not-at-top
    from lingvo_b200 import model_registry  # pylint: disable=g-import-not-at-top
    code.interact(local={'lingvo_b200': lingvo_b200, 'model_registry': model_registry})
    return 0
  import lingvo_b200.model_imports  # noqa: F401  pylint: disable=g-import-not-at-top
  kernelapp.launch_new_instance(argv=argv or sys.argv)
  return 0


if __name__ == '__main__':
  sys.exit(main())
