"""Input placement policy (ref `lingvo/core/input_policy.py`): the input generator's
tensors are produced on the cluster's input device (host), never on the GPU."""


def Apply(input_params):

  class _UseInputDevice(input_params.cls):
    """Keeps generated batches on the host; `DevicePrefetcher` moves them later."""

    def SplitInputBatch(self, num_splits):
      return super().SplitInputBatch(num_splits)

  _UseInputDevice.__name__ = input_params.cls.__name__
  return input_params.Copy().Set(cls=_UseInputDevice)
