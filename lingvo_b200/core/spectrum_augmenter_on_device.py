"""On-device SpecAugment (ref `lingvo/core/spectrum_augmenter_on_device.py`): the base
implementation is already batched tensor code, so this is the same layer."""
from lingvo_b200.core import spectrum_augmenter


class SpectrumAugmenterOnDevice(spectrum_augmenter.SpectrumAugmenter):
  pass
