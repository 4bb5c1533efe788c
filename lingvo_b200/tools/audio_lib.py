"""Audio helpers (ref `lingvo/tools/audio_lib.py`): decode WAV bytes and compute
log-mel filterbanks with the task's `MelAsrFrontend`."""
import io
import wave

import numpy as np
import torch

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.asr import frontend as asr_frontend


def DecodeWav(wav_bytes):
  """→ (sample_rate, float32 [samples] in 16-bit units)."""
  with wave.open(io.BytesIO(wav_bytes), 'rb') as w:
    sr, n, width, ch = w.getframerate(), w.getnframes(), w.getsampwidth(), w.getnchannels()
    raw = w.readframes(n)
  assert width == 2, 'only 16-bit PCM is supported'
  pcm = np.frombuffer(raw, np.int16).astype(np.float32)
  if ch > 1:
    pcm = pcm.reshape(-1, ch).mean(1)
  return sr, pcm


def ExtractLogMelFeatures(wav_bytes, num_bins=80, sample_rate=16000.0):
  sr, pcm = DecodeWav(wav_bytes)
  p = asr_frontend.MelAsrFrontend.Params().Set(sample_rate=float(sr or sample_rate),
                                               num_bins=num_bins, noise_scale=0.0)
  fe = p.Instantiate()
  x = torch.from_numpy(pcm).unsqueeze(0)
  out = fe.FPropDefaultTheta(NestedMap(src_inputs=x, paddings=torch.zeros_like(x)))
  return out.src_inputs[0, :, :, 0].numpy()
