"""LibriSpeech data preparation (ref `lingvo/tasks/asr/tools/librispeech.0[1-4].*.sh`,
`librispeech_lib.sh`): download the OpenSLR tarballs and parameterise them into TFRecords
of log-mel frames + transcripts with `lingvo_b200.tools.create_asr_features`.

  python -m lingvo_b200.models.asr.tools.librispeech_get_data --root=/tmp/librispeech \
      [--sets=train-clean-100,dev-clean,test-clean] [--mirror=/local/tarballs]
"""

from __future__ import annotations

import argparse
import os
import shutil
import sys
import urllib.request

BASE_URL = 'http://www.openslr.org/resources/12/'
TRAIN = ['train-clean-100', 'train-clean-360', 'train-other-500']
DEVTEST = ['dev-clean', 'dev-other', 'test-clean', 'test-other']
SHARDS = {'train-clean-100': 10, 'train-clean-360': 36, 'train-other-500': 50}


def Download(root, name, mirror=''):
  dst = os.path.join(root, 'raw', name + '.tar.gz')
  os.makedirs(os.path.dirname(dst), exist_ok=True)
  if os.path.exists(dst):
    return dst
  if mirror and os.path.exists(os.path.join(mirror, name + '.tar.gz')):
    shutil.copy(os.path.join(mirror, name + '.tar.gz'), dst)
    return dst
  tmp = dst + '.part'
  with urllib.request.urlopen(BASE_URL + name + '.tar.gz') as r, open(tmp, 'wb') as f:   # noqa: S310
    shutil.copyfileobj(r, f, 1 << 20)
  os.replace(tmp, dst)
  return dst


def _DecodeFlac(path):
  """FLAC → 16-bit mono WAV bytes (soundfile if installed, else the `flac` / `ffmpeg` CLI)."""
  import io  # pylint: disable=g-import-not-at-top
  import subprocess  # pylint: disable=g-import-not-at-top
  try:
    import soundfile as sf  # pylint: disable=g-import-not-at-top
    data, rate = sf.read(path, dtype='int16')
    buf = io.BytesIO()
    sf.write(buf, data, rate, format='WAV', subtype='PCM_16')
    return buf.getvalue()
  except ImportError:
    pass
  for cmd in (['flac', '-d', '-c', '-s', path], ['ffmpeg', '-v', 'quiet', '-i', path, '-f', 'wav', '-']):
    if shutil.which(cmd[0]):
      return subprocess.run(cmd, check=True, capture_output=True).stdout   # noqa: S603
  raise RuntimeError('need the `soundfile` package or the flac / ffmpeg binary to decode ' + path)


def _Utterances(extracted_dir):
  """Walks `LibriSpeech/<set>/<speaker>/<chapter>/`: (utt id, transcript, wav bytes)."""
  for dirpath, _, files in sorted(os.walk(extracted_dir)):
    for tf_name in (f for f in files if f.endswith('.trans.txt')):
      with open(os.path.join(dirpath, tf_name), encoding='utf-8') as f:
        for line in f:
          uttid, text = line.strip().split(' ', 1)
          flac = os.path.join(dirpath, uttid + '.flac')
          if os.path.exists(flac):
            yield uttid, text, _DecodeFlac(flac)


def Parameterize(root, name, tarball):
  """Tarball → `<root>/<train|devtest>/<name>.tfrecords-*` of 80-dim log-mel features."""
  import tarfile  # pylint: disable=g-import-not-at-top
  from lingvo_b200.tools import create_asr_features  # pylint: disable=g-import-not-at-top
  extracted = os.path.join(root, 'extracted', name)
  if not os.path.exists(extracted):
    with tarfile.open(tarball) as t:
      t.extractall(extracted)   # noqa: S202
  sub = 'train' if name in TRAIN else 'devtest'
  out = os.path.join(root, sub)
  os.makedirs(out, exist_ok=True)
  shards = SHARDS.get(name, 1)
  n = create_asr_features.WriteFeatures(
      _Utterances(extracted), os.path.join(out, name + '.tfrecords-%5.5d-of-%5.5d'), shards)
  print('%s: %d utterances' % (name, n))


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--root', required=True)
  ap.add_argument('--sets', default=','.join(TRAIN + DEVTEST))
  ap.add_argument('--mirror', default='')
  ap.add_argument('--download_only', action='store_true')
  a = ap.parse_args(argv)
  for name in a.sets.split(','):
    marker = os.path.join(a.root, name + '.done')
    if os.path.exists(marker):
      print('[skip] %s' % name)
      continue
    tarball = Download(a.root, name, a.mirror)
    if not a.download_only:
      Parameterize(a.root, name, tarball)
    open(marker, 'w').close()
  return 0


if __name__ == '__main__':
  sys.exit(main())
