"""ThreadSanitizer run of the native record pipeline.

Builds `csrc_host/records.cpp` + `tests/yielder_stress.cpp` with `-fsanitize=thread` into a
standalone executable and runs it; any data race makes TSAN exit non-zero.

  python tools/tsan_host.py            # build + run, prints TSAN_OK / the report
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'lingvo_b200', 'ops', 'csrc_host')


def Run(sanitizer='thread'):
  with tempfile.TemporaryDirectory() as tmp:
    exe = os.path.join(tmp, 'yielder_stress')
    cmd = ['g++', '-O1', '-g', '-std=c++17', '-pthread', '-fsanitize=' + sanitizer,
           '-fno-omit-frame-pointer', os.path.join(SRC, 'records.cpp'),
           os.path.join(SRC, 'tests', 'yielder_stress.cpp'), '-o', exe, '-lz']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      return False, 'build failed:\n' + r.stdout + r.stderr
    data = os.path.join(tmp, 'data')
    os.makedirs(data)
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=1 second_deadlock_stack=1')
    r = subprocess.run([exe, data], capture_output=True, text=True, timeout=600, env=env)
    ok = r.returncode == 0 and 'YIELDER_STRESS_OK' in r.stdout
    return ok, r.stdout + r.stderr


if __name__ == '__main__':
  ok, out = Run(sys.argv[1] if len(sys.argv) > 1 else 'thread')
  print(out[-4000:])
  print('TSAN_OK' if ok else 'TSAN_FAILED')
  sys.exit(0 if ok else 1)
