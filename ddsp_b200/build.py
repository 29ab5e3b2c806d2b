"""Builds libddsp_b200.so in-tree with nvcc for sm_100a (B200) only."""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libddsp_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
]


def _sources():
  return [os.path.join(CSRC, 'capi.cu')]


def _newest_mtime():
  newest = 0.0
  for root in (CSRC, os.path.join(_HERE, '..', 'include')):
    for name in os.listdir(root):
      if name.endswith(('.cu', '.cuh', '.h', '.inc')):
        newest = max(newest, os.path.getmtime(os.path.join(root, name)))
  return newest


def is_stale():
  return (not os.path.exists(LIB_PATH) or
          os.path.getmtime(LIB_PATH) < _newest_mtime())


def build(force=False, verbose=False):
  """Compiles the CUDA library if missing or older than its sources."""
  if not force and not is_stale():
    return LIB_PATH
  nvcc = os.environ.get('NVCC', 'nvcc')
  extra = os.environ.get('DDSP_B200_NVCC_EXTRA', '').split()   # e.g. -DDDSP_HV3_MIN_CTAS=5
  cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + [
      '-o', LIB_PATH] + _sources()
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), proc.stderr))
  if verbose:
    sys.stderr.write(proc.stderr)
  return LIB_PATH


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
