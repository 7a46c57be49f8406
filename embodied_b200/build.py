"""Compiles embodied_b200/csrc/*.cu into libembodied_b200.so (in-tree).

nvcc cross-compiles for sm_100a without a GPU.  The built .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import pathlib
import shutil
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / 'csrc'
LIB = ROOT / 'libembodied_b200.so'
SOURCES = ['abi.cu', 'rows.cu', 'rssm_fwd.cu', 'rssm_bwd.cu', 'norm.cu', 'optim.cu', 'spatial.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
]


def _nvcc():
  exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not pathlib.Path(exe).exists():
    raise RuntimeError('nvcc not found; cannot build libembodied_b200.so')
  return exe


def stale():
  if not LIB.exists():
    return True
  built = LIB.stat().st_mtime
  deps = list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh'))
  deps.append(ROOT.parent / 'include' / 'embodied_b200.h')
  return any(p.stat().st_mtime > built for p in deps)


def build(force=False, verbose=False):
  if not force and not stale():
    return LIB
  cmd = [_nvcc(), *NVCC_FLAGS]
  if verbose:
    cmd += ['-Xptxas', '-v']
  cmd += ['-o', str(LIB), *[str(CSRC / s) for s in SOURCES]]
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError(f'nvcc failed:\n{proc.stdout}\n{proc.stderr}')
  if verbose:
    print(proc.stderr)
  return LIB


if __name__ == '__main__':
  print(build(force=True, verbose=True))
