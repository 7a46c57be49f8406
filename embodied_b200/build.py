"""Compiles embodied_b200/csrc/*.cu into libembodied_b200.so (in-tree).

nvcc cross-compiles for sm_100a without a GPU.  Every translation unit is
compiled to an object file in parallel (only the stale ones), then linked.  The
built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import pathlib
import shutil
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / 'csrc'
OBJ = ROOT / 'build'
LIB = ROOT / 'libembodied_b200.so'
SOURCES = ['abi.cu', 'rows.cu', 'rssm_fwd.cu', 'rssm_fwd_tma.cu', 'rssm_bwd.cu', 'rssm_bwd_tma.cu', 'norm.cu', 'optim.cu', 'spatial.cu', 'kl.cu', 'probe.cu', 'thinconv.cu', 'gru.cu', 'pack.cu', 'conv_tc.cu', 'losses.cu', 'ppo.cu', 'noise.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC',
]


def _nvcc():
  exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not pathlib.Path(exe).exists():
    raise RuntimeError('nvcc not found; cannot build libembodied_b200.so')
  return exe


def _headers():
  return list(CSRC.glob('*.cuh')) + [ROOT.parent / 'include' / 'embodied_b200.h']


def stale():
  if not LIB.exists():
    return True
  built = LIB.stat().st_mtime
  deps = [CSRC / s for s in SOURCES] + _headers()
  return any(p.stat().st_mtime > built for p in deps)


def _compile(src, force, verbose):
  obj = OBJ / (src[:-3] + '.o')
  deps = [CSRC / src] + _headers()
  if not force and obj.exists() and all(p.stat().st_mtime <= obj.stat().st_mtime for p in deps):
    return obj, ''
  cmd = [_nvcc(), *NVCC_FLAGS, '-c', '-o', str(obj), str(CSRC / src)]
  if verbose:
    cmd += ['-Xptxas', '-v']
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError(f'nvcc failed on {src}:\n{proc.stdout}\n{proc.stderr}')
  return obj, proc.stderr


def build(force=False, verbose=False):
  if not force and not stale():
    return LIB
  OBJ.mkdir(exist_ok=True)
  with concurrent.futures.ThreadPoolExecutor(len(SOURCES)) as pool:
    done = list(pool.map(lambda s: _compile(s, force, verbose), SOURCES))
  if verbose:
    print('\n'.join(log for _, log in done if log))
  cmd = [_nvcc(), '-shared', '-o', str(LIB), *[str(o) for o, _ in done], '-ldl']
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError(f'link failed:\n{proc.stdout}\n{proc.stderr}')
  return LIB


if __name__ == '__main__':
  import sys
  print(build(force='--force' in sys.argv, verbose=True))
