"""Weight gradient of the two thin convolutions' matrix products (millions of pixel rows x a small
matrix): ops.ThinMatmul (tcgen05 weight-gradient kernel, ksize = 1) against the library's split-K
GEMM, timed alone with CUDA events, L2 flushed.  Usage: python tools/bench_thin.py"""
import json
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200.dreamerv3 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, n=10):
  for _ in range(3):
    fn()
  ms = []
  for _ in range(n):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
  return sorted(ms)[len(ms) // 2]


for name, P, K, N in (('enc/cnn0 (patches @ w)', 1024 * 64 * 64, 80, 128), ('dec/imgout (x @ taps)', 1024 * 32 * 32, 128, 80)):
  g = torch.Generator(device='cuda').manual_seed(0)
  a = torch.randn((P, K), generator=g, device='cuda').to(torch.bfloat16)
  gy = torch.randn((P, N), generator=g, device='cuda').to(torch.bfloat16)
  m_is_in = K in (128, 256)
  m, n = (K, N) if m_is_in else (N, K)
  dw = torch.zeros((1, m, (n + 63) // 64 * 64), dtype=torch.float32, device='cuda')
  own = timed(lambda: ops._conv_general(x=a.data_ptr(), gy=gy.data_ptr(), dw=dw.data_ptr(), n=P // 64, h=1, w=64,
                                        cin=K, cout=N, ksize=1, m_is_in=int(m_is_in), gy_up=1, gy_phase=0))
  lib = timed(lambda: a.t() @ gy)
  nbytes = P * (K + N) * 2
  print(json.dumps({'layer': name, 'P': P, 'K': K, 'N': N, 'own_ms': own, 'library_ms': lib,
                    'own_GBs': nbytes / own / 1e6, 'library_GBs': nbytes / lib / 1e6}))
