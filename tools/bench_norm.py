"""Times the channel-row norm kernels (emb_rmsnorm_act_fwd / _bwd) at the shapes of
the size200m conv stacks with CUDA events (L2 flushed between launches) and prints
achieved HBM GB/s: fwd moves 2 x bytes(x), bwd 3 x bytes(x)."""
import json
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200.dreamerv3 import ops  # noqa: E402

ROOT = pathlib.Path(__file__).resolve().parent.parent
try:
  PEAK = json.load(open(ROOT / 'MEASURED_PEAKS.json'))['hbm_gbs']
except Exception:
  PEAK = 6650.0


def main():
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  for shape in [(1024, 32, 32, 128), (1024, 16, 16, 192), (1024, 8, 8, 256), (16384, 1024)]:
    x = torch.randn(*shape, device='cuda').to(torch.bfloat16).requires_grad_(True)
    scale = torch.ones(shape[-1], device='cuda', requires_grad=True)
    bias = torch.zeros(shape[-1], device='cuda', requires_grad=True) if len(shape) == 4 else None
    gy = torch.randn(*shape, device='cuda').to(torch.bfloat16)
    tf, tb = [], []
    for _ in range(6):
      flush.zero_()
      e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
      e[0].record()
      y = ops.rmsnorm_act(x, scale, True, 1e-4, bias)
      e[1].record()
      flush.zero_()
      e[2].record()
      y.backward(gy)
      e[3].record()
      torch.cuda.synchronize()
      tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
    nbytes = x.numel() * 2
    f, b = sorted(tf)[len(tf) // 2] * 1e-3, sorted(tb)[len(tb) // 2] * 1e-3
    print(json.dumps({'shape': shape, 'fwd_us': f * 1e6, 'fwd_GBs': 2 * nbytes / f / 1e9,
                      'bwd_us': b * 1e6, 'bwd_GBs': 3 * nbytes / b / 1e9, 'peak_GBs': PEAK}))


if __name__ == '__main__':
  main()
