"""Times emb_conv5x5_nhwc_tc against the library convolution (cuDNN via torch) at the
dreamerv3 size200m layer shapes, B*T = 1024 images, bf16; CUDA events, L2 flushed."""
import json
import pathlib
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200.dreamerv3 import ops  # noqa: E402

WG = True
LAYERS = [('enc/cnn1', 32, 128, 192), ('enc/cnn2', 16, 192, 256), ('enc/cnn3', 8, 256, 256),
          ('dec/conv2', 8, 256, 256), ('dec/conv1', 16, 256, 192), ('dec/conv0', 32, 192, 128)]


def timeit(fn, flush, reps=10):
  for _ in range(3):
    fn()
  ts = []
  for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return float(np.median(ts))


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  torch.backends.cudnn.benchmark = True
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  for name, hw, cin, cout in LAYERS:
    x = torch.randn((n, hw, hw, cin), device='cuda').to(torch.bfloat16)
    w = (torch.randn((5, 5, cin, cout), device='cuda') / 50).to(torch.bfloat16)
    wp = ops.pack_conv_weight(w)
    xc = x.permute(0, 3, 1, 2)
    wc = w.permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last)
    t_tc = timeit(lambda: ops.conv_tc(x, wp), flush)
    t_lib = timeit(lambda: F.conv2d(xc, wc, padding=2), flush)
    flops = 2.0 * n * hw * hw * 25 * cin * cout
    gy = torch.randn((n, hw, hw, cout), device='cuda').to(torch.bfloat16)
    gyc = gy.permute(0, 3, 1, 2)
    t_wg = timeit(lambda: ops.conv_wgrad(x, gy, 5), flush)
    t_wg_lib = timeit(lambda: torch.nn.grad.conv2d_weight(xc, (cout, cin, 5, 5), gyc, padding=2), flush)
    print(json.dumps({'layer': name, 'n': n, 'hw': hw, 'cin': cin, 'cout': cout,
                      'tc_ms': t_tc, 'tc_tflops': flops / t_tc / 1e9,
                      'cudnn_ms': t_lib, 'cudnn_tflops': flops / t_lib / 1e9,
                      'wgrad_tc_ms': t_wg, 'wgrad_tc_tflops': flops / t_wg / 1e9,
                      'wgrad_cudnn_ms': t_wg_lib, 'wgrad_cudnn_tflops': flops / t_wg_lib / 1e9}))


if __name__ == '__main__':
  main()
