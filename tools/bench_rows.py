"""Micro-benchmark of the row engine at BASELINE config-2 sizes (CUDA events on
the launching stream, L2 flushed between iterations).  Prints JSON lines."""
import json
import sys
import pathlib

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200 import _lib  # noqa: E402

PEAK = 6545.6
try:
  PEAK = json.load(open(pathlib.Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json'))['hbm_gbs']
except Exception:
  pass


def timeit(fn, iters=20, warmup=5, flush=None):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  times = []
  for _ in range(iters):
    if flush is not None:
      flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b) * 1e-3)
  return float(np.median(times)), float(np.min(times))


def main():
  lib = _lib.load()
  stream = torch.cuda.current_stream().cuda_stream
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  rows_table = 64 * 1024
  keys = {'image': 12288, 'dyn/deter': 32768, 'dyn/stoch': 8192, 'stepid': 20,
          'reward': 4, 'is_first': 1, 'is_last': 1, 'is_terminal': 1, 'action': 4}
  tables = {k: torch.randint(0, 255, (rows_table, rb), dtype=torch.uint8, device='cuda')
            for k, rb in keys.items()}
  rng = np.random.default_rng(0)
  for name, B, L in [('gather_B16_L65', 16, 65), ('gather_B64_L65', 64, 65),
                     ('gather_B128_L257', 128, 257)]:
    n = B * L
    starts = rng.integers(0, rows_table - L, B)
    src = torch.from_numpy((starts[:, None] + np.arange(L)[None]).reshape(-1).astype(np.int64)).cuda()
    outs = {k: torch.empty((n, rb), dtype=torch.uint8, device='cuda') for k, rb in keys.items()}
    kl = []
    for k, rb in keys.items():
      op, aux, aux_stride = _lib.OP_COPY, None, 0
      if k == 'is_first':
        op = _lib.OP_FIRST
      if k == 'is_last':
        op, aux, aux_stride = _lib.OP_LAST, tables['is_first'].data_ptr(), 1
      kl.append(_lib.Key(src=tables[k].data_ptr(), dst=outs[k].data_ptr(), aux=aux,
                         aux_stride=aux_stride, src_stride=rb, dst_stride=rb,
                         row_bytes=rb, op=op))
    cons = torch.empty(n, dtype=torch.int32, device='cuda')
    kl.append(_lib.Key(dst=cons.data_ptr(), dst_stride=4, row_bytes=4, op=_lib.OP_FILL32))
    arr = _lib.keys_array(kl)
    fn = lambda: _lib.check(lib.emb_replay_gather(arr, len(kl), src.data_ptr(), n, L, stream))
    med, best = timeit(fn, flush=flush)
    nbytes = 2 * n * (sum(keys.values()) + 4)
    print(json.dumps({'kernel': name, 'us_median': med * 1e6, 'us_min': best * 1e6,
                      'algorithmic_bytes': nbytes, 'GBs': nbytes / med / 1e9,
                      'frac_of_measured_peak': nbytes / med / 1e9 / PEAK}), flush=True)
  # append: N=256 rows of every key from a dense staging block
  for name, N in [('append_N256', 256), ('append_N1024', 1024)]:
    stag = {k: torch.randint(0, 255, (N, rb), dtype=torch.uint8, device='cuda') for k, rb in keys.items()}
    dst = torch.from_numpy((np.arange(N) * 61 % rows_table).astype(np.int64)).cuda()
    kl = [_lib.Key(src=stag[k].data_ptr(), dst=tables[k].data_ptr(), src_stride=rb,
                   dst_stride=rb, row_bytes=rb) for k, rb in keys.items()]
    arr = _lib.keys_array(kl)
    fn = lambda: _lib.check(lib.emb_replay_append_rows(arr, len(kl), dst.data_ptr(), N, stream))
    med, best = timeit(fn, flush=flush)
    nbytes = 2 * N * sum(keys.values())
    print(json.dumps({'kernel': name, 'us_median': med * 1e6, 'us_min': best * 1e6,
                      'algorithmic_bytes': nbytes, 'GBs': nbytes / med / 1e9,
                      'frac_of_measured_peak': nbytes / med / 1e9 / PEAK}), flush=True)
  # stage obs: 256 images u8 -> table + f32 normalised
  N = 256
  img = torch.randint(0, 255, (N, 12288), dtype=torch.uint8, device='cuda')
  norm = torch.empty((N, 12288), dtype=torch.float32, device='cuda')
  dst = torch.from_numpy((np.arange(N) * 61 % rows_table).astype(np.int64)).cuda()
  arr = _lib.keys_array([_lib.Key(src=img.data_ptr(), dst=tables['image'].data_ptr(),
                                  dst2=norm.data_ptr(), src_stride=12288, dst_stride=12288,
                                  dst2_stride=12288 * 4, row_bytes=12288,
                                  op=_lib.OP_NORM_U8_F32)])
  fn = lambda: _lib.check(lib.emb_driver_stage_obs(arr, 1, dst.data_ptr(), N, stream))
  med, best = timeit(fn, flush=flush)
  nbytes = N * 12288 * (1 + 1 + 4)
  print(json.dumps({'kernel': 'stage_obs_N256', 'us_median': med * 1e6, 'us_min': best * 1e6,
                    'algorithmic_bytes': nbytes, 'GBs': nbytes / med / 1e9,
                    'frac_of_measured_peak': nbytes / med / 1e9 / PEAK}), flush=True)
  # reference point: torch D2D copy of the same number of bytes as gather_B16_L65
  nbytes = 16 * 65 * 53283
  a = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
  b = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
  med, best = timeit(lambda: b.copy_(a), flush=flush)
  print(json.dumps({'kernel': 'torch_copy_same_bytes', 'us_median': med * 1e6,
                    'us_min': best * 1e6, 'GBs': 2 * nbytes / med / 1e9}), flush=True)


if __name__ == '__main__':
  main()
