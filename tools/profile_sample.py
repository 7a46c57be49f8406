"""Host-side cost of Replay.sample (config-5 point B x T, default dreamerv3 rows): cProfile of the
call path + wall time per batch with the stream kept busy (no synchronise between batches).
Usage: python tools/profile_sample.py [B] [T]"""
import cProfile
import pathlib
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import embodied_b200 as embodied  # noqa: E402

Bx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Tx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
Lx, workers = Tx + 1, 64
spec = bench.ROW_KINDS['default']
g = torch.Generator(device='cuda').manual_seed(0)
replay = embodied.Replay(Lx, None, chunksize=1024, seed=0, staging_rows=workers, workers=workers)
step = {}
for k, (d, sh) in spec.items():
  if d is bool:
    step[k] = torch.zeros((workers, *sh), dtype=torch.bool, device='cuda')
  elif d is np.uint8:
    step[k] = torch.randint(0, 256, (workers, *sh), generator=g, device='cuda', dtype=torch.uint8)
  elif d is np.int32:
    step[k] = torch.randint(0, 5, (workers, *sh), generator=g, device='cuda', dtype=torch.int32)
  else:
    step[k] = torch.randn((workers, *sh), generator=g, device='cuda')
for _ in range(max(8 * Bx * Tx // workers, 2 * Lx)):
  replay.add_batch(step)
stream = iter(embodied.streams.Consec(embodied.streams.Stateless(replay.sample, Bx, 'train'),
                                      length=Tx, consec=1, prefix=1, strict=True, contiguous=True))
for _ in range(20):
  next(stream)
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
  batch = next(stream)
host = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
full = (time.perf_counter() - t0) / n
print(f'B={Bx} T={Tx}: host enqueue {host * 1e6:.1f} us / batch, to completion {full * 1e6:.1f} us / batch')
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
  batch = next(stream)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
