"""Where one benchmark step (bench.py Loop.step_resident) spends its time:
policy / append / sample / train / update, each synchronised and timed on the
host, averaged over a few steps.  Usage: python tools/profile_loop.py [steps]"""
import pathlib
import sys
import time

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
loop = bench.Loop(torch, 0, int(2e5))
while len(loop.replay) < 4 * bench.B * bench.L:
  loop.driver(loop.agent.policy, steps=bench.NENVS)
loop.make_resident()
acc = {}


def tick(name, fn):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  out = fn()
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  t2 = time.perf_counter()
  a = acc.setdefault(name, [0.0, 0.0, 0])
  a[0] += t1 - t0
  a[1] += t2 - t0
  a[2] += 1
  return out


def step():
  loop.pcarry, acts, outs = tick('policy', lambda: loop.agent.policy(loop.pcarry, loop.res))
  tick('add_batch', lambda: loop.replay.add_batch({**loop.res, **acts, **outs}))
  for _ in range(bench.TRAINS_PER_STEP):
    batch = tick('sample', lambda: next(loop.stream))
    loop.carry, o, m = tick('train', lambda: loop.agent.train(loop.carry, batch))
    tick('update', lambda: loop.replay.update(o['replay']))


for _ in range(2):
  step()
acc.clear()
t0 = time.perf_counter()
for _ in range(steps):
  step()
torch.cuda.synchronize()
total = (time.perf_counter() - t0) / steps
print(f'step (synchronised parts): {total * 1e3:.1f} ms')
for k, (host, full, n) in acc.items():
  print(f'  {k:10s} n/step={n // steps:2d}  host enqueue {host / n * 1e3:8.2f} ms   '
        f'to completion {full / n * 1e3:8.2f} ms   per step {full / steps * 1e3:8.2f} ms')
t0 = time.perf_counter()
for _ in range(steps):
  loop.learner_on = True
  loop.step_resident()
torch.cuda.synchronize()
print(f'step_resident (async): {(time.perf_counter() - t0) / steps * 1e3:.1f} ms')
t0 = time.perf_counter()
for _ in range(steps):
  loop.step_e2e()
torch.cuda.synchronize()
print(f'step_e2e: {(time.perf_counter() - t0) / steps * 1e3:.1f} ms')
