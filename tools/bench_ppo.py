"""ppo at the reference's default size (ppo/configs.yaml:94-109) on the dummy env's spaces:
time of one update on a (16, 64) replay batch and of one policy step over 16 envs.
Usage: python tools/bench_ppo.py [bfloat16|float32]"""
import pathlib
import sys
import time

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent / 'tests'))
from embodied_b200 import _lib, ppo                   # noqa: E402
from embodied_b200.ppo import config as configlib     # noqa: E402
import ppo_cases as cases                             # noqa: E402

obs, act = cases.dummy_spaces()
cfg = configlib.make(compute_dtype=sys.argv[1] if len(sys.argv) > 1 else 'bfloat16')
print('compute_dtype:', cfg.compute_dtype)
agent = ppo.Agent(obs, act, cfg)
print('parameters:', agent.store.count)
B, T, N = 16, 64, 16
data = cases.to_device(cases.batch(cfg, obs, act, B, T, seed=0))
carry = agent.init_train(B)
for _ in range(3):
  carry, _, mets = agent.train(carry, data)
torch.cuda.synchronize()
before, t0 = _lib.launch_count(), time.perf_counter()
n = 10
for _ in range(n):
  carry, _, mets = agent.train(carry, data)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f'update (B={B}, T={T}): {dt * 1e3:.1f} ms  = {B * T / dt:.0f} samples/s, loss {float(mets["loss"]):.4f}, '
      f'{(_lib.launch_count() - before) // n} own launches per update')
g = torch.Generator().manual_seed(0)
o = cases.to_device(cases.obs_batch(obs, (N,), g))
pc = agent.init_policy(N)
for _ in range(3):
  pc, a, e = agent.policy(pc, o)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
  pc, a, e = agent.policy(pc, o)
torch.cuda.synchronize()
print(f'policy step ({N} envs): {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms')
