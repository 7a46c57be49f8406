"""Kernel-level time breakdown of one dreamerv3 train step / policy step
(torch profiler, CUDA activities).  Usage: python tools/profile_train.py [size] [dtype]"""
import sys
import pathlib
import time

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200 import dreamerv3, elements  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else 'size200m'
dtype = sys.argv[2] if len(sys.argv) > 2 else 'bfloat16'
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 40
S = elements.Space
obs = {'image': S(np.uint8, (64, 64, 3)), 'reward': S(np.float32), 'is_first': S(bool),
       'is_last': S(bool), 'is_terminal': S(bool)}
act = {'reset': S(bool), 'action': S(np.int32, (), 0, 5)}
graph = sys.argv[4] if len(sys.argv) > 4 else 'auto'
agent = dreamerv3.Agent(obs, act, dreamerv3.config.make(size, compute_dtype=dtype, graph=graph))
cfg = agent.cfg
B, T, L = 16, 64, 65
g = torch.Generator(device='cuda').manual_seed(0)
data = {
    'image': torch.randint(0, 256, (B, L, 64, 64, 3), generator=g, device='cuda', dtype=torch.uint8),
    'reward': torch.randn(B, L, generator=g, device='cuda'),
    'is_first': torch.zeros(B, L, dtype=torch.bool, device='cuda'),
    'is_last': torch.zeros(B, L, dtype=torch.bool, device='cuda'),
    'is_terminal': torch.zeros(B, L, dtype=torch.bool, device='cuda'),
    'action': torch.randint(0, 5, (B, L), generator=g, device='cuda', dtype=torch.int32),
    'dyn/deter': torch.zeros(B, L, cfg.deter, device='cuda'),
    'dyn/stoch': torch.zeros(B, L, cfg.stoch, cfg.classes, device='cuda'),
    'stepid': torch.zeros(B, L, 20, dtype=torch.uint8, device='cuda'),
    'consec': torch.zeros(B, L, dtype=torch.int32, device='cuda')}
carry = agent.init_train(B)
for _ in range(3):
  carry, outs, mets = agent.train(carry, data)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
  carry, outs, mets = agent.train(carry, data)
torch.cuda.synchronize()
print(f'train step: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms   loss {float(mets["loss"]):.3f}')
print(f'max memory: {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')

# host enqueue time vs device time of one step: is the step launch-bound?
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
carry, outs, mets = agent.train(carry, data)
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f'host enqueue {t_host * 1e3:.1f} ms   device span {e0.elapsed_time(e1):.1f} ms')

# phases of the step (CUDA events): replay context + noise, forward, backward, optimiser
def phases():
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
  c, obs_, pa, sid = agent._apply_replay_context(carry, data)
  ev[0].record()
  noise = agent.make_noise(B, T)
  agent.store.begin_step(); agent.store.grad.zero_()
  ev[1].record()
  total, c2, o2, m2 = agent.model.loss(c, obs_, pa, noise, update=True)
  ev[2].record()
  total.backward()
  ev[3].record()
  agent.opt.step(); agent.opt.update_slow(); agent.store.begin_step()  # eager
  ev[4].record()
  torch.cuda.synchronize()
  return [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
phases()
ph = phases()
print('phases ms: noise %.2f  forward %.2f  backward %.2f  optimiser %.2f' % tuple(ph))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=graph == 'off') as prof:
  carry, outs, mets = agent.train(carry, data)
  torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=rows, max_name_column_width=90))
if graph == 'off':
  print(prof.key_averages(group_by_input_shape=True).table(
      sort_by='self_cuda_time_total', row_limit=3 * rows, max_name_column_width=60,
      max_shapes_column_width=90))

pobs = {'image': data['image'][:, 0].repeat(16, 1, 1, 1), 'is_first': torch.zeros(256, dtype=torch.bool, device='cuda')}
pc = agent.init_policy(256)
for _ in range(3):
  pc, a, o = agent.policy(pc, pobs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
  pc, a, o = agent.policy(pc, pobs)
torch.cuda.synchronize()
print(f'policy step (256 envs): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms')
