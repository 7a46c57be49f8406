"""Forward-pass time of one dreamerv3 update by model section (CUDA events around
the Model methods, eager launches).  Usage: python tools/profile_sections.py [size]"""
import collections
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200 import dreamerv3, elements  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else 'size200m'
S = elements.Space
obs = {'image': S(np.uint8, (64, 64, 3)), 'reward': S(np.float32), 'is_first': S(bool),
       'is_last': S(bool), 'is_terminal': S(bool)}
act = {'reset': S(bool), 'action': S(np.int32, (), 0, 5)}
agent = dreamerv3.Agent(obs, act, dreamerv3.config.make(size, compute_dtype='bfloat16', graph='off'))
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

m = agent.model
events = []


def wrap(name):
  fn = getattr(m, name)

  def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn(*a, **k)
    e1.record()
    events.append((name, e0, e1))
    return out
  setattr(m, name, timed)


for name in ('encoder', 'observe', 'prior', 'kl_losses', 'decoder', 'imagine', 'imag_loss', 'head',
             'twohot_loss', 'twohot_pred', 'slow_value_logits', 'lambda_return'):
  wrap(name)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
carry, outs, mets = agent.train(carry, data)
e1.record()
torch.cuda.synchronize()
tot = collections.OrderedDict()
for name, a, b in events:
  tot[name] = tot.get(name, 0.0) + a.elapsed_time(b)
print(f'update (eager): {e0.elapsed_time(e1):.1f} ms')
for k, v in tot.items():
  print(f'  {k:20s} {v:7.2f} ms  (forward, inclusive of nested sections)')
