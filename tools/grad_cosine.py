"""Per-tensor cosine between the gradients of one dreamerv3 update in fp32 and in bf16
compute (same parameters, batch and injected noise) -- a screen for tensors the bf16
path treats differently.  Usage: python tools/grad_cosine.py [size] [B] [T]"""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
from embodied_b200 import dreamerv3, elements  # noqa: E402


def main():
  size = sys.argv[1] if len(sys.argv) > 1 else 'size12m'
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
  T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
  S = elements.Space
  obs = {'image': S(np.uint8, (64, 64, 3)), 'reward': S(np.float32), 'is_first': S(bool),
         'is_last': S(bool), 'is_terminal': S(bool)}
  act = {'reset': S(bool), 'action': S(np.int32, (), 0, 5)}
  g = torch.Generator(device='cuda').manual_seed(0)
  L = T + 1
  data = {
      'image': torch.randint(0, 256, (B, L, 64, 64, 3), generator=g, device='cuda', dtype=torch.uint8),
      'reward': torch.randn(B, L, generator=g, device='cuda'),
      'is_first': torch.zeros(B, L, dtype=torch.bool, device='cuda'),
      'is_last': torch.zeros(B, L, dtype=torch.bool, device='cuda'),
      'is_terminal': torch.zeros(B, L, dtype=torch.bool, device='cuda'),
      'action': torch.randint(0, 5, (B, L), generator=g, device='cuda', dtype=torch.int32),
      'stepid': torch.zeros(B, L, 20, dtype=torch.uint8, device='cuda'),
      'consec': torch.zeros(B, L, dtype=torch.int32, device='cuda'),
  }
  grads, noise = [], None
  for dtype in ('float32', 'bfloat16'):
    agent = dreamerv3.Agent(obs, act, dreamerv3.config.make(size, compute_dtype=dtype, graph='off'))
    if noise is None:
      first = agent
      noise = agent.make_noise(B, T)
    else:
      agent.store.master.copy_(first.store.master)
      agent.store.refresh_low()
      agent.store.version += 1
    carry = agent.init_train(B)
    cfg = agent.cfg
    full = dict(data)
    full['dyn/deter'] = torch.zeros(B, L, cfg.deter, device='cuda')
    full['dyn/stoch'] = torch.zeros(B, L, cfg.stoch, cfg.classes, device='cuda')
    agent.train(carry, full, noise)
    grads.append({k: agent.store.view('grad', k).double().clone() for k in agent.store.specs})
  rows = []
  for k in grads[0]:
    a, b = grads[0][k], grads[1][k]
    na, nb = float(a.norm()), float(b.norm())
    cos = float((a * b).sum() / (na * nb)) if na > 0 and nb > 0 else float('nan')
    rows.append((cos if cos == cos else -2.0, k, na, nb))
  rows.sort()
  for cos, k, na, nb in rows[:25]:
    print(f'{cos:+.4f}  |g32|={na:.3e}  |g16|={nb:.3e}  {k}')
  print('tensors:', len(rows), ' cos<0.9:', sum(r[0] < 0.9 for r in rows), ' cos<0.99:', sum(r[0] < 0.99 for r in rows))


if __name__ == '__main__':
  main()
