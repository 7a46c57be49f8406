"""TEST INFRASTRUCTURE -- writes tests/golden/ppo_tiny.npz: the ppo oracle's own numbers over
three policy steps and three updates on the dummy env's spaces (debug-size networks), so that a
change of the oracle shows up as a diff and the GPU suite has a fixture that travels.
Run: python -m oracle.gen_ppo_golden"""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))


def run():
  from oracle import ppo_oracle as po
  import ppo_cases as cases
  torch.manual_seed(0)
  torch.set_num_threads(1)
  obs, act = cases.dummy_spaces()
  cfg = po.tiny_config(warmup=2)
  model, _ = cases.oracle_for(cfg, obs, act)
  out = {}
  B, T = 3, 8
  carry = (model.initial(B), {k: torch.zeros(B, *v.shape, dtype=torch.int32 if v.discrete else torch.float32)
                              for k, v in act.items()})
  g = torch.Generator().manual_seed(5)
  for step in range(3):
    o = cases.obs_batch(obs, (B,), g)
    o['is_first'][:] = step == 0
    carry, acts, ext = model.policy(carry, o, po.make_noise(act, (B,), 10 + step))
    for k, v in {**acts, **ext}.items():
      out[f'policy{step}/{k}'] = v.numpy()
  carry = (model.initial(B), {k: torch.zeros(B, *v.shape) for k, v in act.items()})
  for step in range(3):
    data = cases.batch(cfg, obs, act, B, T, seed=20 + step)
    carry, _, mets, grads, _ = model.train(carry, data)
    for k, v in mets.items():
      out[f'train{step}/{k}'] = np.asarray(float(v), np.float64)
    out[f'train{step}/gradsum'] = np.asarray(sum(float(g.double().abs().sum()) for g in grads.values()))
  out['params/abssum'] = np.asarray(sum(float(v.double().abs().sum()) for v in model.p.values()))
  return out


if __name__ == '__main__':
  path = ROOT / 'tests' / 'golden' / 'ppo_tiny.npz'
  np.savez(path, **run())
  print('wrote', path)
