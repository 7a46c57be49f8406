"""Shared seeded inputs for the dreamerv3 parity tests (oracle vs product)."""
import numpy as np
import torch

from oracle import dreamer_oracle as do


def batch(cfg, B, T, seed=1, flags=True):
  L = T + cfg.replay_context
  g = torch.Generator().manual_seed(seed)
  data = {
      'image': torch.randint(0, 256, (B, L, *cfg.image), generator=g, dtype=torch.uint8),
      'reward': torch.randn(B, L, generator=g),
      'is_first': torch.zeros(B, L, dtype=torch.bool),
      'is_last': torch.zeros(B, L, dtype=torch.bool),
      'is_terminal': torch.zeros(B, L, dtype=torch.bool),
      'action': torch.randint(0, cfg.actions, (B, L), generator=g, dtype=torch.int32),
      'dyn/deter': torch.randn(B, L, cfg.deter, generator=g) * 0.1,
      'dyn/stoch': torch.nn.functional.one_hot(
          torch.randint(0, cfg.classes, (B, L, cfg.stoch), generator=g), cfg.classes).float(),
      'stepid': torch.randint(0, 256, (B, L, 20), generator=g, dtype=torch.uint8),
      'consec': torch.zeros(B, L, dtype=torch.int32),
  }
  if flags and T >= 5:
    data['is_first'][0, 3] = True
    data['is_last'][0, 2] = True
    data['is_terminal'][0, 2] = True
    data['is_terminal'][-1, 4] = True
    data['is_last'][-1, 4] = True
  return data


def product_config(ocfg, dtype='float32'):
  from embodied_b200.dreamerv3 import config as C
  cfg = C.make('size1m')
  for k in ocfg:
    if k in cfg:
      cfg[k] = ocfg[k]
  cfg['compute_dtype'] = dtype
  return cfg


def to_device(tree, device='cuda'):
  return {k: v.to(device) for k, v in tree.items()}
