"""Shared seeded inputs for the ppo parity tests (oracle vs product)."""
import numpy as np
import torch

from oracle import ppo_oracle as po


def dummy_spaces():
  """The spaces of embodied/envs/dummy.py:6-59 ('disc' / 'cont' share one layout)."""
  from embodied_b200.envs import dummy
  env = dummy.Dummy('disc')
  act = {k: v for k, v in env.act_space.items() if k != 'reset'}
  return env.obs_space, act


def vector_spaces():
  """The dummy env without its image: the encoder is DictEmbed + MLP only."""
  obs, act = dummy_spaces()
  return {k: v for k, v in obs.items() if k != 'image'}, act


def small_spaces():
  """A 16x16 two-camera + vector case: the 3x3/stride-2 SAME pool on odd and even sizes."""
  from embodied_b200 import elements
  S = elements.Space
  obs = {'cam': S(np.uint8, (18, 18, 1)), 'vector': S(np.float32, (5,)), 'token': S(np.int32, (), 0, 7),
         'reward': S(np.float32), 'is_first': S(bool), 'is_last': S(bool), 'is_terminal': S(bool)}
  act = {'action': S(np.float32, (3,), -1.0, 1.0)}
  return obs, act


def obs_batch(obs_space, lead, g):
  data = {}
  for k, s in obs_space.items():
    shape = (*lead, *s.shape)
    if k in ('is_first', 'is_last', 'is_terminal'):
      data[k] = torch.zeros(shape, dtype=torch.bool)
    elif s.dtype == np.uint8:
      data[k] = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    elif s.discrete:
      hi = int(np.asarray(s.classes).max())
      data[k] = torch.randint(0, hi, shape, generator=g, dtype=torch.int32)
    else:
      data[k] = torch.randn(shape, generator=g)
  return data


def act_batch(act_space, lead, g):
  data = {}
  for k, s in act_space.items():
    shape = (*lead, *s.shape)
    if s.discrete:
      data[k] = torch.randint(0, int(np.asarray(s.classes).max()), shape, generator=g, dtype=torch.int32)
    else:
      data[k] = torch.rand(shape, generator=g) * 2 - 1
  return data


def batch(cfg, obs_space, act_space, B, T, seed=1, flags=True):
  L = T + cfg.replay_context
  g = torch.Generator().manual_seed(seed)
  data = obs_batch(obs_space, (B, L), g)
  data.update(act_batch(act_space, (B, L), g))
  for k in act_space:
    data[f'logp/{k}'] = -torch.rand(B, L, generator=g) - 0.5
  if cfg.recurrent:
    data['memory'] = torch.randn(B, L, cfg.rnn_units, generator=g) * 0.3
  data['stepid'] = torch.randint(0, 256, (B, L, 20), generator=g, dtype=torch.uint8)
  data['consec'] = torch.zeros(B, L, dtype=torch.int32)
  if flags and L >= 6:
    data['is_first'][0, 3] = True
    data['is_last'][0, 2] = True
    data['is_terminal'][0, 2] = True
    data['is_last'][-1, 4] = True            # time-limit end: last but not terminal
    data['is_first'][-1, 5] = True
  return data


def to_device(tree, device='cuda'):
  if isinstance(tree, dict):
    return {k: to_device(v, device) for k, v in tree.items()}
  if isinstance(tree, (tuple, list)):
    return type(tree)(to_device(v, device) for v in tree)
  return tree.to(device) if torch.is_tensor(tree) else tree


def product_config(ocfg, **over):
  from embodied_b200.ppo import config as configlib
  keys = configlib.make().keys()
  return configlib.make(**{k: v for k, v in ocfg.items() if k in keys}).update(over)


def oracle_for(ocfg, obs_space, act_space, seed=0, outscale=1.0):
  vals = po.init_params(ocfg, obs_space, act_space, seed, outscale_override=outscale)
  return po.PPO(ocfg, obs_space, act_space, {k: v.clone() for k, v in vals.items()}), vals
