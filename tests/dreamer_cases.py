"""Shared seeded inputs for the dreamerv3 parity tests (oracle vs product)."""
import numpy as np
import torch

from oracle import dreamer_oracle as do


SPACE_CASES = {
    # name: (observation keys, action keys) as {key: (dtype, shape, classes or None)}
    'mixed': ({'image': ('uint8', (64, 64, 3), None), 'vector': ('float32', (7,), None),
               'token': ('int32', (), 10), 'grid': ('int32', (2, 3), 4)},
              {'act_disc': ('int32', (), 5), 'act_cont': ('float32', (6,), None)}),
    # DMC-proprio shape (BASELINE config 4): vectors only, one continuous action
    'proprio': ({'orientations': ('float32', (14,), None), 'height': ('float32', (), None),
                 'velocity': ('float32', (9,), None)},
                {'action': ('float32', (6,), None)}),
    'twoimages': ({'cam0': ('uint8', (64, 64, 1), None), 'cam1': ('uint8', (64, 64, 2), None)},
                  {'action': ('int32', (2,), 3)}),
}


def spaces_of(name):
  """elements.Space dicts of a SPACE_CASES entry (+ reward / flags / reset)."""
  import numpy as np
  from embodied_b200 import elements
  S = elements.Space
  def space(dtype, shape, classes):
    if classes:
      return S(getattr(np, dtype), shape, 0, classes)
    if dtype == 'uint8':
      return S(np.uint8, shape)
    return S(getattr(np, dtype), shape, -1.0, 1.0) if dtype == 'float32' else S(getattr(np, dtype), shape)
  obs, act = SPACE_CASES[name]
  obs_space = {k: space(*v) for k, v in obs.items()}
  obs_space.update(reward=S(np.float32), is_first=S(bool), is_last=S(bool), is_terminal=S(bool))
  act_space = {k: space(*v) for k, v in act.items()}
  act_space['reset'] = S(bool)
  return obs_space, act_space


def oracle_config_for(name, **over):
  """tiny oracle config whose space description matches SPACE_CASES[name]."""
  from embodied_b200.dreamerv3 import spaces as spacelib
  obs_space, act_space = spaces_of(name)
  info = spacelib.analyze(obs_space, act_space)
  ocfg = do.tiny_config(**over)
  ocfg.update(image=info['image'], imgkeys=info['imgkeys'], vecspec=info['vecspec'],
              actspec=info['actspec'], actions=sum(spacelib.width(a) for a in info['actspec']))
  return ocfg, obs_space, act_space


def batch(cfg, B, T, seed=1, flags=True):
  L = T + cfg.replay_context
  g = torch.Generator().manual_seed(seed)
  if cfg.get('actspec'):
    return _batch_general(cfg, B, L, T, g, flags)
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


def _batch_general(cfg, B, L, T, g, flags):
  data = {}
  for key, ch in cfg.get('imgkeys') or []:
    data[key] = torch.randint(0, 256, (B, L, *cfg.image[:2], ch), generator=g, dtype=torch.uint8)
  for key, kind, shape, classes in list(cfg.vecspec) + list(cfg.actspec):
    if kind == 'disc':
      data[key] = torch.randint(0, classes, (B, L, *shape), generator=g, dtype=torch.int32)
    else:
      data[key] = torch.randn(B, L, *shape, generator=g).clamp(-1, 1) * (3.0 if key in dict(
          (v[0], 0) for v in cfg.vecspec) else 1.0)
  data.update({
      'reward': torch.randn(B, L, generator=g),
      'is_first': torch.zeros(B, L, dtype=torch.bool),
      'is_last': torch.zeros(B, L, dtype=torch.bool),
      'is_terminal': torch.zeros(B, L, dtype=torch.bool),
      'dyn/deter': torch.randn(B, L, cfg.deter, generator=g) * 0.1,
      'dyn/stoch': torch.nn.functional.one_hot(
          torch.randint(0, cfg.classes, (B, L, cfg.stoch), generator=g), cfg.classes).float(),
      'stepid': torch.randint(0, 256, (B, L, 20), generator=g, dtype=torch.uint8),
      'consec': torch.zeros(B, L, dtype=torch.int32)})
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
  return {k: (to_device(v, device) if isinstance(v, dict) else v.to(device)) for k, v in tree.items()}


def clone(tree):
  return {k: (clone(v) if isinstance(v, dict) else v.clone()) for k, v in tree.items()}
