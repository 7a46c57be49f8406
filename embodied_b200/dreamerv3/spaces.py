"""What the agent's networks see of the observation / action spaces.

Reference: `Encoder` / `Decoder` split the observation keys into images (3-D uint8, concatenated
along channels in sorted key order, dreamerv3/rssm.py:196-197,226-231) and vectors (<= 2-D: floats
squashed by symlog, integers one-hot, flattened and concatenated in sorted key order by
`nets.DictConcat`, embodied/jax/nets.py:467-500); the action dict goes through the same DictConcat
(rssm.py:78) and the policy head has one distribution per action key (dreamerv3/agent.py:61-64).
A key is described by the tuple (name, kind, shape, classes): kind 'disc' (integer / bool space,
`classes` values per element) or 'cont'.
"""
import numpy as np

EXCLUDE = ('is_first', 'is_last', 'is_terminal', 'reward')      # dreamerv3/agent.py:37-39


def keyspec(name, space):
  shape = tuple(int(x) for x in space.shape)
  if space.discrete:
    classes = np.asarray(space.classes).flatten()
    assert (classes == classes[0]).all(), (name, classes)       # nets.py:486-487
    return (name, 'disc', shape, int(classes[0]))
  return (name, 'cont', shape, 0)


def size(spec):
  return int(np.prod(spec[2], dtype=np.int64))


def width(spec):
  """Columns the key occupies after DictConcat."""
  return size(spec) * (spec[3] if spec[1] == 'disc' else 1)


def analyze(obs_space, act_space):
  """-> dict(imgkeys=[(name, channels)], image=(H, W, C_total) | None, vecspec=[...], actspec=[...])."""
  obs = {k: v for k, v in obs_space.items() if k not in EXCLUDE and not k.startswith('log/')}
  assert all(len(s.shape) <= 3 for s in obs.values()), obs            # rssm.py:193
  imgkeys = sorted(k for k, s in obs.items() if len(s.shape) == 3)
  for k in imgkeys:
    assert obs[k].dtype == np.uint8, (k, obs[k].dtype)                # rssm.py:228
  image = None
  if imgkeys:
    hw = {tuple(obs[k].shape[:2]) for k in imgkeys}
    assert len(hw) == 1, hw
    image = (*hw.pop(), sum(int(obs[k].shape[2]) for k in imgkeys))
  vec = [keyspec(k, obs[k]) for k in sorted(obs) if len(obs[k].shape) <= 2]
  act = [keyspec(k, act_space[k]) for k in sorted(act_space) if k != 'reset']
  return dict(imgkeys=[(k, int(obs[k].shape[2])) for k in imgkeys], image=image,
              vecspec=vec, actspec=act)
