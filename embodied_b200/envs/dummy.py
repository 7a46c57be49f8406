"""The self-contained test environment of the reference (embodied/envs/dummy.py:6-59)
as a table: every observation key is (dtype, shape, low, high, constant value), and
both ``obs_space`` and the observations are generated from it.  Episodes last
``length`` steps; the first step of an episode pays reward 0, every later one 1;
``count`` is the step index inside the episode.  The `task` argument ('disc' /
'cont' in the reference's configs) does not change the layout.
"""
import numpy as np

from .. import elements
from ..core import base

# key -> (dtype, shape, low, high, constant fill); `size` / `length` are filled in per env
_CONSTANT_KEYS = {
    'image': (np.uint8, 'size+3', None, None, 255),
    'vector': (np.float32, (7,), None, None, 0),
    'token': (np.int32, (), 0, 256, 0),
    'float2d': (np.float32, (4, 5), None, None, 1),
    'int2d': (np.int32, (2, 3), 0, 4, 1),
}
_ACTIONS = {
    'act_disc': (np.int32, (), 0, 5),
    'act_cont': (np.float32, (6,), None, None),
}
_FLAGS = ('is_first', 'is_last', 'is_terminal')


def _space(dtype, shape, low=None, high=None):
  if low is None:
    return elements.Space(dtype, shape)
  return elements.Space(dtype, shape, low, high)


class Dummy(base.Env):

  def __init__(self, task, size=(64, 64), length=100):
    self.size = tuple(size)
    self.length = length
    self.count = 0
    self.done = False
    shape = lambda s: self.size + (3,) if s == 'size+3' else s
    self._layout = {k: (d, shape(s), lo, hi, fill) for k, (d, s, lo, hi, fill) in _CONSTANT_KEYS.items()}
    self._constants = {k: np.full(s, fill, d) for k, (d, s, _, _, fill) in self._layout.items()}

  @property
  def obs_space(self):
    spaces = {k: _space(d, s, lo, hi) for k, (d, s, lo, hi, _) in self._layout.items()}
    spaces['count'] = _space(np.float32, (), 0, self.length)
    spaces['reward'] = _space(np.float32, ())
    spaces.update({k: _space(bool, ()) for k in _FLAGS})
    # the reference's key order: image, vector, token, count, float2d, int2d, reward, flags
    order = ['image', 'vector', 'token', 'count', 'float2d', 'int2d', 'reward', *_FLAGS]
    return {k: spaces[k] for k in order}

  @property
  def act_space(self):
    return {'reset': _space(bool, ()), **{k: _space(*v) for k, v in _ACTIONS.items()}}

  def step(self, action):
    restart = action.pop('reset') or self.done      # the reference pops `reset` too
    if restart:
      self.count, self.done = 0, False
    else:
      self.count += 1
      self.done = self.count >= self.length
    obs = {k: v.copy() for k, v in self._constants.items()}
    obs['count'] = np.float32(self.count)
    obs['reward'] = np.float32(0 if restart else 1)
    obs['is_first'] = bool(restart)
    obs['is_last'] = obs['is_terminal'] = bool(self.done and not restart)
    order = ['image', 'vector', 'token', 'count', 'float2d', 'int2d', 'reward', *_FLAGS]
    return {k: obs[k] for k in order}
