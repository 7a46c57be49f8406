"""An agent that acts uniformly at random and learns nothing -- the plumbing
stand-in of the reference (embodied/core/random.py:4-39), used to exercise
Driver / Replay / run.train without a model.  Stateless: every carry is the
empty tuple, train / report return empty outputs, save returns None.
"""
import numpy as np

from . import base

_NO_CARRY = ()


def _uniform_batch(space, n):
  """n independent draws from `space`, stacked on a new leading axis."""
  return np.stack([space.sample() for _ in range(n)])


class RandomAgent(base.Agent):

  def __init__(self, obs_space, act_space, config=None):
    self.obs_space = obs_space
    self.act_space = act_space
    # `reset` is set by the Driver, never by the policy (driver.py:72-76)
    self._sampled = {k: s for k, s in act_space.items() if k != 'reset'}

  def _carry(self, batch_size):
    return _NO_CARRY

  init_policy = init_train = init_report = _carry

  def policy(self, carry, obs, mode='train'):
    n = len(obs['is_first'])
    return carry, {k: _uniform_batch(s, n) for k, s in self._sampled.items()}, {}

  def train(self, carry, data):
    return carry, {}, {}

  def report(self, carry, data):
    return carry, {}

  def stream(self, st):
    return st

  def save(self):
    return None

  def load(self, data=None):
    return None
