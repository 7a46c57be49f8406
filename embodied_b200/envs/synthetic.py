"""Synthetic benchmark environments: obs of a
given image shape drawn from ``default_rng(1000 + env_id)``, reward ~ N(0, 1),
episode ends every `length` steps, one discrete action.  No simulator."""
import numpy as np

from .. import elements
from ..core import base


class SyntheticImage(base.Env):

  def __init__(self, index=0, size=(64, 64, 3), classes=5, length=500,
               pool=8):
    self.shape = tuple(size)
    self.classes = classes
    self.length = length
    rng = np.random.default_rng(1000 + index)
    # A small pool of pre-drawn frames: the env itself must not dominate the
    # Driver measurement (the reference's Dummy allocates one frame per step).
    self.frames = rng.integers(0, 256, (pool, *self.shape), dtype=np.uint8)
    self.rewards = rng.standard_normal(1024).astype(np.float32)
    self.t = 0
    self.done = False

  @property
  def obs_space(self):
    S = elements.Space
    return {
        'image': S(np.uint8, self.shape),
        'reward': S(np.float32),
        'is_first': S(bool), 'is_last': S(bool), 'is_terminal': S(bool),
    }

  @property
  def act_space(self):
    S = elements.Space
    return {'reset': S(bool), 'action': S(np.int32, (), 0, self.classes)}

  def step(self, action):
    if action['reset'] or self.done:
      self.t, self.done = 0, False
      first = True
    else:
      self.t += 1
      first = False
    self.done = self.t >= self.length
    return dict(
        image=self.frames[self.t % len(self.frames)],
        reward=self.rewards[self.t % 1024],
        is_first=first, is_last=self.done, is_terminal=self.done)


class SyntheticProprio(base.Env):
  """DMC-proprio-shaped synthetic env (BASELINE config 4; dm_control walker
  layout): orientations f32[14], height f32[], velocity f32[9],
  one continuous action f32[6] in [-1, 1].  The next observation depends on the
  action so that a dropped or mis-routed action row shows up in the data."""

  def __init__(self, index=0, length=200):
    self.length = length
    self.rng = np.random.default_rng(2000 + index)
    self.t = 0
    self.done = False
    self.last = np.zeros(6, np.float32)

  @property
  def obs_space(self):
    S = elements.Space
    return {
        'orientations': S(np.float32, (14,)), 'height': S(np.float32), 'velocity': S(np.float32, (9,)),
        'reward': S(np.float32),
        'is_first': S(bool), 'is_last': S(bool), 'is_terminal': S(bool),
    }

  @property
  def act_space(self):
    S = elements.Space
    return {'reset': S(bool), 'action': S(np.float32, (6,), -1, 1)}

  def step(self, action):
    if action['reset'] or self.done:
      self.t, self.done = 0, False
      self.last = np.zeros(6, np.float32)
      return self._obs(True, False)
    self.t += 1
    self.last = np.asarray(action['action'], np.float32)
    self.done = self.t >= self.length
    return self._obs(False, self.done)

  def _obs(self, first, last):
    ori = self.rng.standard_normal(14).astype(np.float32)
    ori[:6] += self.last
    return dict(
        orientations=ori, height=np.float32(self.last.sum()),
        velocity=self.rng.standard_normal(9).astype(np.float32),
        reward=np.float32(np.abs(self.last).mean()),
        is_first=first, is_last=last, is_terminal=False)
