"""Self-contained test environment, same observation/action layout as the
reference's (embodied/envs/dummy.py:6-59): image u8[H,W,3]=255, vector f32[7]=0,
token i32=0, count f32, float2d f32[4,5]=1, int2d i32[2,3]=1, reward 0/1,
episodes of `length` steps."""
import numpy as np

from .. import elements
from ..core import base


class Dummy(base.Env):

  def __init__(self, task, size=(64, 64), length=100):
    del task
    self.size = tuple(size)
    self.length = length
    self.count = 0
    self.done = False

  @property
  def obs_space(self):
    S = elements.Space
    return {
        'image': S(np.uint8, self.size + (3,)),
        'vector': S(np.float32, (7,)),
        'token': S(np.int32, (), 0, 256),
        'count': S(np.float32, (), 0, self.length),
        'float2d': S(np.float32, (4, 5)),
        'int2d': S(np.int32, (2, 3), 0, 4),
        'reward': S(np.float32),
        'is_first': S(bool),
        'is_last': S(bool),
        'is_terminal': S(bool),
    }

  @property
  def act_space(self):
    S = elements.Space
    return {
        'reset': S(bool),
        'act_disc': S(np.int32, (), 0, 5),
        'act_cont': S(np.float32, (6,)),
    }

  def step(self, action):
    if action.pop('reset') or self.done:
      self.count, self.done = 0, False
      return self._obs(0, is_first=True)
    self.count += 1
    self.done = self.count >= self.length
    return self._obs(1, is_last=self.done, is_terminal=self.done)

  def _obs(self, reward, is_first=False, is_last=False, is_terminal=False):
    return dict(
        image=np.full(self.size + (3,), 255, np.uint8),
        vector=np.zeros(7, np.float32),
        token=np.zeros((), np.int32),
        count=np.float32(self.count),
        float2d=np.ones((4, 5), np.float32),
        int2d=np.ones((2, 3), np.int32),
        reward=np.float32(reward),
        is_first=is_first, is_last=is_last, is_terminal=is_terminal)
