"""Environment wrappers that sit between a simulator and the Driver -- the ones the
shipped configs stack in ``wrap_env`` (dreamerv3/main.py:249-258: NormalizeAction,
UnifyDtypes, CheckSpaces, ClipAction) plus TimeLimit.  Behaviour is pinned against
the reference's own classes (embodied/core/wrappers.py:8-55, 76-110, 204-270) by
``tests/test_wrappers_host.py``; the implementation is this project's: every
wrapper is a ``Wrapper`` with two optional hooks, ``map_action`` (outer -> inner
action dict) and ``map_obs`` (inner -> outer observation dict), and ``step`` is
written once.  These run inside the environment process, host Python by
definition; nothing here touches the device.
"""
import numpy as np

from .. import elements

_PLAIN_VALUES = (np.ndarray, np.generic, list, tuple, int, float, bool)


class Wrapper:
  """Transparent proxy around an env.  Attribute lookups fall through to the
  wrapped env; a name the whole stack lacks raises ValueError (as the reference
  does, so that ``hasattr`` on a wrapped env keeps meaning "implemented")."""

  def __init__(self, env):
    self.env = env

  def __getattr__(self, name):
    if name.startswith('__') or 'env' not in self.__dict__:
      raise AttributeError(name)
    try:
      return getattr(self.env, name)
    except AttributeError:
      raise ValueError(name) from None

  def __len__(self):
    return len(self.env)

  def __bool__(self):
    return bool(self.env)

  # hooks -----------------------------------------------------------------
  def map_action(self, action):
    return action

  def map_obs(self, obs):
    return obs

  def step(self, action):
    return self.map_obs(self.env.step(self.map_action(action)))


class TimeLimit(Wrapper):
  """Ends an episode after `duration` steps (0 = never).  When the episode is over --
  by the limit, by the env, or because the caller asks for a reset -- the next step
  starts a new one: with ``reset=True`` the env is really reset, otherwise it keeps
  running and the step is merely labelled ``is_first``.  The caller's action dict
  is updated in place, as in the reference (the Driver reuses it)."""

  def __init__(self, env, duration, reset=True):
    super().__init__(env)
    self._limit = duration
    self._hard = reset
    self._elapsed = 0
    self._over = False

  def step(self, action):
    if action['reset'] or self._over:
      self._elapsed, self._over = 0, False
      action['reset'] = bool(self._hard)
      obs = self.env.step(action)
      if not self._hard:
        obs['is_first'] = True
      return obs
    self._elapsed += 1
    obs = self.env.step(action)
    if self._limit and self._elapsed >= self._limit:
      obs['is_last'] = True
    self._over = obs['is_last']
    return obs


class ClipAction(Wrapper):
  """Clips one action key into [low, high] before the env sees it."""

  def __init__(self, env, key='action', low=-1, high=1):
    super().__init__(env)
    self._key, self._bounds = key, (low, high)

  def map_action(self, action):
    return {**action, self._key: np.clip(action[self._key], *self._bounds)}


class NormalizeAction(Wrapper):
  """Presents the bounded dimensions of a continuous action as [-1, 1] and maps
  them back affinely; unbounded dimensions pass through unchanged."""

  def __init__(self, env, key='action'):
    super().__init__(env)
    inner = env.act_space[key]
    bounded = np.isfinite(inner.low) & np.isfinite(inner.high)
    self._key, self._bounded = key, bounded
    self._lo = np.where(bounded, inner.low, -1)
    self._hi = np.where(bounded, inner.high, 1)
    unit = np.ones_like(self._lo)
    self._outer = elements.Space(
        np.float32, inner.shape,
        np.where(bounded, -unit, self._lo), np.where(bounded, unit, self._hi))

  @property
  def act_space(self):
    return {**self.env.act_space, self._key: self._outer}

  def map_action(self, action):
    value = action[self._key]
    # same operation order as the reference: the result is bit-identical
    restored = (value + 1) / 2 * (self._hi - self._lo) + self._lo
    return {**action, self._key: np.where(self._bounded, restored, value)}


def _unified(dtype):
  """float* -> float32, uint8 stays, other integers -> int32, anything else stays."""
  if np.issubdtype(dtype, np.floating):
    return np.float32
  if np.issubdtype(dtype, np.uint8):
    return np.uint8
  if np.issubdtype(dtype, np.integer):
    return np.int32
  return dtype


class UnifyDtypes(Wrapper):
  """One dtype per kind on the outside (see `_unified`); actions are cast back to
  what the env declared before it sees them."""

  def __init__(self, env):
    super().__init__(env)
    self._inner_act = {k: s.dtype for k, s in env.act_space.items()}
    self._outer_obs = {k: _unified(s.dtype) for k, s in env.obs_space.items()}
    recast = lambda spaces: {
        k: elements.Space(_unified(s.dtype), s.shape, s.low, s.high) for k, s in spaces.items()}
    self._spaces = recast(env.obs_space), recast(env.act_space)

  @property
  def obs_space(self):
    return self._spaces[0]

  @property
  def act_space(self):
    return self._spaces[1]

  def map_action(self, action):
    action = dict(action)
    for key, dtype in self._inner_act.items():
      action[key] = np.asarray(action[key], dtype)
    return action

  def map_obs(self, obs):
    for key, dtype in self._outer_obs.items():
      obs[key] = np.asarray(obs[key], dtype)
    return obs


class CheckSpaces(Wrapper):
  """Every value that crosses the env boundary must lie in its declared space;
  observation and action keys must not collide (they share one transition dict)."""

  def __init__(self, env):
    shared = set(env.obs_space) & set(env.act_space)
    assert not shared, f'keys used as both observation and action: {sorted(shared)}'
    super().__init__(env)

  def map_action(self, action):
    spaces = self.env.act_space
    for key, value in action.items():
      self._require(key, value, spaces[key])
    return action

  def map_obs(self, obs):
    spaces = self.env.obs_space
    for key, value in obs.items():
      self._require(key, value, spaces[key])
    return obs

  @staticmethod
  def _require(key, value, space):
    if not isinstance(value, _PLAIN_VALUES):
      raise TypeError(f'Invalid type {type(value)} for key {key}.')
    if value in space:
      return
    arr = np.asarray(value)
    raise ValueError(
        f"Value for '{key}' with dtype {arr.dtype}, shape {arr.shape}, "
        f'lowest {arr.min()}, highest {arr.max()} is not in {space}.')
