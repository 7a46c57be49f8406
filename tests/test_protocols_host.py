"""The plugin protocols (embodied/core/base.py:1-73): the shipped envs, streams and
doubles are drop-ins, incomplete objects are rejected by name before any work starts."""
import numpy as np
import pytest

import embodied_b200 as embodied
from embodied_b200 import elements
from embodied_b200.core import base, streams
from embodied_b200.envs import dummy, synthetic


def test_base_classes_raise_with_the_signature():
  agent = base.Agent({}, {}, None)
  for name, (args, returns) in base.AGENT_PROTOCOL.items():
    with pytest.raises(NotImplementedError, match=name):
      getattr(agent, name)(*[None] * len(args))
  env = base.Env()
  with pytest.raises(NotImplementedError):
    env.obs_space
  with pytest.raises(NotImplementedError):
    env.step({})
  env.close()
  stream = base.Stream()
  assert iter(stream) is stream
  with pytest.raises(NotImplementedError):
    next(stream)


def test_shipped_objects_are_drop_ins():
  for env in (dummy.Dummy('disc'), synthetic.SyntheticImage(0), synthetic.SyntheticProprio(0)):
    assert base.implements_env(env) == [], type(env).__name__
    assert 'obs_space=' in repr(env)
  agent = embodied.RandomAgent(dummy.Dummy('disc').obs_space, dummy.Dummy('disc').act_space)
  assert base.implements_agent(agent) == []
  assert base.implements_stream(streams.Stateless(lambda: {}, )) == []


def test_incomplete_env_is_rejected_by_the_driver():
  class NoFlags(base.Env):
    obs_space = {'image': elements.Space(np.uint8, (4, 4, 3))}
    act_space = {'action': elements.Space(np.int32, (), 0, 3)}

    def step(self, action):
      return {'image': np.zeros((4, 4, 3), np.uint8)}

  lacking = base.implements_env(NoFlags())
  assert "obs_space['is_first']" in lacking and "act_space['reset']" in lacking
  with pytest.raises(TypeError, match='is_last'):
    embodied.Driver([NoFlags], parallel=False)

  class NoStep:
    obs_space = act_space = {}
  assert 'step' in base.implements_env(NoStep())


def test_incomplete_agent_is_named():
  class PolicyOnly:
    def policy(self, carry, obs, mode='train'):
      return carry, {}, {}
  lacking = base.implements_agent(PolicyOnly())
  assert 'policy' not in lacking and {'train', 'report', 'init_policy', 'save', 'load'} <= set(lacking)


def test_local_clock_semantics(monkeypatch):
  """embodied/core/clock.py:95-118: 0 = off, negative = always, positive = at most once
  per `every` seconds, the first call arms the timer and fires iff `first`."""
  from embodied_b200.core import clock
  now = [100.0]
  monkeypatch.setattr(clock.time, 'time', lambda: now[0])
  assert [clock.LocalClock(0)() for _ in range(3)] == [False] * 3
  assert [clock.LocalClock(-1)() for _ in range(3)] == [True] * 3
  c = clock.LocalClock(10)
  fired = []
  for t in (100, 105, 110, 111, 119.9, 120, 200):
    now[0] = t
    fired.append(c())
  assert fired == [False, False, True, False, False, True, True]
  c = clock.LocalClock(10, first=True)
  now[0] = 300
  assert c() is True and c() is False
  assert c(skip=True) is False
  now[0] = 311
  assert c(skip=True) is False and c() is True      # a skipped call does not touch the timer


def test_wait_returns_the_time_spent(monkeypatch, capsys):
  from embodied_b200.core import limiters
  assert limiters.wait(lambda: True, 'never printed') == 0
  now = [0.0]
  monkeypatch.setattr(limiters.time, 'time', lambda: now[0])
  monkeypatch.setattr(limiters.time, 'sleep', lambda s: now.__setitem__(0, now[0] + 25))
  calls = iter([False] * 5 + [True] * 3)
  spent = limiters.wait(lambda: next(calls), 'Replay sample is waiting', info='0 of 16', notify=60)
  assert spent == 100
  out = capsys.readouterr().out
  assert out.count('Replay sample is waiting') == 1 and '0 of 16' in out
