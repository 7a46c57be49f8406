"""embodied_b200.core.wrappers against the reference's own classes
(embodied/core/wrappers.py:8-55, 76-110, 204-270, loaded verbatim by path): the same
seeded action stream through the same wrapper stack over the same toy env must give the
same spaces, the same observations and the same errors.  Plus behaviour checks that
need no reference (they also run on machines without /root/reference)."""
import numpy as np
import pytest

from embodied_b200 import elements
from embodied_b200.core import wrappers
from oracle import refload

needs_ref = pytest.mark.skipif(
    not refload.available(), reason='/root/reference not on this machine')


class Toy:
  """Deterministic env with a bounded + an unbounded action dimension, float64 / int64 /
  uint8 / bool observations, an episode end every `length` steps; records what it saw."""

  def __init__(self, Space, length=5):
    self.Space = Space
    self.length = length
    self.t = 0
    self.seen = []

  @property
  def obs_space(self):
    S = self.Space
    return {'image': S(np.uint8, (2, 2, 3)), 'vec': S(np.float64, (3,)), 'count': S(np.int64, (), 0, 1000),
            'reward': S(np.float32), 'is_first': S(bool), 'is_last': S(bool), 'is_terminal': S(bool)}

  @property
  def act_space(self):
    S = self.Space
    low = np.array([-2.0, -np.inf, 0.0])
    high = np.array([4.0, np.inf, 1.0])
    return {'reset': S(bool), 'action': S(np.float64, (3,), low, high), 'choice': S(np.int64, (), 0, 4)}

  def step(self, action):
    self.seen.append({k: np.array(v) for k, v in action.items()})
    first = bool(action['reset']) or self.t == 0
    self.t = 1 if first else self.t + 1
    a = np.asarray(action['action'], np.float64)
    return {
        'image': np.full((2, 2, 3), self.t % 256, np.uint8), 'vec': a * 0.5 + self.t,
        'count': np.int64(self.t + int(action['choice'])), 'reward': np.float32(a.sum()),
        'is_first': first, 'is_last': self.t >= self.length, 'is_terminal': False}


def stack(mod, Space, **kw):
  env = Toy(Space, **kw)
  top = mod.NormalizeAction(env)
  top = mod.UnifyDtypes(top)
  top = mod.CheckSpaces(top)
  top = mod.ClipAction(top, low=-1, high=1)
  top = mod.TimeLimit(top, 7, reset=kw.get('length', 5) != 9)
  return env, top


def actions(seed, n):
  rng = np.random.default_rng(seed)
  for i in range(n):
    yield {'reset': bool(i == 0 or rng.random() < 0.1),
           'action': rng.uniform(-1.5, 1.5, 3).astype(np.float32),
           'choice': np.int32(rng.integers(0, 4))}


def same_space(a, b):
  return (np.dtype(a.dtype) == np.dtype(b.dtype) and tuple(a.shape) == tuple(b.shape) and
          np.array_equal(a.low, b.low) and np.array_equal(a.high, b.high))


@needs_ref
@pytest.mark.parametrize('length', [5, 9, 20])
def test_wrapper_stack_equals_the_reference(length):
  ref = refload.load()
  renv, rtop = stack(ref.wrappers, ref.elements.Space, length=length)
  menv, mtop = stack(wrappers, elements.Space, length=length)
  for name in ('obs_space', 'act_space'):
    rs, ms = getattr(rtop, name), getattr(mtop, name)
    assert list(rs) == list(ms)
    for k in rs:
      assert same_space(rs[k], ms[k]), (name, k, rs[k], ms[k])
  for ra, ma in zip(actions(3, 200), actions(3, 200)):
    ro, mo = rtop.step(ra), mtop.step(ma)
    assert ra['reset'] == ma['reset']              # TimeLimit edits the caller's dict in place
    assert list(ro) == list(mo)
    for k in ro:
      x, y = np.asarray(ro[k]), np.asarray(mo[k])
      assert x.dtype == y.dtype and np.array_equal(x, y), (k, x, y)
  assert len(renv.seen) == len(menv.seen)
  for rs, ms in zip(renv.seen, menv.seen):         # what reached the env, bit for bit
    for k in rs:
      assert rs[k].dtype == ms[k].dtype and np.array_equal(rs[k], ms[k]), k


@needs_ref
def test_errors_equal_the_reference():
  ref = refload.load()
  for mod, Space in ((ref.wrappers, ref.elements.Space), (wrappers, elements.Space)):
    env = mod.CheckSpaces(Toy(Space))
    with pytest.raises(ValueError, match="Value for 'choice'"):
      env.step({'reset': True, 'action': np.zeros(3), 'choice': np.int64(9)})
    with pytest.raises(TypeError, match='Invalid type'):
      env.step({'reset': True, 'action': np.zeros(3), 'choice': 'two'})
    with pytest.raises(ValueError, match='no_such_attribute'):
      env.no_such_attribute
    with pytest.raises(AttributeError):
      env.__no_such_dunder__
    assert env.length == 5 and bool(env)           # attributes fall through the stack


def test_time_limit_soft_and_hard_resets():
  env = Toy(elements.Space, length=100)
  hard = wrappers.TimeLimit(env, 3, reset=True)
  act = lambda: {'reset': False, 'action': np.zeros(3), 'choice': 0}
  first = hard.step({**act(), 'reset': True})
  assert first['is_first'] and not first['is_last']
  flags = [hard.step(act())['is_last'] for _ in range(3)]
  assert flags == [False, False, True]
  a = act()
  again = hard.step(a)                              # episode over: a real reset is forced
  assert a['reset'] is True and again['is_first'] and env.seen[-1]['reset']
  soft = wrappers.TimeLimit(Toy(elements.Space, length=100), 2, reset=False)
  soft.step({**act(), 'reset': True})
  soft.step(act()); soft.step(act())
  a = act()
  cont = soft.step(a)                               # labelled first, the env keeps running
  assert a['reset'] is False and cont['is_first'] and cont['count'] == 4


def test_normalize_and_unify_shapes():
  env = wrappers.UnifyDtypes(wrappers.NormalizeAction(Toy(elements.Space)))
  sp = env.act_space['action']
  assert sp.dtype == np.float32
  # unbounded dimensions are also DECLARED as [-1, 1] (reference wrappers.py:94-102) but pass through
  assert np.array_equal(sp.low, [-1, -1, -1]) and np.array_equal(sp.high, [1, 1, 1])
  assert env.obs_space['vec'].dtype == np.float32 and env.obs_space['count'].dtype == np.int32
  assert env.obs_space['image'].dtype == np.uint8 and env.obs_space['is_last'].dtype == bool
  obs = env.step({'reset': True, 'action': np.array([1, 0.7, -1], np.float32), 'choice': np.int32(2)})
  inner = env.env.env.seen[-1]
  assert inner['action'].dtype == np.float64 and inner['choice'].dtype == np.int64
  assert np.allclose(inner['action'], [4.0, 0.7, 0.0])      # bounded dims mapped back, free dim untouched
  assert obs['vec'].dtype == np.float32 and obs['count'].dtype == np.int32


@needs_ref
@pytest.mark.parametrize('size,length', [((64, 64), 4), ((8, 8), 1), ((5, 7), 3)])
def test_dummy_env_equals_the_reference(size, length):
  """embodied_b200.envs.dummy.Dummy against embodied/envs/dummy.py:6-59: spaces (order,
  dtype, shape, bounds) and every value of a reset-laden episode stream."""
  from embodied_b200.envs import dummy
  ref = refload.load()
  renv, menv = ref.dummy.Dummy('disc', size=size, length=length), dummy.Dummy('disc', size=size, length=length)
  for name in ('obs_space', 'act_space'):
    rs, ms = getattr(renv, name), getattr(menv, name)
    assert list(rs) == list(ms), name
    for k in rs:
      assert same_space(rs[k], ms[k]), (name, k, rs[k], ms[k])
  rng = np.random.default_rng(5)
  for i in range(40):
    reset = bool(i == 0 or rng.random() < 0.15)
    act = {'reset': reset, 'act_disc': np.int32(1), 'act_cont': np.zeros(6, np.float32)}
    ro, mo = renv.step(dict(act)), menv.step(dict(act))
    assert list(ro) == list(mo)
    for k in ro:
      x, y = np.asarray(ro[k]), np.asarray(mo[k])
      assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), (i, k, x, y)
