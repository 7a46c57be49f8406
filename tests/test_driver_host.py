"""Driver step semantics (reference tests/test_driver.py:9-112 ported; numpy
mask double, no GPU) + the golden flag/action sequences of the real reference."""
from functools import partial as bind

import numpy as np

import embodied_b200 as embodied
from embodied_b200.envs import dummy
import doubles
import golden_cases


def make_env(length=10):
  return dummy.Dummy('disc', length=length)


def make_agent():
  env = make_env()
  agent = embodied.RandomAgent(env.obs_space, env.act_space)
  env.close()
  return agent


def make_driver(fns):
  return embodied.Driver(fns, parallel=False, ops=doubles.HostOps())


def test_episode_length():
  agent = make_agent()
  driver = make_driver([make_env])
  driver.reset(agent.init_policy)
  seq = []
  driver.on_step(lambda tran, _: seq.append(tran))
  driver(agent.policy, episodes=1)
  assert len(seq) == 11


def test_first_and_last_step():
  agent = make_agent()
  driver = make_driver([make_env])
  driver.reset(agent.init_policy)
  seq = []
  driver.on_step(lambda tran, _: seq.append(tran))
  driver(agent.policy, episodes=2)
  for index in [0, 11]:
    assert seq[index]['is_first'].item() is True
    assert seq[index]['is_last'].item() is False
  for index in [1, 10, 12]:
    assert seq[index]['is_first'].item() is False
  for index in [10, 21]:
    assert seq[index]['is_last'].item() is True
    assert seq[index]['is_first'].item() is False
  for index in [0, 1, 9, 11, 20]:
    assert seq[index]['is_last'].item() is False


def test_env_reset():
  agent = make_agent()
  driver = make_driver([bind(make_env, length=5)])
  driver.reset(agent.init_policy)
  seq = []
  driver.on_step(lambda tran, _: seq.append(tran))
  action = {'act_disc': np.ones(1, int), 'act_cont': np.zeros((1, 6), float)}
  driver(lambda carry, obs: (carry, action, {}), episodes=2)
  assert len(seq) == 12
  seq = {k: np.array([seq[i][k] for i in range(len(seq))]) for k in seq[0]}
  assert (seq['is_first'] == [1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0]).all()
  assert (seq['is_last'] == [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1]).all()
  assert (seq['act_disc'] == [1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 0]).all()
  # the reference's own test also reads seq['reset'], but its Driver no longer
  # puts 'reset' into the transition (driver.py:75-76); the next-step reset is
  # driver.acts['reset']:
  assert 'reset' not in seq
  assert driver.acts['reset'].tolist() == [True]


def test_agent_inputs():
  agent = make_agent()
  driver = make_driver([make_env])
  driver.reset(agent.init_policy)
  inputs, states = [], []

  def policy(carry, obs, mode='train'):
    inputs.append(obs)
    states.append(carry)
    _, act, _ = agent.policy(carry, obs, mode)
    return 'carry', act, {}
  seq = []
  driver.on_step(lambda tran, _: seq.append(tran))
  driver(policy, episodes=2)
  assert len(seq) == 22
  assert states == ([()] + ['carry'] * 21)
  for index in [0, 11]:
    assert inputs[index]['is_first'].item() is True
  for index in [1, 10, 12, 21]:
    assert inputs[index]['is_first'].item() is False
  for index in [10, 21]:
    assert inputs[index]['is_last'].item() is True
  for index in [0, 1, 9, 11, 20]:
    assert inputs[index]['is_last'].item() is False


def test_unexpected_reset():

  class UnexpectedReset(embodied.Wrapper):
    """Send is_first without preceeding is_last."""
    def __init__(self, env, when):
      super().__init__(env)
      self._when = when
      self._step = 0

    def step(self, action):
      if self._step == self._when:
        action = action.copy()
        action['reset'] = np.ones_like(action['reset'])
      self._step += 1
      return self.env.step(action)

  env = UnexpectedReset(make_env(length=4), when=3)
  agent = make_agent()
  driver = make_driver([lambda: env])
  driver.reset(agent.init_policy)
  steps = []
  driver.on_step(lambda tran, _: steps.append(tran))
  driver(agent.policy, episodes=1)
  assert len(steps) == 8
  steps = {k: np.array([x[k] for x in steps]) for k in steps[0]}
  assert (steps['is_first'] == [1, 0, 0, 1, 0, 0, 0, 0]).all()
  assert (steps['is_last'] == [0, 0, 0, 0, 0, 0, 0, 1]).all()


def golden_policy():
  counter = [0]

  def policy(carry, obs):
    n = len(obs['is_first'])
    counter[0] += 1
    act = {
        'act_disc': np.full(n, counter[0] % 5, np.int32) + np.arange(n, dtype=np.int32) % 2,
        'act_cont': (np.arange(n * 6, dtype=np.float32).reshape(n, 6)
                     - 7.5 + counter[0]).astype(np.float32)}
    return carry, act, {'aux': np.full((n, 2), -float(counter[0]), np.float32)}
  return policy


def check_driver_golden(driver):
  fix = np.load(golden_cases.GOLDEN / 'driver_seq.npz')
  rows = []
  driver.on_step(lambda tran, worker: rows.append((worker, tran)))
  driver.reset()
  driver(golden_policy(), steps=45)
  assert [w for w, _ in rows] == fix['workers'].tolist()
  for k in fix.files:
    if k.startswith('tran/'):
      got = np.stack([np.asarray(r[k[5:]]) for _, r in rows])
      assert got.dtype == fix[k].dtype, k
      assert got.tobytes() == fix[k].tobytes(), k


def test_matches_reference_golden_sequence():
  fns = [bind(dummy.Dummy, 'disc', size=(8, 8), length=3 + i) for i in range(3)]
  check_driver_golden(make_driver(fns))


def test_parallel_envs_match_serial():
  fns = [bind(dummy.Dummy, 'disc', size=(8, 8), length=3 + i) for i in range(3)]
  driver = embodied.Driver(fns, parallel=True, ops=doubles.HostOps())
  try:
    check_driver_golden(driver)
  finally:
    driver.close()


def test_replay_add_is_batched_and_in_worker_order():
  replay = embodied.Replay(
      length=3, capacity=100, store=doubles.HostStore(1024, staging_rows=8))
  fns = [bind(dummy.Dummy, 'disc', size=(8, 8), length=4) for _ in range(3)]
  driver = make_driver(fns)
  order = []
  driver.on_step(lambda tran, w: order.append(('a', w)))
  driver.on_step(replay.add)
  driver.on_step(lambda tran, w: order.append(('b', w)))
  batches = []
  driver.on_batch(lambda trans, n: batches.append(n))
  agent = make_agent()
  driver.reset(agent.init_policy)
  driver(agent.policy, steps=30)
  assert batches == [3] * 10
  assert order[:6] == [('a', 0), ('b', 0), ('a', 1), ('b', 1), ('a', 2), ('b', 2)]
  assert len(replay) == 3 * (10 - 3 + 1)
  assert [x for x in replay.store.launches if x[0] == 'append'] == [('append', 3)] * 10
  seq = replay.sample(4)
  assert (np.diff(seq['count'], axis=1) >= -4).all()
