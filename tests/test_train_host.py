"""run.train loop counters and wiring (reference tests/test_train.py:12-33,
tests/utils.py:8-104 TestAgent ported): the agent asserts obs['count']
continuity across policy calls and inside every sampled batch row -- a semantic
oracle for the Driver + Replay wiring.  Numpy doubles, no GPU."""
import time

import numpy as np

import embodied_b200 as embodied
from embodied_b200 import elements
from embodied_b200.envs import dummy
import doubles


class CountAgent:

  device_obs = False

  def __init__(self, obs_space, act_space):
    self.obs_space, self.act_space = obs_space, act_space
    self.stats_ = {'env_steps': 0, 'replay_steps': 0, 'reports': 0,
                   'saves': 0, 'loads': 0, 'created': time.time()}

  def stats(self):
    out = dict(self.stats_)
    out['lifetime'] = time.time() - out.pop('created')
    return out

  def init_policy(self, batch_size):
    return (np.zeros(batch_size),)

  def init_train(self, batch_size):
    return (np.zeros(batch_size),)

  def init_report(self, batch_size):
    return ()

  def policy(self, carry, obs, mode='train'):
    assert set(obs.keys()) == set(self.obs_space.keys())
    B = len(obs['is_first'])
    self.stats_['env_steps'] += B
    carry, = carry
    assert np.asarray(carry).shape == (B,)
    target = (carry + 1) * (1 - obs['is_first'])
    assert (obs['count'] == target).all()
    act = {k: np.stack([v.sample() for _ in range(B)])
           for k, v in self.act_space.items() if k != 'reset'}
    return (target,), act, {}

  def train(self, carry, data):
    data = {k: np.asarray(v) for k, v in data.items()}
    expected = sorted(
        set(self.obs_space) | set(self.act_space) | {'stepid', 'consec'})
    expected.remove('reset')
    assert sorted(data.keys()) == expected, (sorted(data.keys()), expected)
    B, T = data['count'].shape
    carry, = carry
    assert carry.shape == (B,)
    self.stats_['replay_steps'] += B * T
    for t in range(T):
      current = data['count'][:, t]
      reset = data['is_first'][:, t]
      target = (1 - reset) * (carry + 1) + reset * current
      assert (current == target).all()
      carry = current
    return (carry,), {}, {'loss': np.float32(0)}

  def report(self, carry, data):
    self.stats_['reports'] += 1
    return carry, {'scalar': np.float32(0), 'vector': np.zeros(10)}

  def stream(self, st):
    return st

  def save(self):
    self.stats_['saves'] += 1
    return self.stats_

  def load(self, data):
    self.stats_ = data
    self.stats_['loads'] += 1


def test_run_loop(tmpdir):
  args = elements.Config(
      steps=1000, train_ratio=32.0, log_every=0.1, report_every=0.2,
      save_every=0.2, report_batches=1, from_checkpoint='',
      usage=dict(psutil=True), debug=True, logdir=str(tmpdir), envs=4,
      batch_size=8, batch_length=16, replay_context=0, report_length=8)
  make_env = lambda index: dummy.Dummy('disc', size=(64, 64), length=100)
  env = make_env(0)
  agent = CountAgent(env.obs_space, env.act_space)

  def make_replay():
    return embodied.Replay(
        length=args.batch_length, capacity=1e4,
        store=doubles.HostStore(1024, staging_rows=16))

  def make_stream(replay, mode):
    fn = embodied.streams.Stateless(replay.sample, args.batch_size, mode)
    return embodied.streams.Consec(
        fn, length=args.batch_length, consec=1, prefix=0, strict=True,
        contiguous=True)

  def make_logger():
    return elements.Logger(elements.Counter(), [elements.logger.TerminalOutput()])

  slow_policy = agent.policy

  def policy(carry, obs, mode='train'):
    time.sleep(0.004)   # lifetime >= 1 s so the clocks fire (test_train.py:21)
    return slow_policy(carry, obs, mode)
  agent.policy = policy

  args = args.update(driver_ops=doubles.HostOps())
  if True:
    embodied.run.train(
        lambda: agent, make_replay, make_env, make_stream, make_logger, args)
    stats = agent.stats()
    replay_steps = args.steps * args.train_ratio
    assert stats['lifetime'] >= 1
    assert np.allclose(stats['env_steps'], args.steps, 100, 0.1)
    assert np.allclose(stats['replay_steps'], replay_steps, 100, 0.1)
    assert stats['reports'] >= 1
    assert stats['saves'] >= 2
    assert stats['loads'] == 0
    args2 = args.update(steps=2 * args.steps)
    embodied.run.train(
        lambda: agent, make_replay, make_env, make_stream, make_logger, args2)
    stats = agent.stats()
    assert stats['loads'] == 1
    assert np.allclose(stats['env_steps'], args2.steps, 100, 0.1)


def test_episode_stats_aggregate_log_keys():
  """`log/<k>` scalars become `<k>/avg|max|sum` over the episode (embodied/run/train.py:44-46),
  reset on is_first, emitted when the episode ends."""
  import importlib
  trainlib = importlib.import_module("embodied_b200.run.train")

  class Sink:
    def __init__(self):
      self.rows = []
    def add(self, mapping, prefix=None):
      self.rows.append((prefix, dict(mapping)))
  logger, epstats = Sink(), Sink()
  stats = trainlib._EpisodeStats(logger, epstats)
  vals = np.array([[1.0, 5.0], [3.0, -2.0], [2.0, 4.0]])       # (t, env)
  for t in range(3):
    stats({'reward': np.array([1.0, 0.5 * t]), 'is_first': np.array([t == 0, t == 0]),
           'is_last': np.array([t == 2, t == 1]), 'log/x': vals[t]}, 2)
  # env 1 ended at t = 1 (two steps), env 0 at t = 2 (three steps)
  (_, first), (_, second) = epstats.rows
  assert first['log/x/sum'] == 3.0 and first['log/x/max'] == 5.0 and first['log/x/avg'] == 1.5
  assert second['log/x/sum'] == 6.0 and second['log/x/max'] == 3.0 and second['log/x/avg'] == 2.0
  assert [r[1]['length'] for r in logger.rows] == [2, 3]
