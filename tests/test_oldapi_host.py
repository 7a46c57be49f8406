"""SURVEY 8f rank 3 / BASELINE config 4, the adapter half: an agent written against the OLD
embodied API generation (director/jaxagent.py:104-230 -- `policy(obs, state)`, `train(data,
state)`, `report(data)`, `dataset(generator_fn)`, five-factory run.train) runs through
embodied_b200's Driver + Replay + run.train on the DMC-proprio-shaped dummy.  Numpy doubles."""
import numpy as np
import pytest

import embodied_b200 as embodied
from embodied_b200 import director, elements
from embodied_b200.envs import synthetic
import doubles


class OldGenerationAgent:
  """Old API.  Checks the wiring semantically: the action it chose at step t must come back as
  the `action` of the replayed transition, and `orientations[:, :6]` of the next step carries it
  (SyntheticProprio adds the last action into the observation)."""

  def __init__(self, obs_space, act_space):
    self.obs_space, self.act_space = obs_space, act_space
    self.calls = dict(policy=0, train=0, report=0, dataset=0, batches=0, saves=0, loads=0)

  policy_keys = '/(wm|task_behavior)/'

  def init_policy(self, batch_size):
    return {'step': np.zeros(batch_size, np.int64)}

  def init_train(self, batch_size):
    return {'updates': 0}

  def policy(self, obs, state, mode='train'):
    self.calls['policy'] += 1
    assert mode in ('train', 'eval', 'explore')
    assert set(obs) == set(self.obs_space), sorted(obs)
    ori = np.asarray(obs['orientations'])
    act = np.tanh(ori[:, 6:12]).astype(np.float32)
    step = np.where(np.asarray(obs['is_first']), 0, state['step'] + 1)
    return {'action': act, 'log_entropy': np.zeros(len(act), np.float32)}, {'step': step}

  def train(self, data, state):
    self.calls['train'] += 1
    assert 'stepid' not in data and 'consec' not in data, sorted(data)
    assert 'rng' in data                                             # added by dataset(): jaxagent.py:222-225
    act, ori = np.asarray(data['action']), np.asarray(data['orientations'])
    first, last = np.asarray(data['is_first']), np.asarray(data['is_last'])
    expect = np.tanh(ori[:, :, 6:12]).astype(np.float32)
    live = ~last                                                     # the Driver zeroes actions on is_last
    np.testing.assert_allclose(act[live], expect[live], rtol=1e-6)
    assert (act[last] == 0).all()
    return {}, {'updates': state['updates'] + 1}, {'model_loss': np.float32(1.5)}

  def report(self, data):
    self.calls['report'] += 1
    return {'openl': np.float32(0.0)}

  def dataset(self, generator_fn):
    self.calls['dataset'] += 1
    def gen():
      for batch in generator_fn():
        self.calls['batches'] += 1
        yield {**batch, 'rng': np.zeros(2, np.uint32)}
    return gen()

  def save(self):
    self.calls['saves'] += 1
    return dict(self.calls)

  def load(self, state):
    self.calls['loads'] += 1


def test_adapter_swaps_the_call_conventions():
  env = synthetic.SyntheticProprio(0)
  old = OldGenerationAgent(env.obs_space, env.act_space)
  agent = director.OldApiAgent(old)
  assert embodied.core.base.implements_agent(agent) == []
  assert 'reset' not in agent.act_space and agent.policy_keys == old.policy_keys
  obs = {k: np.stack([np.asarray(env.step({'reset': True, 'action': np.zeros(6, np.float32)})[k])] * 3)
         for k in env.obs_space}
  carry, act, out = agent.policy(agent.init_policy(3), obs, mode='eval')
  assert set(act) == {'action'} and out == {}                        # log_* entries are not replayed
  assert (carry['step'] == 0).all()
  carry, rep = agent.report((), {'stepid': 0, 'consec': 0, 'x': 1})
  assert carry == () and set(rep) == {'openl'}
  with pytest.raises(KeyError, match='action'):
    old.policy = lambda obs, state, mode='train': ({}, state)
    agent.policy(None, obs)


def test_config4_old_api_agent_through_five_factory_train(tmpdir):
  n = 32                                                             # a slice of config 4's 512 envs
  args = elements.Config(
      steps=1200, train_ratio=16.0, log_every=0.05, report_every=0.05, save_every=0.05,
      report_batches=1, from_checkpoint='', usage=dict(psutil=False), debug=True,
      logdir=str(tmpdir), envs=n, batch_size=8, batch_length=16, replay_context=0,
      report_length=16, driver_ops=doubles.HostOps())
  make_env = lambda index: synthetic.SyntheticProprio(index, length=11 + index % 7)
  env = make_env(0)
  old = OldGenerationAgent(env.obs_space, env.act_space)
  make_replay = lambda: embodied.Replay(
      length=args.batch_length, capacity=1e4, chunksize=256, store=doubles.HostStore(256, staging_rows=n))
  make_logger = lambda: elements.Logger(elements.Counter(), [elements.logger.TerminalOutput()])
  director.train(lambda: old, make_replay, make_env, make_logger, args)
  calls = old.calls
  assert calls['policy'] * n >= args.steps
  assert calls['dataset'] == 2                                       # train and report streams
  want = args.steps * args.train_ratio / (args.batch_size * args.batch_length)
  assert 0.5 * want <= calls['train'] <= 1.1 * want, (calls, want)
  assert calls['batches'] >= calls['train'] and calls['saves'] >= 1
