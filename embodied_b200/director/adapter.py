"""Adapter: an old-generation embodied agent behind the current Agent protocol.

Old generation (director/jaxagent.py:43-250, director/agent.py:51-128):
  init_policy(batch) / init_train(batch)         -> state
  policy(obs, state, mode='train')               -> (outs, state)      outs = actions (+ extras)
  train(data, state)                             -> (outs, state, metrics)
  report(data)                                   -> metrics
  dataset(generator_fn)                          -> iterator of train batches
  save() / load(state)
Current generation (embodied/core/base.py:1-31):
  policy(carry, obs, mode) -> (carry, act, out);  train(carry, data) -> (carry, out, metrics);
  report(carry, data) -> (carry, metrics);  stream(st) -> st;  init_report(batch).
"""
import functools

from ..core import base


class OldApiAgent(base.Agent):

  def __init__(self, agent, act_space=None):
    self.agent = agent
    self.obs_space = getattr(agent, 'obs_space', None)
    space = act_space if act_space is not None else getattr(agent, 'act_space', {})
    self.act_space = {k: v for k, v in space.items() if k != 'reset'}
    # an old-API agent that keeps its observations on the device may say so itself
    self.device_obs = bool(getattr(agent, 'device_obs', False))

  @property
  def policy_keys(self):
    return getattr(self.agent, 'policy_keys', '.*')

  @property
  def ext_space(self):
    return dict(getattr(self.agent, 'ext_space', {}) or {})

  def init_policy(self, batch_size):
    return self.agent.init_policy(batch_size)

  def init_train(self, batch_size):
    return self.agent.init_train(batch_size)

  def init_report(self, batch_size):
    return ()

  def policy(self, carry, obs, mode='train', **kwargs):
    outs, carry = self.agent.policy(obs, carry, mode=mode, **kwargs)
    # the old Driver merged everything policy() returned into the transition
    # (director-era driver: `acts = {k: v for k in outs if not k.startswith('log_')}`); the current
    # one takes actions and replay extras apart (embodied/core/driver.py:68-76)
    act = {k: v for k, v in outs.items() if k in self.act_space}
    out = {k: v for k, v in outs.items() if k not in self.act_space and not k.startswith('log_')}
    missing = set(self.act_space) - set(act)
    if missing:
      raise KeyError(f'old-API policy returned no value for action keys {sorted(missing)}')
    return carry, act, out

  def train(self, carry, data):
    # the old generation knows nothing of the replay-context keys the current streams add
    data = {k: v for k, v in data.items() if k not in ('stepid', 'consec')}
    outs, carry, metrics = self.agent.train(data, carry)
    return carry, dict(outs or {}), dict(metrics or {})

  def report(self, carry, data):
    data = {k: v for k, v in data.items() if k not in ('stepid', 'consec')}
    return carry, dict(self.agent.report(data) or {})

  def stream(self, st):
    dataset = getattr(self.agent, 'dataset', None)
    if dataset is None:
      return st
    return dataset(lambda: iter(st))                         # jaxagent.py:221-226 takes a generator FUNCTION

  def save(self):
    return self.agent.save()

  def load(self, data, regex=None):
    if regex:
      raise NotImplementedError('old-API agents restore whole checkpoints (jaxagent.py:236-250)')
    return self.agent.load(data)


def train(make_agent, make_replay, make_env, make_logger, args):
  """The five-factory `embodied.run.train` of the director generation (director/train.py:61-65):
  the train stream is built here from `args.batch_size / batch_length` (the old loop called
  `replay.dataset(batch)`), the agent is wrapped, and the current loop runs."""
  from .. import run
  from ..core import streams

  def make_stream(replay, mode):
    fn = functools.partial(replay.sample, args.batch_size, mode)
    length = args.batch_length if mode == 'train' else args.get('report_length', args.batch_length)
    return streams.Consec(streams.Stateless(fn), length=length, consec=1, prefix=0,
                          strict=(mode == 'train'), contiguous=True)

  def wrapped_agent():
    agent = make_agent()
    return agent if not base.implements_agent(agent) and not _is_old(agent) else OldApiAgent(agent)

  return run.train(wrapped_agent, make_replay, make_env, make_stream, make_logger, args)


def _is_old(agent):
  """Old generation: no `init_report`, and `report` takes the batch only."""
  import inspect
  if not hasattr(agent, 'init_report'):
    return True
  try:
    params = list(inspect.signature(agent.report).parameters)
  except (TypeError, ValueError):
    return False
  return len(params) == 1
