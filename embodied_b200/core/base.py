"""The plugin protocols, unchanged from the reference
(embodied/core/base.py:1-73): an Agent is anything with
init_policy/init_train/init_report/policy/train/report/stream/save/load, an Env
anything with obs_space/act_space/step/close, a Stream an iterator with
save/load.  Batches handed to ``Agent.train`` are dicts of *device* tensors
(B, T, ...); observations handed to ``Agent.policy`` are dicts of numpy arrays
(N, ...) unless the agent sets ``device_obs = True``.
"""


class Agent:

  device_obs = False   # True: policy() receives the staged device tensors

  def __init__(self, obs_space, act_space, config):
    pass

  def init_train(self, batch_size):
    raise NotImplementedError('init_train(batch_size) -> carry')

  def init_report(self, batch_size):
    raise NotImplementedError('init_report(batch_size) -> carry')

  def init_policy(self, batch_size):
    raise NotImplementedError('init_policy(batch_size) -> carry')

  def train(self, carry, data):
    raise NotImplementedError('train(carry, data) -> carry, out, metrics')

  def report(self, carry, data):
    raise NotImplementedError('report(carry, data) -> carry, metrics')

  def policy(self, carry, obs, mode):
    raise NotImplementedError('policy(carry, obs, mode) -> carry, act, out')

  def stream(self, st):
    raise NotImplementedError('stream(st) -> st')

  def save(self):
    raise NotImplementedError('save() -> data')

  def load(self, data):
    raise NotImplementedError('load(data) -> None')


class Env:

  def __repr__(self):
    return (f'{type(self).__name__}(obs_space={self.obs_space}, '
            f'act_space={self.act_space})')

  @property
  def obs_space(self):
    # Must contain is_first, is_last, is_terminal (and usually reward, image);
    # keys starting with 'log/' bypass the agent and the replay.
    raise NotImplementedError('Returns: dict of spaces')

  @property
  def act_space(self):
    # Must contain the reset key as well as any actions.
    raise NotImplementedError('Returns: dict of spaces')

  def step(self, action):
    raise NotImplementedError('Returns: dict')

  def close(self):
    pass


class Stream:

  def __iter__(self):
    return self

  def __next__(self):
    raise NotImplementedError

  def save(self):
    raise NotImplementedError

  def load(self, state):
    raise NotImplementedError
