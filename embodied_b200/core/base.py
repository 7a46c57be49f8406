"""The three plugin protocols of the hot path, name for name what the reference
defines in embodied/core/base.py:1-73, written as tables so that the contract is
data: the same tables drive the stub methods of the base classes and the
``implements_*`` checks with which the Driver's in-process env pool and
``run.train`` reject an object that is not a drop-in before any device work starts.

Conventions on this side of the boundary: batches handed to ``Agent.train`` /
``Agent.report`` are dicts of *device* tensors (B, T, ...); observations handed
to ``Agent.policy`` are dicts of numpy arrays (N, ...) unless the agent sets
``device_obs = True`` (then they are the staged device tensors).
"""

# method -> (argument names after self, what it returns); reference line in base.py
AGENT_PROTOCOL = {
    'init_policy': (('batch_size',), 'carry'),                      # :12
    'init_train': (('batch_size',), 'carry'),                       # :6
    'init_report': (('batch_size',), 'carry'),                      # :9
    'policy': (('carry', 'obs', 'mode'), 'carry, act, out'),        # :21
    'train': (('carry', 'data'), 'carry, out, metrics'),            # :15
    'report': (('carry', 'data'), 'carry, metrics'),                # :18
    'stream': (('st',), 'st'),                                      # :24
    'save': ((), 'data'),                                           # :27
    'load': (('data',), 'None'),                                    # :30
}
ENV_PROTOCOL = {
    'step': (('action',), 'dict of numpy values'),                  # :54
    'close': ((), 'None'),                                          # :57
}
ENV_SPACES = ('obs_space', 'act_space')                             # :42-52
STREAM_PROTOCOL = {
    '__next__': ((), 'element'),                                    # :66
    'save': ((), 'state'),                                          # :69
    'load': (('state',), 'None'),                                   # :72
}
# every Env must expose these observation keys and the `reset` action (base.py:44-52 comments)
REQUIRED_OBS = ('is_first', 'is_last', 'is_terminal')
REQUIRED_ACT = ('reset',)


def _stub(name, args, returns):
  text = f'{name}({", ".join(args)}) -> {returns}'

  def method(self, *a, **kw):
    raise NotImplementedError(f'{type(self).__name__} must implement {text}')
  method.__name__ = name
  method.__doc__ = text
  return method


def _with_stubs(table, skip=()):
  def decorate(cls):
    for name, (args, returns) in table.items():
      if name not in skip and name not in cls.__dict__:
        setattr(cls, name, _stub(name, args, returns))
    return cls
  return decorate


def _missing(obj, table):
  return [n for n in table if not callable(getattr(obj, n, None))]


def implements_agent(obj):
  """Names of the Agent protocol `obj` lacks (empty list = drop-in)."""
  return _missing(obj, AGENT_PROTOCOL)


def implements_env(obj):
  """Names of the Env protocol `obj` lacks, including required space keys."""
  lacking = _missing(obj, ENV_PROTOCOL)
  for prop, required in zip(ENV_SPACES, (REQUIRED_OBS, REQUIRED_ACT)):
    try:
      space = getattr(obj, prop)
    except NotImplementedError:
      lacking.append(prop)
      continue
    lacking += [f'{prop}[{k!r}]' for k in required if k not in space]
  return lacking


def implements_stream(obj):
  return _missing(obj, STREAM_PROTOCOL) + ([] if hasattr(obj, '__iter__') else ['__iter__'])


@_with_stubs(AGENT_PROTOCOL)
class Agent:

  device_obs = False   # True: policy() receives the staged device tensors

  def __init__(self, obs_space, act_space, config):
    pass


@_with_stubs(ENV_PROTOCOL, skip=('close',))
class Env:

  @property
  def obs_space(self):
    """dict of spaces; must hold is_first, is_last, is_terminal (usually reward
    too).  Keys starting with 'log/' reach neither the agent nor the replay."""
    raise NotImplementedError(f'{type(self).__name__}.obs_space -> dict of spaces')

  @property
  def act_space(self):
    """dict of spaces; must hold `reset` next to the actions."""
    raise NotImplementedError(f'{type(self).__name__}.act_space -> dict of spaces')

  def close(self):
    pass

  def __repr__(self):
    spaces = ', '.join(f'{p}={getattr(self, p)}' for p in ENV_SPACES)
    return f'{type(self).__name__}({spaces})'


@_with_stubs(STREAM_PROTOCOL)
class Stream:

  def __iter__(self):
    return self
