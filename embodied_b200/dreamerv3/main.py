"""Config-level entry point: the reference's `dreamerv3/main.py` contract on this runtime.

    python -m embodied_b200.dreamerv3.main --configs size12m debug --task dummy_disc \\
        --run.envs 4 --logdir /tmp/run  [--configs-file /path/to/configs.yaml]

A `configs.yaml` written for the reference (dreamerv3/configs.yaml: a `defaults` block plus named
blocks whose keys may be dotted paths or regular expressions such as `.*\\.rssm`, YAML anchors
and merge keys included) is consumed unchanged: named blocks are applied in order, then
`--dotted.key value` flags typed by the default they replace (main.py:23-31).  Without a file
the same tree is generated from `dreamerv3/config.py`.  The factories keep the reference's
names and wiring (main.py:127-272): `make_agent`, `make_replay`, `make_env`, `wrap_env`,
`make_stream`, `make_logger`; `main()` hands them to `embodied_b200.run.train`.

Only what the hot-path scope covers is constructed here: scripts other than `train`, the
simulator adapters of `embodied/envs/` (besides `dummy`) and non-uniform replay selectors
raise with the reference location they correspond to.
"""
import functools
import sys

import numpy as np

from .. import elements
from ..core import replay as replaylib
from ..core import streams, wrappers
from ..core.random import RandomAgent
from . import config as configlib

SIZE_BLOCKS = {
    name: {r'.*\.rssm': {k: spec[k] for k in ('deter', 'hidden', 'classes')},
           r'.*\.depth': spec['depth'], r'.*\.units': spec['units']}
    for name, spec in configlib.SIZES.items()}


def builtin_configs():
  """`defaults` + the size blocks, in the reference's layout (configs.yaml:1-145)."""
  defaults = dict(
      logdir='~/logdir/{timestamp}', replica=0, replicas=1, task='dummy_disc', seed=0,
      script='train', batch_size=16, batch_length=64, report_length=32, consec_train=1,
      consec_report=1, replay_context=1, random_agent=False,
      logger=dict(outputs=['jsonl'], filter='score|length|fps|ratio|train/loss/', timer=True),
      env=dict(dummy=dict(), synthetic=dict(size=[64, 64, 3], classes=5, length=500)),
      replay=dict(size=5e6, online=True, chunksize=1024,
                  fracs=dict(uniform=1.0, priority=0.0, recency=0.0),
                  prio=dict(exponent=0.8, maxfrac=0.5, initial=float('inf'), zero_on_sample=True),
                  priosignal='model', recexp=1.0),
      run=dict(steps=1e10, train_ratio=32.0, log_every=120, report_every=300, save_every=900,
               envs=16, eval_envs=4, eval_eps=1, report_batches=1, from_checkpoint='', debug=True,
               usage=dict(psutil=True)),
      jax=dict(platform='cuda', compute_dtype='bfloat16'),
      agent=configlib.schema('size200m'))
  debug = {
      'batch_size': 8, 'batch_length': 10, 'report_length': 5, 'replay.size': 1e4,
      'run': dict(envs=4, report_every=10, log_every=5, save_every=15, train_ratio=8, debug=True),
      'agent': {r'.*\.bins': 5, r'.*\.layers': 1, r'.*\.units': 8, r'.*\.stoch': 2, r'.*\.classes': 4,
                r'.*\.deter': 8, r'.*\.hidden': 3, r'.*\.blocks': 4, r'.*\.depth': 2}}
  return {'defaults': defaults, 'debug': debug, **SIZE_BLOCKS}


def parse_flags(argv):
  """['--a.b', '1', '--c=x', '--names', 'p', 'q'] -> {'a.b': ['1'], 'c': ['x'], 'names': ['p', 'q']}."""
  out, key = {}, None
  for token in argv:
    if token.startswith('--'):
      key, _, value = token[2:].partition('=')
      out[key] = [value] if value else []
    elif key is None:
      raise ValueError(f'positional argument {token!r}: flags are --name value')
    else:
      out[key].append(token)
  return out


def _typed(old, tokens, key):
  """Parse flag tokens with the type of the default they replace."""
  def one(proto, text):
    if isinstance(proto, bool):
      if text.lower() in ('true', '1', 'yes'):
        return True
      if text.lower() in ('false', '0', 'no'):
        return False
      raise ValueError(f'--{key}: expected a boolean, got {text!r}')
    if isinstance(proto, int):
      return int(float(text))
    if isinstance(proto, float):
      return float(text)
    return text
  if isinstance(old, (tuple, list)):
    proto = old[0] if len(old) else ''
    return tuple(one(proto, t) for text in tokens for t in text.split(',') if t != '')
  if len(tokens) != 1:
    raise ValueError(f'--{key} takes one value, got {tokens}')
  return one(old, tokens[0])


def _yaml_load(text):
  """safe_load with YAML 1.2 floats: the reference reads its configs with ruamel.yaml, for which
  `1e5` / `3e-4` / `inf` are numbers; PyYAML's 1.1 resolver leaves them strings."""
  import re
  import yaml

  class Loader(yaml.SafeLoader):
    pass
  Loader.add_implicit_resolver(
      'tag:yaml.org,2002:float',
      re.compile(r'^[-+]?(?:[0-9][0-9_]*\.?[0-9_]*|\.[0-9_]+)(?:[eE][-+]?[0-9]+)?$|^[-+]?\.?(?:inf|Inf|INF)$|^\.?(?:nan|NaN|NAN)$'),
      list('-+0123456789.in'))
  def as_float(loader, node):
    text = loader.construct_scalar(node).replace('_', '').lower().replace('.inf', 'inf').replace('.nan', 'nan')
    return float(text)
  Loader.add_constructor('tag:yaml.org,2002:float', as_float)
  return yaml.load(text, Loader=Loader)


def load_config(argv=None, configs_file=None, builtin=None):
  """The run configuration as an elements.Config (main.py:22-31)."""
  flags = parse_flags(list(sys.argv[1:] if argv is None else argv))
  configs_file = (flags.pop('configs-file', None) or [configs_file])[0]
  if configs_file:
    blocks = _yaml_load(elements.Path(configs_file).read())
  else:
    blocks = (builtin or builtin_configs)()
  config = elements.Config(blocks['defaults'])
  for name in flags.pop('configs', []):
    if name == 'defaults':
      continue
    if name not in blocks:
      raise KeyError(f'no config block {name!r}; available: {sorted(blocks)}')
    config = config.update(blocks[name])
  flat = config.flat
  for key, tokens in flags.items():
    if key not in flat:
      raise KeyError(f'unknown flag --{key}')
    config = config.update({key: _typed(flat[key], tokens, key)})
  return config.update(logdir=str(config.logdir).format(timestamp=elements.timestamp()))


def run_args(config):
  """The `args` of run.train (main.py:55-66)."""
  return elements.Config(
      **config.run, replica=config.replica, replicas=config.replicas, logdir=config.logdir,
      batch_size=config.batch_size, batch_length=config.batch_length,
      report_length=config.report_length, consec_train=config.consec_train,
      consec_report=config.consec_report, replay_context=config.replay_context)


# ------------------------------------------------------------------- factories
def make_env(config, index, **overrides):                   # main.py:199-241
  suite, task = config.task.split('_', 1)
  kwargs = dict(config.env.get(suite, {}))
  kwargs.update(overrides)
  if suite == 'dummy':
    from ..envs import dummy
    env = dummy.Dummy(task, **kwargs)
  elif suite == 'synthetic':
    from ..envs import synthetic
    kwargs['size'] = tuple(kwargs.get('size', (64, 64, 3)))
    env = synthetic.SyntheticImage(index, **kwargs)
  else:
    raise NotImplementedError(
        f"environment suite {suite!r}: the simulator adapters of embodied/envs/ are outside "
        "the hot-path scope (SURVEY.md 8); wrap your simulator in an embodied Env "
        "(embodied_b200.core.base.Env) and pass make_env to run.train directly")
  return wrap_env(env, config)


def wrap_env(env, config):                                   # main.py:244-253
  for name, space in env.act_space.items():
    if not space.discrete:
      env = wrappers.NormalizeAction(env, name)
  env = wrappers.UnifyDtypes(env)
  env = wrappers.CheckSpaces(env)
  for name, space in env.act_space.items():
    if not space.discrete:
      env = wrappers.ClipAction(env, name)
  return env


def make_agent(config):                                      # main.py:127-150
  env = make_env(config, 0)
  obs_space = {k: v for k, v in env.obs_space.items() if not k.startswith('log/')}
  act_space = {k: v for k, v in env.act_space.items() if k != 'reset'}
  env.close()
  if config.random_agent:
    return RandomAgent(obs_space, act_space)
  from .agent import Agent
  return Agent(obs_space, act_space, configlib.from_reference(
      config.agent, seed=config.seed, replay_context=config.replay_context,
      compute_dtype=config.jax.compute_dtype))


def make_replay(config, folder, mode='train', **kwargs):     # main.py:183-196
  batlen = config.batch_length if mode == 'train' else config.report_length
  consec = config.consec_train if mode == 'train' else config.consec_report
  capacity = config.replay.size if mode == 'train' else config.replay.size / 10
  length = consec * batlen + config.replay_context
  assert config.batch_size * length <= capacity, (config.batch_size, length, capacity)
  if config.replay.fracs.uniform < 1 and mode == 'train':                 # main.py:196-206
    assert config.jax.compute_dtype in ('bfloat16', 'float32'), config.jax.compute_dtype
    from ..core import selectors
    recency = 1.0 / np.arange(1, int(capacity) + 1) ** config.replay.recexp
    kwargs['selector'] = selectors.Mixture(dict(
        uniform=selectors.Uniform(),
        priority=selectors.Prioritized(**config.replay.prio),
        recency=selectors.Recency(recency),
    ), dict(config.replay.fracs))
  directory = elements.Path(config.logdir) / folder
  if config.replicas > 1:
    directory /= f'{config.replica:05}'
  kwargs.setdefault('workers', int(config.run.envs))
  kwargs.setdefault('staging_rows', max(int(config.run.envs), 16))
  return replaylib.Replay(
      length=length, capacity=int(capacity), online=config.replay.online,
      chunksize=config.replay.chunksize, directory=directory, **kwargs)


def make_stream(config, replay, mode):                       # main.py:256-268
  fn = functools.partial(replay.sample, config.batch_size, mode)
  stream = streams.Stateless(fn)
  return streams.Consec(
      stream,
      length=config.batch_length if mode == 'train' else config.report_length,
      consec=config.consec_train if mode == 'train' else config.consec_report,
      prefix=config.replay_context, strict=(mode == 'train'), contiguous=True)


def make_logger(config):                                     # main.py:153-180
  outputs = [elements.logger.TerminalOutput(config.logger.filter, 'Agent')]
  for output in config.logger.outputs:
    if output == 'jsonl':
      outputs.append(elements.logger.JSONLOutput(config.logdir, 'metrics.jsonl'))
      outputs.append(elements.logger.JSONLOutput(config.logdir, 'scores.jsonl', 'episode/score'))
    elif output in ('scope', 'tensorboard', 'wandb', 'expa'):
      print(f"logger output '{output}' is not available in this build; skipped")
    else:
      raise NotImplementedError(output)
  multiplier = config.env.get(config.task.split('_')[0], {}).get('repeat', 1)
  return elements.Logger(elements.Counter(), outputs, multiplier)


def main(argv=None):
  config = load_config(argv)
  logdir = elements.Path(config.logdir)
  print('Logdir:', logdir)
  if config.script not in ('train', 'train_eval'):
    raise NotImplementedError(
        f"script {config.script!r}: `train` (embodied/run/train.py) and `train_eval` "
        "(embodied/run/train_eval.py) are built; eval_only / parallel* are SURVEY.md section 8 "
        "'out of scope' (portal RPC)")
  logdir.mkdir()
  import yaml
  (logdir / 'config.yaml').write(yaml.safe_dump(_plain(config)))
  from .. import run
  bind = functools.partial
  if config.script == 'train':
    run.train(bind(make_agent, config), bind(make_replay, config, 'replay'), bind(make_env, config),
              bind(make_stream, config), bind(make_logger, config), run_args(config))
  else:                                                                   # main.py:77-86
    run.train_eval(
        bind(make_agent, config), bind(make_replay, config, 'replay'),
        bind(make_replay, config, 'eval_replay', 'eval'), bind(make_env, config),
        bind(make_env, config), bind(make_stream, config), bind(make_logger, config),
        run_args(config))


def _plain(config):
  out = {}
  for k, v in config.items():
    if hasattr(v, 'items'):
      out[k] = _plain(v)
    elif isinstance(v, tuple):
      out[k] = [x.item() if isinstance(x, np.generic) else x for x in v]
    else:
      out[k] = v.item() if isinstance(v, np.generic) else v
  return out


if __name__ == '__main__':
  main()
