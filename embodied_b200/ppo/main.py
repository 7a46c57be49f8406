"""Config-level entry point of ppo: the reference's `ppo/main.py` contract on this runtime.

    python -m embodied_b200.ppo.main --configs debug --task dummy_disc --logdir /tmp/run \\
        [--configs-file /path/to/ppo/configs.yaml]

The reference's two `main.py` files are the same program around a different agent import (they
differ in one `errfile` argument), so this module reuses the loader and factories of
`embodied_b200.dreamerv3.main` -- regex / dotted config blocks, typed flags, `make_replay`,
`make_env`, `wrap_env`, `make_stream`, `make_logger` -- and swaps `make_agent` and the built-in
`defaults` (ppo/configs.yaml:1-109: batch 16 x 64, replay.size 1e5, train_ratio 3, the `agent:`
block of ppo/config.py).  A `configs.yaml` written for the reference is consumed unchanged.
"""
import functools

from .. import elements
from ..core.random import RandomAgent
from ..dreamerv3 import main as shared
from . import config as configlib

make_env, wrap_env = shared.make_env, shared.wrap_env
make_replay, make_stream, make_logger = shared.make_replay, shared.make_stream, shared.make_logger
run_args = shared.run_args


def agent_schema():
  """The nested `agent:` block (ppo/configs.yaml:94-109) generated from the flat defaults."""
  flat = configlib.make()
  tree = {}
  for name, path in configlib._NESTED.items():
    node = tree
    *parents, leaf = path.split('.')
    for part in parents:
      node = node.setdefault(part, {})
    value = flat[name]
    node[leaf] = list(value) if isinstance(value, tuple) else value
  for path, value in configlib._FIXED.items():
    node = tree
    *parents, leaf = path.split('.')
    for part in parents:
      node = node.setdefault(part, {})
    node[leaf] = value
  tree['value']['act'], tree['value']['norm'] = tree['policy']['act'], tree['policy']['norm']
  tree['advnorm'] = dict(tree['valnorm'])
  tree['loss_scales'] = dict(flat['scales'])
  return tree


def builtin_configs():
  base = shared.builtin_configs()['defaults']
  defaults = dict(base)
  defaults.update(
      batch_size=16, batch_length=64, report_length=32,
      replay=dict(base['replay'], size=1e5),
      run=dict(base['run'], train_ratio=3.0),
      agent=agent_schema())
  debug = {
      'batch_size': 8, 'batch_length': 12, 'report_length': 12, 'replay.size': 1e4,
      'jax.compute_dtype': 'float32',
      'run': dict(envs=4, report_every=10, log_every=5, save_every=15, train_ratio=8, debug=True),
      'agent': {'enc.impala': dict(depth=2, outmult=2), r'.*\.layers': 1, r'.*\.units': 8}}
  return {'defaults': defaults, 'debug': debug}


def load_config(argv=None, configs_file=None):
  return shared.load_config(argv, configs_file, builtin=builtin_configs)


def make_agent(config):                                      # ppo/main.py:127-150
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


def main(argv=None):
  config = load_config(argv)
  logdir = elements.Path(config.logdir)
  print('Logdir:', logdir)
  if config.script not in ('train', 'train_eval'):
    raise NotImplementedError(f'script {config.script!r}: train / train_eval are built')
  logdir.mkdir()
  import yaml
  (logdir / 'config.yaml').write(yaml.safe_dump(shared._plain(config)))
  from .. import run
  bind = functools.partial
  if config.script == 'train':
    run.train(bind(make_agent, config), bind(make_replay, config, 'replay'), bind(make_env, config),
              bind(make_stream, config), bind(make_logger, config), run_args(config))
  else:
    run.train_eval(
        bind(make_agent, config), bind(make_replay, config, 'replay'),
        bind(make_replay, config, 'eval_replay', 'eval'), bind(make_env, config),
        bind(make_env, config), bind(make_stream, config), bind(make_logger, config),
        run_args(config))


if __name__ == '__main__':
  main()
