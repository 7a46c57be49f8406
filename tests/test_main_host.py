"""Config-level drop-in (dreamerv3/main.py:22-31,127-272): a configs.yaml written for the
reference is consumed unchanged -- named blocks with dotted and regular-expression keys, YAML
anchors / merge keys, typed command-line flags -- and the factories wire Replay / Consec / env
wrappers the way the reference does."""
import pathlib

import numpy as np
import pytest

from embodied_b200 import elements
from embodied_b200.dreamerv3 import config as configlib, main as mainlib
import doubles

REF_YAML = pathlib.Path('/root/reference/dreamerv3/configs.yaml')
needs_ref = pytest.mark.skipif(not REF_YAML.exists(), reason='/root/reference not on this machine')

MINI_YAML = r"""
defaults:
  logdir: /tmp/x/{timestamp}
  replica: 0
  replicas: 1
  task: dummy_disc
  seed: 0
  script: train
  batch_size: 16
  batch_length: 64
  report_length: 32
  consec_train: 1
  consec_report: 1
  replay_context: 1
  random_agent: False
  logger: {outputs: [jsonl], filter: 'score', timer: True}
  env: {dummy: {}}
  replay: {size: 5e6, online: True, chunksize: 1024, fracs: {uniform: 1.0, priority: 0.0, recency: 0.0}}
  run: {steps: 1e10, train_ratio: 32.0, envs: 16, log_every: 120, debug: True}
  jax: {compute_dtype: bfloat16}
  agent:
    dyn: {typ: rssm, rssm: {deter: 8192, hidden: 1024, stoch: 32, classes: 64, blocks: 8}}
    enc: {typ: simple, simple: {depth: 64, mults: [2, 3, 4, 4], units: 1024, kernel: 5}}
    dec: {typ: simple, simple: {depth: 64, mults: [2, 3, 4, 4], units: 1024, kernel: 5, bspace: 8}}
    rewhead: {layers: 1, units: 1024, bins: 255}
    value: {layers: 3, units: 1024, bins: 255}

small: &small
  .*\.rssm: {deter: 512, hidden: 64, classes: 4}
  .*\.depth: 4
  .*\.units: 64

proprio:
  <<: *small
  task: dummy_cont
  env.dummy.length: 7
  run: {envs: 2, train_ratio: 1024}
"""


def test_blocks_regex_keys_anchors_and_flags(tmp_path):
  path = tmp_path / 'configs.yaml'
  path.write_text(MINI_YAML)
  config = mainlib.load_config(
      ['--configs', 'proprio', '--batch_size', '4', '--run.train_ratio=64', '--agent.enc.simple.mults',
       '1,2', '--random_agent', 'True'], configs_file=str(path))
  assert config.task == 'dummy_cont' and config.env.dummy.length == 7
  assert config.agent.dyn.rssm.deter == 512 and config.agent.dyn.rssm.stoch == 32
  assert config.agent.enc.simple.depth == config.agent.dec.simple.depth == 4
  assert {config.agent[k].units for k in ('rewhead', 'value')} == {64}
  assert config.agent.enc.simple.units == 64
  assert config.batch_size == 4 and isinstance(config.batch_size, int)
  assert config.run.train_ratio == 64.0 and config.run.envs == 2
  assert config.agent.enc.simple.mults == (1, 2)
  assert config.random_agent is True
  assert '{timestamp}' not in config.logdir
  with pytest.raises(KeyError):
    mainlib.load_config(['--no.such.flag', '1'], configs_file=str(path))
  with pytest.raises(KeyError):
    mainlib.load_config(['--configs', 'nope'], configs_file=str(path))


def test_builtin_tree_round_trips_to_the_flat_hyperparameters():
  for size in ('size1m', 'size12m', 'size200m'):
    config = mainlib.load_config(['--configs', size])
    flat = configlib.from_reference(config.agent)
    want = configlib.make(size)
    assert {k: flat[k] for k in want} == dict(want), size
  debug = mainlib.load_config(['--configs', 'debug'])
  flat = configlib.from_reference(debug.agent)
  assert (flat.deter, flat.hidden, flat.stoch, flat.classes, flat.blocks, flat.depth, flat.bins) == (
      8, 3, 2, 4, 4, 2, 5)
  assert debug.batch_size == 8 and debug.run.envs == 4 and debug.replay.size == 1e4


def test_unsupported_options_are_refused_by_name():
  config = mainlib.load_config(['--agent.ac_grads', 'True', '--agent.retnorm.impl', 'meanstd'])
  with pytest.raises(NotImplementedError, match='ac_grads.*retnorm.impl|retnorm.impl.*ac_grads'):
    configlib.from_reference(config.agent)
  with pytest.raises(NotImplementedError, match='eval_only'):
    mainlib.main(['--script', 'eval_only', '--logdir', '/tmp/never_made'])
  with pytest.raises(NotImplementedError, match='atari'):
    mainlib.make_env(mainlib.load_config(['--task', 'atari_pong']), 0)


@needs_ref
def test_reference_configs_yaml_is_consumed_unchanged():
  config = mainlib.load_config(['--configs', 'size12m', 'debug', '--run.envs', '3'],
                               configs_file=str(REF_YAML))
  # debug comes last: its regex keys win over size12m's (dreamerv3/configs.yaml:120-145, debug block)
  assert config.agent.dyn.rssm.deter == 8 and config.agent.dyn.rssm.hidden == 3
  assert config.agent.enc.simple.depth == 2 and config.agent.value.bins == 5
  assert config.batch_size == 8 and config.batch_length == 10 and config.run.envs == 3
  assert config.jax.platform == 'cpu'
  flat = configlib.from_reference(config.agent, seed=config.seed)
  assert (flat.deter, flat.stoch, flat.classes, flat.blocks, flat.units) == (8, 2, 4, 4, 8)
  assert flat.scales['image'] == 1.0 and flat.scales['rep'] == 0.1
  # every shipped block that only re-parameterises the agent / run parses; size blocks give the table
  for name, deter in (('size1m', 512), ('size50m', 4096), ('size200m', 8192), ('size400m', 12288)):
    c = mainlib.load_config(['--configs', name], configs_file=str(REF_YAML))
    f = configlib.from_reference(c.agent)
    assert f.deter == deter and dict(f) == dict(configlib.make(name, seed=f.seed)), name
  c = mainlib.load_config(['--configs', 'dmc_proprio'], configs_file=str(REF_YAML))
  assert c.env.dmc.image is False and c.agent.dyn.rssm.deter == 512 and c.run.train_ratio == 1024


def test_factories_wire_replay_stream_and_wrappers(tmp_path):
  config = mainlib.load_config(['--configs', 'debug', '--logdir', str(tmp_path), '--task', 'dummy_cont',
                                '--consec_train', '2'])
  args = mainlib.run_args(config)
  assert args.batch_size == 8 and args.train_ratio == 8 and args.logdir == str(tmp_path)
  env = mainlib.make_env(config, 0)
  names = []
  e = env
  while hasattr(e, 'env'):
    names.append(type(e).__name__)
    e = e.env
  assert names == ['ClipAction', 'CheckSpaces', 'UnifyDtypes', 'NormalizeAction'], names
  replay = mainlib.make_replay(config, 'replay', store=doubles.HostStore(1024, staging_rows=16))
  assert replay.length == 2 * 10 + 1 and replay.capacity == 10000 and replay.online
  assert str(replay.directory) == str(tmp_path / 'replay')
  report = mainlib.make_replay(config, 'eval_replay', 'report', store=doubles.HostStore(1024, 16))
  assert report.length == 5 + 1 and report.capacity == 1000
  stream = mainlib.make_stream(config, replay, 'train')
  assert (stream.length, stream.consec, stream.prefix, stream.strict) == (10, 2, 1, True)
  agent = mainlib.make_agent(config.update(random_agent=True))
  assert set(agent.act_space) == {'act_disc', 'act_cont'} or 'reset' not in agent.act_space


def test_replay_fracs_build_the_selector_mixture(tmp_path):
  """replay.fracs.uniform < 1 -> Mixture(Uniform, Prioritized(**replay.prio), Recency(1 / age ** recexp))
  (dreamerv3/main.py:196-206); fractions of 0 drop their selector (selectors.py:203-206)."""
  from embodied_b200.core import selectors
  import doubles
  config = mainlib.load_config(['--replay.fracs.uniform', '0.5', '--replay.fracs.priority', '0.5',
                                '--replay.size', '40000', '--logdir', str(tmp_path)])
  replay = mainlib.make_replay(config, 'replay', store=doubles.HostStore(1024, staging_rows=16))
  mix = replay.sampler
  assert isinstance(mix, selectors.Mixture)
  assert [type(s) for s in mix.selectors] == [selectors.Prioritized, selectors.Uniform]
  prio = mix.selectors[0]
  assert (prio.exponent, prio.maxfrac, prio.initial, prio.zero_on_sample) == (0.8, 0.5, float('inf'), True)
  # evaluation replays always sample uniformly (main.py:196 `mode == 'train'`)
  assert isinstance(mainlib.make_replay(config, 'eval_replay', 'eval',
                                        store=doubles.HostStore(1024, staging_rows=16)).sampler,
                    selectors.Uniform)


# ------------------------------------------------------------------------------ ppo/main.py
PPO_YAML = pathlib.Path('/root/reference/ppo/configs.yaml')


def test_yaml_12_floats():
  """`1e5`, `3e-4`, `inf` are numbers for the reference's ruamel.yaml reader (configs.yaml:41,109)."""
  tree = mainlib._yaml_load('a: 1e5\nb: 3e-4\nc: inf\nd: 16\ne: 1.5\nf: name\ng: [1, 2e2]\nh: True\n')
  assert tree == dict(a=1e5, b=3e-4, c=float('inf'), d=16, e=1.5, f='name', g=[1, 200.0], h=True)
  assert isinstance(tree['d'], int) and isinstance(tree['a'], float)


def test_ppo_builtin_tree_round_trips():
  from embodied_b200.ppo import config as pcfg, main as pmain
  config = pmain.load_config(['--logdir', '/tmp/x'])
  assert pcfg.from_reference(config.agent) == pcfg.make()
  assert (config.batch_size, config.batch_length, config.run.train_ratio, config.replay.size) == (16, 64, 3.0, 1e5)
  debug = pmain.load_config(['--logdir', '/tmp/x', '--configs', 'debug'])
  flat = pcfg.from_reference(debug.agent)
  assert flat == pcfg.debug(), {k: (flat[k], pcfg.debug()[k]) for k in flat if flat[k] != pcfg.debug()[k]}
  with pytest.raises(NotImplementedError, match='policy_dist_cont'):
    pcfg.from_reference(config.update({'agent.policy_dist_cont': 'normal_logstd'}).agent)


@pytest.mark.skipif(not PPO_YAML.exists(), reason='/root/reference not on this machine')
def test_ppo_reference_configs_yaml_is_consumed_unchanged():
  from embodied_b200.ppo import config as pcfg, main as pmain
  ref = pmain.load_config(['--logdir', '/tmp/x'], configs_file=str(PPO_YAML))
  own = pmain.load_config(['--logdir', '/tmp/x'])
  assert pcfg.from_reference(ref.agent) == pcfg.from_reference(own.agent)
  for key in ('batch_size', 'batch_length', 'replay_context', 'consec_train'):
    assert ref[key] == own[key], key
  assert ref.replay.size == own.replay.size == 1e5 and ref.run.train_ratio == own.run.train_ratio
  dbg = pmain.load_config(['--logdir', '/tmp/x', '--configs', 'debug'], configs_file=str(PPO_YAML))
  assert pcfg.from_reference(dbg.agent) == pcfg.debug()
  assert (dbg.batch_size, dbg.batch_length, dbg.run.envs) == (8, 12, 4)
