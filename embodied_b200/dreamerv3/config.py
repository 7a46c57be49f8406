"""dreamerv3 hyper-parameters (dreamerv3/configs.yaml:81-145), flattened.

`make(size='size200m', **overrides)`; the size presets are the reference's
regex blocks (configs.yaml:120-145: `.*\\.rssm`, `.*\\.depth`, `.*\\.units`).
"""

SIZES = {
    'size1m': dict(deter=512, hidden=64, classes=4, depth=4, units=64),
    'size12m': dict(deter=2048, hidden=256, classes=16, depth=16, units=256),
    'size25m': dict(deter=3072, hidden=384, classes=24, depth=24, units=384),
    'size50m': dict(deter=4096, hidden=512, classes=32, depth=32, units=512),
    'size100m': dict(deter=6144, hidden=768, classes=48, depth=48, units=768),
    'size200m': dict(deter=8192, hidden=1024, classes=64, depth=64, units=1024),
    'size400m': dict(deter=12288, hidden=1536, classes=96, depth=96, units=1536),
}


class Config(dict):
  __getattr__ = dict.__getitem__

  def update(self, *a, **kw):
    super().update(*a, **kw)
    return self


def make(size='size200m', **over):
  cfg = Config(
      # agent.dyn.rssm
      deter=8192, hidden=1024, stoch=32, classes=64, blocks=8, unimix=0.01,
      free_nats=1.0, imglayers=2, obslayers=1, dynlayers=1,
      # agent.enc.simple / agent.dec.simple
      depth=64, mults=(2, 3, 4, 4), kernel=5, units=1024, bspace=8,
      # heads
      bins=255, rew_layers=1, con_layers=1, pol_layers=3, val_layers=3,
      # imagination / losses
      imag_length=15, horizon=333, contdisc=True, lam=0.95, actent=3e-4,
      slowreg=1.0, slowrate=0.02, replay_context=1,
      retnorm_rate=0.01, retnorm_limit=1.0, perclo=5.0, perchi=95.0,
      scales=dict(image=1.0, rew=1.0, con=1.0, dyn=1.0, rep=0.1,
                  policy=1.0, value=1.0, repval=0.3),
      # agent.opt
      lr=4e-5, agc=0.3, eps=1e-20, beta1=0.9, beta2=0.999, warmup=1000,
      pmin=1e-3,
      # spaces (filled by the Agent from obs_space / act_space)
      image=(64, 64, 3), actions=5,
      # runtime (the `jax:` block's role): compute dtype and seed
      compute_dtype='bfloat16', seed=0,
  )
  cfg.update(SIZES[size])
  cfg.update(over)
  return cfg


# ---------------------------------------------------------------------------
# The reference's nested `agent:` block (dreamerv3/configs.yaml:81-118) <-> the flat
# hyper-parameters above.  `schema()` generates the nested defaults (so a run without
# a YAML file has the same tree to apply `--agent.dyn.rssm.deter 512`-style flags and
# `.*\.units`-style regex blocks to); `from_reference()` reads a nested block back.

_NESTED = {          # flat name -> dotted path inside `agent`
    'deter': 'dyn.rssm.deter', 'hidden': 'dyn.rssm.hidden', 'stoch': 'dyn.rssm.stoch',
    'classes': 'dyn.rssm.classes', 'blocks': 'dyn.rssm.blocks', 'unimix': 'dyn.rssm.unimix',
    'free_nats': 'dyn.rssm.free_nats', 'imglayers': 'dyn.rssm.imglayers',
    'obslayers': 'dyn.rssm.obslayers', 'dynlayers': 'dyn.rssm.dynlayers',
    'depth': 'enc.simple.depth', 'mults': 'enc.simple.mults', 'kernel': 'enc.simple.kernel',
    'units': 'enc.simple.units', 'bspace': 'dec.simple.bspace',
    'bins': 'rewhead.bins', 'rew_layers': 'rewhead.layers', 'con_layers': 'conhead.layers',
    'pol_layers': 'policy.layers', 'val_layers': 'value.layers',
    'imag_length': 'imag_length', 'horizon': 'horizon', 'contdisc': 'contdisc',
    'lam': 'imag_loss.lam', 'actent': 'imag_loss.actent', 'slowreg': 'imag_loss.slowreg',
    'slowrate': 'slowvalue.rate', 'retnorm_rate': 'retnorm.rate', 'retnorm_limit': 'retnorm.limit',
    'perclo': 'retnorm.perclo', 'perchi': 'retnorm.perchi',
    'lr': 'opt.lr', 'agc': 'opt.agc', 'eps': 'opt.eps', 'beta1': 'opt.beta1', 'beta2': 'opt.beta2',
    'warmup': 'opt.warmup',
}
# places that must agree with a flat value (the reference repeats them per module)
_MIRRORS = {
    'depth': ('dec.simple.depth',), 'mults': ('dec.simple.mults',), 'kernel': ('dec.simple.kernel',),
    'units': ('dec.simple.units', 'rewhead.units', 'conhead.units', 'policy.units', 'value.units'),
    'bins': ('value.bins',), 'lam': ('repl_loss.lam',), 'slowreg': ('repl_loss.slowreg',),
}
# options of the reference this build implements only at one value: anything else is refused
_FIXED = {
    'ac_grads': False, 'dyn.typ': 'rssm', 'enc.typ': 'simple', 'dec.typ': 'simple',
    'dyn.rssm.act': 'silu', 'dyn.rssm.norm': 'rms', 'dyn.rssm.absolute': False,
    'dyn.rssm.obslayers': 1, 'dyn.rssm.dynlayers': 1,
    'enc.simple.act': 'silu', 'enc.simple.norm': 'rms', 'enc.simple.outer': False,
    'enc.simple.strided': False, 'dec.simple.act': 'silu', 'dec.simple.norm': 'rms',
    'dec.simple.outer': False, 'dec.simple.strided': False,
    'rewhead.output': 'symexp_twohot', 'conhead.output': 'binary', 'value.output': 'symexp_twohot',
    'policy_dist_disc': 'categorical', 'imag_last': 0,
    'imag_loss.slowtar': False, 'repl_loss.slowtar': False, 'slowvalue.every': 1,
    'retnorm.impl': 'perc', 'retnorm.debias': False, 'valnorm.impl': 'none', 'advnorm.impl': 'none',
    'reward_grad': True, 'repval_loss': True, 'repval_grad': True,
    'opt.momentum': True, 'opt.wd': 0.0, 'opt.schedule': 'const', 'opt.anneal': 0,
}


def _put(tree, path, value):
  *parents, leaf = path.split('.')
  for p in parents:
    tree = tree.setdefault(p, {})
  tree[leaf] = value


def _get(tree, path, default=None):
  for p in path.split('.'):
    if not isinstance(tree, dict) or p not in tree:
      return default
    tree = tree[p]
  return tree


def schema(size='size200m'):
  """The nested `agent` block with this build's defaults at `size`."""
  flat = make(size)
  tree = {}
  for name, path in _NESTED.items():
    value = flat[name]
    _put(tree, path, list(value) if isinstance(value, tuple) else value)
    for mirror in _MIRRORS.get(name, ()):
      _put(tree, mirror, list(value) if isinstance(value, tuple) else value)
  for path, value in _FIXED.items():
    if _get(tree, path) is None:
      _put(tree, path, value)
  scales = dict(flat['scales'])
  scales['rec'] = scales.pop('image')            # configs.yaml:86 calls the image term `rec`
  tree['loss_scales'] = scales
  for head, outscale in (('rewhead', 0.0), ('conhead', 1.0), ('policy', 0.01), ('value', 0.0)):
    _put(tree, f'{head}.outscale', outscale)
  return tree


def from_reference(agent, **over):
  """Nested `agent` block (a dict / elements.Config in the reference's layout) -> flat Config.
  Raises NotImplementedError for options outside what this build computes."""
  agent = agent._plain() if hasattr(agent, '_plain') else dict(agent)
  refused = []
  for path, want in _FIXED.items():
    got = _get(agent, path, want)
    if isinstance(want, float):
      bad = abs(float(got) - want) > 1e-12
    else:
      bad = got != want
    if bad:
      refused.append(f'agent.{path}={got!r} (implemented: {want!r})')
  for name, mirrors in _MIRRORS.items():
    base = _get(agent, _NESTED[name])
    for m in mirrors:
      got = _get(agent, m, base)
      norm = lambda v: list(v) if isinstance(v, (tuple, list)) else v
      if base is not None and norm(got) != norm(base):
        refused.append(f'agent.{m}={got!r} differs from agent.{_NESTED[name]}={base!r}')
  if refused:
    raise NotImplementedError(
        'this build implements the shipped dreamerv3 configuration family; unsupported: ' +
        '; '.join(refused))
  cfg = make('size200m')
  for name, path in _NESTED.items():
    value = _get(agent, path)
    if value is not None:
      cfg[name] = tuple(value) if isinstance(value, (list, tuple)) else type(cfg[name])(value)
  scales = _get(agent, 'loss_scales')
  if scales:
    scales = dict(scales)
    if 'rec' in scales:
      scales['image'] = scales.pop('rec')
    cfg['scales'] = {k: float(v) for k, v in scales.items()}
  for k in ('seed', 'replay_context'):
    if k in agent:
      cfg[k] = int(agent[k])
  jax = agent.get('jax') or {}
  if 'compute_dtype' in jax:
    cfg['compute_dtype'] = jax['compute_dtype']
  cfg.update(over)
  return cfg
