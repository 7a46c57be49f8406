"""ppo hyper-parameters (ppo/configs.yaml:94-109 `agent:` block), flattened, and the mapping
from the reference's nested block.  Options this build does not compute are refused by name."""


class Config(dict):
  __getattr__ = dict.__getitem__

  def update(self, *a, **kw):
    super().update(*a, **kw)
    return self


def make(**over):
  cfg = Config(
      # agent.enc.impala
      depth=32, mults=(1, 2, 2), outmult=16, blocks=2, enc_layers=5, enc_units=1024,
      enc_act='relu', enc_norm='none', symlog=True,
      # agent.rnn / actemb
      recurrent=True, rnn_units=1024, rnn_norm='layer', rnnact=True, actemb_units=1024,
      # agent.policy / agent.value
      pol_layers=0, pol_units=1024, val_layers=0, val_units=1024, head_act='relu',
      head_norm='layer', minstd=0.1, maxstd=1.0, pol_outscale=0.0, val_outscale=0.0,
      # agent.ppo_loss / loss_scales / valnorm / advnorm
      actent=1e-2, hor=200, lam=0.8, trclip=0.2, tarclip=10.0,
      scales=dict(policy=1.0, value=0.5), norm_rate=0.01, norm_limit=1e-8,
      # agent.opt
      lr=3e-4, eps=1e-7, clip=10.0, wd=0.0, warmup=1000, wdregex=r'/kernel$',
      replay_context=1, seed=0, compute_dtype='float32', graph='auto')
  cfg.update(over)
  return cfg


def debug(**over):
  """The `debug` block (ppo/configs.yaml:121-131)."""
  return make(depth=2, outmult=2, enc_layers=1, enc_units=8, rnn_units=8, actemb_units=8,
              pol_layers=1, pol_units=8, val_layers=1, val_units=8).update(over)


_FIXED = {           # dotted path in `agent` -> the only value computed here
    'enc.typ': 'impala', 'policy_dist_disc': 'categorical', 'policy_dist_cont': 'bounded_normal',
    'value.output': 'mse', 'valnorm.impl': 'meanstd', 'advnorm.impl': 'meanstd',
    'policy.unimix': 0.0, 'enc.impala.winit': 'trunc_normal_in', 'rnn.winit': 'trunc_normal_in',
}
_NESTED = {
    'depth': 'enc.impala.depth', 'mults': 'enc.impala.mults', 'outmult': 'enc.impala.outmult',
    'enc_layers': 'enc.impala.layers', 'enc_units': 'enc.impala.units', 'enc_act': 'enc.impala.act',
    'enc_norm': 'enc.impala.norm', 'symlog': 'enc.impala.symlog', 'recurrent': 'recurrent',
    'rnn_units': 'rnn.units', 'rnn_norm': 'rnn.norm', 'rnnact': 'rnnact', 'actemb_units': 'actemb.units',
    'pol_layers': 'policy.layers', 'pol_units': 'policy.units', 'head_act': 'policy.act',
    'head_norm': 'policy.norm', 'minstd': 'policy.minstd', 'maxstd': 'policy.maxstd',
    'pol_outscale': 'policy.outscale', 'val_layers': 'value.layers', 'val_units': 'value.units',
    'val_outscale': 'value.outscale', 'actent': 'ppo_loss.actent', 'hor': 'ppo_loss.hor',
    'lam': 'ppo_loss.lam', 'trclip': 'ppo_loss.trclip', 'tarclip': 'ppo_loss.tarclip',
    'norm_rate': 'valnorm.rate', 'norm_limit': 'valnorm.limit', 'lr': 'opt.lr', 'eps': 'opt.eps',
    'clip': 'opt.clip', 'wd': 'opt.wd', 'warmup': 'opt.warmup',
}


def _get(tree, path, default=None):
  for part in path.split('.'):
    if not hasattr(tree, 'get') or part not in tree:
      return default
    tree = tree[part]
  return tree


def from_reference(agent, **over):
  """Nested `agent:` block of ppo/configs.yaml -> flat Config."""
  refused = []
  for path, want in _FIXED.items():
    got = _get(agent, path, want)
    if got != want:
      refused.append(f'agent.{path}={got!r} (implemented: {want!r})')
  for a, b in (('valnorm.rate', 'advnorm.rate'), ('valnorm.limit', 'advnorm.limit'),
               ('policy.act', 'value.act'), ('policy.norm', 'value.norm')):
    if _get(agent, a) != _get(agent, b):
      refused.append(f'agent.{a} != agent.{b}')
  if refused:
    raise NotImplementedError('ppo: unsupported options: ' + '; '.join(refused))
  cfg = make()
  for name, path in _NESTED.items():
    value = _get(agent, path)
    if value is not None:
      cfg[name] = tuple(value) if isinstance(value, (list, tuple)) else type(cfg[name])(value)
  scales = _get(agent, 'loss_scales')
  if scales:
    cfg['scales'] = {k: float(v) for k, v in dict(scales).items()}
  cfg.update(over)
  return cfg
