"""TEST INFRASTRUCTURE -- CPU restatement (fp32 torch) of the reference's ppo agent.

Follows, function by function: ppo/agent.py:19-235 (Agent.policy / train / loss, Model,
ppo_loss, the optax chain of _make_opt), ppo/nets.py:10-71 (ImpalaEncoder) and the pieces of
embodied/jax they call: nets.py:230-251 (Linear), :284-323 (Conv2D NHWC/HWIO SAME), :361-399
(Norm 'none' / 'layer'), :503-562 (DictEmbed, one-hot impl), :565-587 (MLP), :634-669 (GRU),
:76-100 (mask / available), heads.py:16-155 (MLPHead, categorical / bounded_normal / mse
heads), outs.py:129-141,161-179,208-234 (MSE, Normal, Categorical), utils.py:16-91
(Normalize 'meanstd', debias), opt.py:16-80 (Optimizer wrapper).

**Parity unpinned**: JAX cannot be installed here and the reference has no test that touches
ppo numerics, so this file is a restatement checked only by structural self-tests
(tests/test_ppo_oracle.py).  Sampling noise is INJECTED (Gumbel for categorical actions,
standard normal for Normal.sample) instead of jax.random.  Third-party arithmetic restated
from its published definition: optax `clip_by_global_norm`, `scale_by_adam` (b1 0.9, b2
0.999, bias-corrected, eps outside the root), `add_decayed_weights`, `scale_by_learning_rate`
over `linear_schedule(0, lr, warmup)` (the schedule's own count starts at 0: the first update
has learning rate 0).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math
import re

import numpy as np
import torch
import torch.nn.functional as F

f32 = torch.float32
EXCLUDE = ('is_first', 'is_last', 'is_terminal', 'reward')


class Config(dict):
  __getattr__ = dict.__getitem__


def default_config(**over):
  """ppo/configs.yaml:94-109 `agent:` block, flattened."""
  cfg = Config(
      depth=32, mults=(1, 2, 2), outmult=16, blocks=2, enc_layers=5, enc_units=1024,
      enc_act='relu', enc_norm='none', symlog=True,
      recurrent=True, rnn_units=1024, rnn_norm='layer', rnnact=True, actemb_units=1024,
      pol_layers=0, pol_units=1024, val_layers=0, val_units=1024, head_act='relu', head_norm='layer',
      minstd=0.1, maxstd=1.0, pol_outscale=0.0, val_outscale=0.0,
      actent=1e-2, hor=200, lam=0.8, trclip=0.2, tarclip=10.0,
      scales=dict(policy=1.0, value=0.5), norm_rate=0.01, norm_limit=1e-8,
      lr=3e-4, eps=1e-7, clip=10.0, wd=0.0, warmup=1000, wdregex=r'/kernel$',
      replay_context=1, norm_eps=1e-4)
  cfg.update(over)
  return cfg


def tiny_config(**over):
  """The `debug` block of ppo/configs.yaml:121-131: depth 2, outmult 2, layers 1, units 8."""
  cfg = default_config(depth=2, outmult=2, enc_layers=1, enc_units=8, rnn_units=8, actemb_units=8,
                       pol_layers=1, pol_units=8, val_layers=1, val_units=8)
  cfg.update(over)
  return cfg


# ------------------------------------------------------------------ spaces
def space_kind(space):
  """('disc', shape, classes) / ('cont', shape, None) for an elements-like Space."""
  if space.discrete:
    return 'disc', tuple(space.shape), int(np.asarray(space.classes).max())
  return 'cont', tuple(space.shape), None


def enc_spaces(obs_space):
  spaces = {k: v for k, v in obs_space.items() if k not in EXCLUDE}
  vec = {k: v for k, v in spaces.items() if len(v.shape) <= 2}
  img = {k: v for k, v in spaces.items() if len(v.shape) == 3}
  return vec, img


# ------------------------------------------------------------------ layers
def linear(p, name, x):                                          # nets.py:240-248
  return x @ p[f'{name}/kernel'] + p[f'{name}/bias']


def norm(p, name, x, impl, eps=1e-4):                            # nets.py:374-399
  if impl == 'none':
    return x
  assert impl == 'layer', impl
  mean = x.mean(-1, keepdim=True)
  mean2 = (x * x).mean(-1, keepdim=True)
  var = torch.clamp(mean2 - mean * mean, min=0)
  return (x - mean) * (torch.rsqrt(var + eps) * p[f'{name}/scale']) + p[f'{name}/shift']


def act(name, x):
  return {'relu': torch.relu, 'silu': F.silu, 'none': lambda y: y}[name](x)


def conv3(p, name, x):                                           # nets.py:299-320 (NHWC in, HWIO kernel)
  w = p[f'{name}/kernel'].permute(3, 2, 0, 1)
  y = F.conv2d(x.permute(0, 3, 1, 2), w, p[f'{name}/bias'], padding=1)
  return y.permute(0, 2, 3, 1)


def maxpool_3x3_s2_same(x):                                      # ppo/nets.py:49-50 (reduce_window, init -inf)
  def pads(n):
    out = -(-n // 2)
    total = max((out - 1) * 2 + 3 - n, 0)
    return total // 2, total - total // 2
  (t, b), (l, r) = pads(x.shape[1]), pads(x.shape[2])
  y = F.pad(x.permute(0, 3, 1, 2), (l, r, t, b), value=float('-inf'))
  return F.max_pool2d(y, 3, 2).permute(0, 2, 3, 1)


def symlog(x):
  return torch.sign(x) * torch.log1p(torch.abs(x))


def available(x, bdims):                                         # nets.py:80-100
  if x.dtype.is_floating_point:
    m = x != float('-inf')
  elif x.dtype in (torch.int8, torch.int16, torch.int32, torch.int64):
    m = x != -1
  else:
    return torch.ones(x.shape[:bdims], dtype=torch.bool)
  return m.reshape(*x.shape[:bdims], -1).all(-1)


def dict_embed(p, name, spaces, xs, bshape, units, squish):      # nets.py:520-562
  total = p[f'{name}/init'].expand(*bshape, units)
  for key in sorted(spaces):
    kind, shape, classes = space_kind(spaces[key])
    x = xs[key]
    m = available(x, len(bshape))
    x = torch.where(m.reshape(*bshape, *([1] * (x.ndim - len(bshape)))), x, torch.zeros_like(x))
    if kind == 'disc':
      x = F.one_hot(x.long(), classes).to(f32)
    else:
      x = squish(x.to(f32))
    x = linear(p, f'{name}/{key}', x.reshape(*bshape, -1))
    total = total + torch.where(m[..., None], x, torch.zeros_like(x))
  return total


class PPO:
  """Functional model + training state of ppo/agent.py.  `p`: dict name -> fp32 tensor."""

  def __init__(self, cfg, obs_space, act_space, params):
    self.cfg, self.obs_space, self.act_space = cfg, obs_space, act_space
    self.vec, self.img = enc_spaces(obs_space)
    self.p = params
    self.opt = dict(count=0, mu={k: torch.zeros_like(v) for k, v in params.items()},
                    nu={k: torch.zeros_like(v) for k, v in params.items()})
    self.norms = {n: dict(mean=0.0, sqrs=0.0, corr=0.0) for n in ('advnorm', 'valnorm')}

  # -- ppo/nets.py:29-71 ---------------------------------------------------------
  def encoder(self, p, obs, bdims):
    cfg = self.cfg
    bshape = tuple(next(iter(obs.values())).shape[:bdims])
    outs = []
    if self.vec:
      squish = symlog if cfg.symlog else (lambda y: y)
      x = dict_embed(p, 'enc/emb', self.vec, obs, bshape, cfg.enc_units, squish)
      x = x.reshape(-1, x.shape[-1])
      for i in range(cfg.enc_layers - 1):
        x = linear(p, f'enc/mlp/linear{i}', x)
        x = act(cfg.enc_act, norm(p, f'enc/mlp/norm{i}', x, cfg.enc_norm))
      outs.append(x)
    if self.img:
      x = torch.cat([obs[k] for k in sorted(self.img)], -1)
      assert x.dtype == torch.uint8
      x = x.to(f32) * 255 - 0.5                                   # sic (ppo/nets.py:46)
      x = x.reshape(-1, *x.shape[-3:])
      for s, mult in enumerate(cfg.mults):
        x = conv3(p, f'enc/s{s}in', x)
        x = maxpool_3x3_s2_same(x)
        for b in range(cfg.blocks):
          skip = x
          x = act(cfg.enc_act, norm(p, f'enc/s{s}b{b}n1', x, cfg.enc_norm))
          x = conv3(p, f'enc/s{s}b{b}c1', x)
          x = act(cfg.enc_act, norm(p, f'enc/s{s}b{b}n2', x, cfg.enc_norm))
          x = conv3(p, f'enc/s{s}b{b}c2', x)
          x = x + skip
      x = x.reshape(x.shape[0], -1)
      x = act(cfg.enc_act, norm(p, 'enc/outn1', x, cfg.enc_norm))
      x = linear(p, 'enc/outl', x)
      x = act(cfg.enc_act, norm(p, 'enc/outn2', x, cfg.enc_norm))
      outs.append(x)
    x = torch.cat(outs, -1)
    return x.reshape(*bshape, -1)

  # -- nets.py:657-669 -------------------------------------------------------------
  def gru_step(self, p, carry, inp, reset):
    cfg = self.cfg
    carry = torch.where(reset[:, None], torch.zeros_like(carry), carry)
    x = torch.cat([carry, inp], -1)
    x = norm(p, 'rnn/norm', x, cfg.rnn_norm)
    x = linear(p, 'rnn/linear', x)
    res, cand, update = torch.chunk(x, 3, -1)
    cand = torch.tanh(torch.sigmoid(res) * cand)
    update = torch.sigmoid(update - 1.0)
    carry = update * cand + (1 - update) * carry
    return carry, carry

  def head_mlp(self, p, name, x, layers):
    for i in range(layers):
      x = linear(p, f'{name}/mlp/linear{i}', x)
      x = act(self.cfg.head_act, norm(p, f'{name}/mlp/norm{i}', x, self.cfg.head_norm))
    return x

  def policy_outputs(self, p, feat):                             # heads.py:100-112,146-155
    cfg = self.cfg
    h = self.head_mlp(p, 'policy', feat, cfg.pol_layers)
    outs = {}
    for key, space in self.act_space.items():
      kind, shape, classes = space_kind(space)
      if kind == 'disc':
        y = linear(p, f'policy/head/{key}/logits', h)
        outs[key] = y.reshape(*y.shape[:-1], *shape, classes)
      else:
        mean = linear(p, f'policy/head/{key}/mean', h)
        std = linear(p, f'policy/head/{key}/stddev', h)
        std = (cfg.maxstd - cfg.minstd) * torch.sigmoid(std + 2.0) + cfg.minstd
        outs[key] = (torch.tanh(mean).reshape(*mean.shape[:-1], *shape), std.reshape(*std.shape[:-1], *shape))
    return outs

  def logp_entropy(self, outs, acts):
    """{key: logp}, {key: entropy}, event dims summed (outs.Agg, heads.py:89-90)."""
    logps, ents = {}, {}
    for key, space in self.act_space.items():
      kind, shape, classes = space_kind(space)
      if kind == 'disc':
        la = torch.log_softmax(outs[key], -1)
        lp = (la * F.one_hot(acts[key].long(), classes)).sum(-1)
        en = -(torch.softmax(outs[key], -1) * la).sum(-1)
      else:
        mean, std = outs[key]
        lp = -0.5 * ((acts[key].to(f32) - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
        en = 0.5 * torch.log(2 * math.pi * std * std) + 0.5
      for _ in shape:
        lp, en = lp.sum(-1), en.sum(-1)
      logps[key], ents[key] = lp, en
    return logps, ents

  def sample(self, outs, noise):
    acts = {}
    for key, space in self.act_space.items():
      kind, shape, classes = space_kind(space)
      if kind == 'disc':
        acts[key] = torch.argmax(outs[key] + noise[key], -1).to(torch.int32)
      else:
        mean, std = outs[key]
        acts[key] = noise[key] * std + mean
    return acts

  def value_pred(self, p, feat):
    h = self.head_mlp(p, 'value', feat, self.cfg.val_layers)
    return linear(p, 'value/head/pred', h).squeeze(-1)

  # -- ppo/agent.py:162-183 ----------------------------------------------------------
  def model(self, p, memory, obs, prevact, value=True, single=False):
    cfg = self.cfg
    bdims = 1 if single else 2
    bshape = tuple(obs['is_first'].shape[:bdims])
    embed = self.encoder(p, {k: obs[k] for k in list(self.vec) + list(self.img)}, bdims)
    if cfg.recurrent:
      if cfg.rnnact:
        first = obs['is_first']
        masked = {k: torch.where(first.reshape(*bshape, *([1] * (v.ndim - bdims))), torch.zeros_like(v), v)
                  for k, v in prevact.items()}
        clip = lambda x: x / torch.clamp(torch.abs(x), min=1.0).detach()
        inputs = torch.cat([embed, dict_embed(p, 'actemb', self.act_space, masked, bshape,
                                              cfg.actemb_units, clip)], -1)
      else:
        inputs = embed
      if single:
        memory, feat = self.gru_step(p, memory, inputs, obs['is_first'])
      else:
        feats = []
        for t in range(bshape[1]):
          memory, out = self.gru_step(p, memory, inputs[:, t], obs['is_first'][:, t])
          feats.append(out)
        feat = torch.stack(feats, 1)
    else:
      feat = embed
    policy = self.policy_outputs(p, feat)
    val = self.value_pred(p, feat) if value else None
    return memory, feat, policy, val

  def initial(self, batch):
    return torch.zeros(batch, self.cfg.rnn_units) if self.cfg.recurrent else ()

  def policy(self, carry, obs, noise):                            # agent.py:71-81
    memory, prevact = carry
    with torch.no_grad():
      memory, feat, pol, _ = self.model(self.p, memory, obs, prevact, value=False, single=True)
      acts = self.sample(pol, noise)
      logps, _ = self.logp_entropy(pol, acts)
    out = {f'logp/{k}': v for k, v in logps.items()}
    if self.cfg.recurrent:
      out['memory'] = memory
    return (memory, acts), acts, out

  # -- utils.py:39-91 ('meanstd', debias) -----------------------------------------------
  def norm_stats(self, name):
    st, cfg = self.norms[name], self.cfg
    corr = 1.0 / max(cfg.norm_rate, st['corr'])
    mean = st['mean'] * corr
    std = math.sqrt(max(st['sqrs'] * corr - mean ** 2, 0.0))
    return mean, max(cfg.norm_limit, std)

  def norm_update(self, name, x):
    st, r = self.norms[name], self.cfg.norm_rate
    x = x.detach().to(f32)
    st['mean'] = (1 - r) * st['mean'] + r * float(x.mean())
    st['sqrs'] = (1 - r) * st['sqrs'] + r * float((x * x).mean())
    st['corr'] = (1 - r) * st['corr'] + r * 1.0

  # -- ppo/agent.py:186-235 ---------------------------------------------------------------
  def ppo_loss(self, data, policy, value, update=True):
    cfg = self.cfg
    acts = {k: data[k] for k in self.act_space}
    logps, ents = self.logp_entropy(policy, acts)
    logpi = sum(logps.values())
    logdata = sum(data['logp/' + k] for k in self.act_space)
    rew, last, term = data['reward'], data['is_last'], data['is_terminal']
    mask = (~last & ~term).to(f32)
    ratio = torch.exp(logpi - logdata.detach())
    voffset, vscale = self.norm_stats('valnorm')
    val = value * vscale + voffset
    live = (~term).to(f32)[:, 1:] * (1 - 1 / cfg.hor)
    cont = (~last & ~term).to(f32)[:, 1:] * cfg.lam
    delta = rew[:, 1:] + live * val[:, 1:] - val[:, :-1]
    advs = [torch.zeros_like(delta[:, 0])]
    for t in reversed(range(delta.shape[1])):
      advs.append(delta[:, t] + live[:, t] * cont[:, t] * advs[-1])
    adv = torch.stack(list(reversed(advs))[:-1], 1)
    tar = adv + val[:, :-1]
    if update:
      self.norm_update('valnorm', tar)
    voffset, vscale = self.norm_stats('valnorm')
    tarnormed = (tar - voffset) / vscale
    if cfg.tarclip:
      tarnormed = torch.clamp(tarnormed, -cfg.tarclip, cfg.tarclip)
    padded = torch.cat([tarnormed, 0 * tarnormed[:, :1]], 1)
    losses = {'value': (value - padded.detach()) ** 2 * mask}
    if update:
      self.norm_update('advnorm', adv)
    aoffset, ascale = self.norm_stats('advnorm')
    advnormed = (adv - aoffset) / ascale
    reinforce = ratio[:, :-1] * advnormed.detach()
    maxent = cfg.actent * sum(ents.values())[:, :-1]
    upper = (ratio[:, :-1] < 1 + cfg.trclip) | (advnormed < 0)
    lower = (ratio[:, :-1] > 1 - cfg.trclip) | (advnormed > 0)
    tr = (upper & lower).to(f32)
    losses['policy'] = -(reinforce + maxent) * mask[:, :-1] * tr
    metrics = {f'ent/{k}': v.mean() for k, v in ents.items()}
    for k, space in self.act_space.items():                       # agent.py:222-224: heads.py:108-109,152-153
      kind, shape, classes = space_kind(space)                    # set minent / maxent; outs.Agg (shaped keys) has none
      if not shape:
        if kind == 'disc':
          lo, hi = 0.0, math.log(classes)
        else:
          lo, hi = [0.5 * math.log(2 * math.pi * s * s) + 0.5 for s in (cfg.minstd, cfg.maxstd)]
        metrics[f'rand/{k}'] = (ents[k].mean() - lo) / (hi - lo)
    metrics.update(rew=rew.mean(), val=val.mean(), tar=tar.mean(), adv=adv.mean(),
                   advmag=adv.abs().mean(), ratio=ratio.mean(), clipfrac=(1 - tr).mean(),
                   td=(value[:, :-1] - tarnormed).abs().mean())
    return losses, metrics

  def loss(self, p, memory, data, prevact, update=True):          # agent.py:108-118
    memory, feat, policy, value = self.model(p, memory, data, prevact)
    losses, metrics = self.ppo_loss(data, policy, value, update)
    for k, v in losses.items():
      metrics[f'{k}_loss'] = v.mean()
      metrics[f'{k}_loss_std'] = v.std(unbiased=False)          # jnp.std: population (ddof = 0)
    total = sum(v.mean() * self.cfg.scales[k] for k, v in losses.items())
    return total, (memory, metrics, losses)

  def train(self, carry, data):                                   # agent.py:83-103
    cfg = self.cfg
    memory, prevact = carry
    if cfg.replay_context:
      K = cfg.replay_context
      prevact = {k: data[k][:, K - 1:-1] for k in self.act_space}
      data = {k: v[:, K:] for k, v in data.items()}
      if cfg.recurrent:
        data = dict(data)
        memory = data.pop('memory').to(f32)[:, K - 1]            # sic: index K-1 of the SLICED rows
    else:
      prepend = lambda x, y: torch.cat([x[:, None], y[:, :-1]], 1)
      prevact = {k: prepend(prevact[k], data[k]) for k in self.act_space}
    p = {k: v.clone().requires_grad_(True) for k, v in self.p.items()}
    total, (memory, metrics, losses) = self.loss(p, memory, data, prevact)
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    metrics['loss'] = total.detach()
    metrics.update(self.apply_updates(grads))
    prevact = {k: data[k][:, -1] for k in self.act_space}
    memory = memory.detach() if cfg.recurrent else memory
    return (memory, prevact), {}, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in metrics.items()}, grads, losses

  def learning_rate(self, count):                                 # optax.linear_schedule(0, lr, warmup)
    return self.cfg.lr * min(count / self.cfg.warmup, 1.0) if self.cfg.warmup else self.cfg.lr

  def apply_updates(self, grads):                                 # agent.py:120-131 + opt.py:60-64
    cfg, st = self.cfg, self.opt
    gnorm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    factor = 1.0 if gnorm < cfg.clip else cfg.clip / gnorm       # optax.clip_by_global_norm
    count = st['count'] + 1
    lr = self.learning_rate(st['count'])                          # the schedule's count lags by one
    b1, b2 = 0.9, 0.999
    pattern = re.compile(cfg.wdregex)
    for k, g in grads.items():
      g = g * factor
      st['mu'][k] = b1 * st['mu'][k] + (1 - b1) * g
      st['nu'][k] = b2 * st['nu'][k] + (1 - b2) * g * g
      mhat = st['mu'][k] / (1 - b1 ** count)
      vhat = st['nu'][k] / (1 - b2 ** count)
      upd = mhat / (torch.sqrt(vhat) + cfg.eps)
      if cfg.wd and pattern.search('/' + k):
        upd = upd + cfg.wd * self.p[k]
      self.p[k] = self.p[k] - lr * upd
    st['count'] = count
    return {'opt/grad_norm': gnorm, 'opt/updates': count}


# ------------------------------------------------------------------ parameters
def param_shapes(cfg, obs_space, act_space):
  """name -> (shape, fan-in shape for the initialiser, outscale).  Names follow the module
  tree of ppo/agent.py:134-159 (enc / actemb / rnn / policy / value)."""
  vec, img = enc_spaces(obs_space)
  shapes = {}

  def lin(name, i, o, outscale=1.0):
    shapes[f'{name}/kernel'] = ((i, o), 'trunc', outscale)
    shapes[f'{name}/bias'] = ((o,), 'zeros', 1.0)

  def nrm(name, n, impl):
    if impl != 'none':
      shapes[f'{name}/scale'] = ((n,), 'ones', 1.0)
      shapes[f'{name}/shift'] = ((n,), 'zeros', 1.0)

  def embed(name, spaces, units):
    shapes[f'{name}/init'] = ((units,), 'trunc_out', 1.0)
    for key in sorted(spaces):
      kind, shape, classes = space_kind(spaces[key])
      width = int(np.prod(shape, dtype=np.int64)) * (classes if kind == 'disc' else 1)
      lin(f'{name}/{key}', width, units)

  width = 0
  if vec:
    embed('enc/emb', vec, cfg.enc_units)
    for i in range(cfg.enc_layers - 1):
      lin(f'enc/mlp/linear{i}', cfg.enc_units, cfg.enc_units)
      nrm(f'enc/mlp/norm{i}', cfg.enc_units, cfg.enc_norm)
    width += cfg.enc_units
  if img:
    first = next(iter(img.values()))
    h, w = first.shape[:2]
    chans = sum(v.shape[-1] for v in img.values())
    for s, mult in enumerate(cfg.mults):
      d = cfg.depth * mult
      shapes[f'enc/s{s}in/kernel'] = ((3, 3, chans, d), 'trunc', 1.0)
      shapes[f'enc/s{s}in/bias'] = ((d,), 'zeros', 1.0)
      h, w = -(-h // 2), -(-w // 2)
      for b in range(cfg.blocks):
        for c in ('c1', 'c2'):
          nrm(f'enc/s{s}b{b}n{c[1]}', d, cfg.enc_norm)
          shapes[f'enc/s{s}b{b}{c}/kernel'] = ((3, 3, d, d), 'trunc', 1.0)
          shapes[f'enc/s{s}b{b}{c}/bias'] = ((d,), 'zeros', 1.0)
      chans = d
    nrm('enc/outn1', h * w * chans, cfg.enc_norm)
    lin('enc/outl', h * w * chans, cfg.outmult * cfg.depth)
    nrm('enc/outn2', cfg.outmult * cfg.depth, cfg.enc_norm)
    width += cfg.outmult * cfg.depth
  feat = width
  if cfg.recurrent:
    inp = width
    if cfg.rnnact:
      embed('actemb', act_space, cfg.actemb_units)
      inp += cfg.actemb_units
    nrm('rnn/norm', cfg.rnn_units + inp, cfg.rnn_norm)
    lin('rnn/linear', cfg.rnn_units + inp, 3 * cfg.rnn_units)
    feat = cfg.rnn_units

  def head_mlp(name, layers, units):
    n = feat
    for i in range(layers):
      lin(f'{name}/mlp/linear{i}', n, units)
      nrm(f'{name}/mlp/norm{i}', units, cfg.head_norm)
      n = units
    return n

  n = head_mlp('policy', cfg.pol_layers, cfg.pol_units)
  for key, space in act_space.items():
    kind, shape, classes = space_kind(space)
    size = int(np.prod(shape, dtype=np.int64))
    if kind == 'disc':
      lin(f'policy/head/{key}/logits', n, size * classes, cfg.pol_outscale)
    else:
      lin(f'policy/head/{key}/mean', n, size, cfg.pol_outscale)
      lin(f'policy/head/{key}/stddev', n, size, cfg.pol_outscale)
  n = head_mlp('value', cfg.val_layers, cfg.val_units)
  lin('value/head/pred', n, 1, cfg.val_outscale)
  return shapes


def init_params(cfg, obs_space, act_space, seed=0, outscale_override=None):
  """Truncated normal x 1.1368 / sqrt(fan) (nets.py:168-170; fan-in for kernels, fan-out for the
  embedding's init vector), zeros / ones otherwise.  Not the reference's random stream."""
  g = torch.Generator().manual_seed(seed)
  out = {}
  for name, (shape, kind, outscale) in param_shapes(cfg, obs_space, act_space).items():
    if outscale_override is not None and outscale != 1.0:
      outscale = outscale_override
    if kind == 'zeros':
      out[name] = torch.zeros(shape)
    elif kind == 'ones':
      out[name] = torch.ones(shape)
    else:
      if kind == 'trunc_out':
        fan = shape[0]
      elif len(shape) == 2:
        fan = shape[0]
      else:
        fan = shape[-2] * int(np.prod(shape[:-2]))
      x = torch.empty(shape)
      torch.nn.init.trunc_normal_(x, 0.0, 1.0, -2.0, 2.0, generator=g)
      out[name] = x * (1.1368 * math.sqrt(1 / fan)) * outscale
  return out


def make_noise(act_space, lead, seed=0):
  """Injected sampling noise per action key: Gumbel (categorical) / standard normal."""
  g = torch.Generator().manual_seed(seed)
  noise = {}
  for key, space in act_space.items():
    kind, shape, classes = space_kind(space)
    if kind == 'disc':
      u = torch.rand(*lead, *shape, classes, generator=g).clamp_(1e-6, 1 - 1e-6)
      noise[key] = -torch.log(-torch.log(u))
    else:
      noise[key] = torch.randn(*lead, *shape, generator=g)
  return noise
