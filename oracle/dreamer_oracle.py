"""fp32 torch-CPU restatement of the reference's dreamerv3 math (TEST INFRASTRUCTURE).

PARITY UNPINNED: the reference holds no test or golden tensor for dreamerv3
 and JAX/ninjax/optax cannot be installed in this image, so
this file cannot be checked against the running original.  It restates, line by
line, the code cited below; sampling noise is INJECTED (Gumbel tensors) so that
the product kernels and this file can be compared on identical draws
(jax.random.categorical(key, l) == argmax(l + gumbel(key))).

Restated here (reference file:line):
  linear / block_linear / conv / rms      embodied/jax/nets.py:230-251, 254-281, 284-323, 361-399
  onehot_dist (unimix, straight-through)  embodied/jax/outs.py:208-270
  twohot (bins, symmetric pred, loss)     embodied/jax/heads.py:132-144, outs.py:273-330
  binary                                   embodied/jax/outs.py:189-205
  Encoder / Decoder                        dreamerv3/rssm.py:210-250, 288-359
  RSSM core / observe / prior / loss       dreamerv3/rssm.py:61-92, 120-176
  RSSM imagine                             dreamerv3/rssm.py:94-118
  Agent.loss / policy / replay context     dreamerv3/agent.py:115-135, 156-245, 312-340
  imag_loss / repl_loss / lambda_return    dreamerv3/agent.py:382-490
  Normalize('perc')                        embodied/jax/utils.py:16-91
  optimizer chain                          embodied/jax/opt.py:109-164, dreamerv3/agent.py:342-379
  SlowModel.update                         embodied/jax/utils.py:113-119

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

f32 = torch.float32


class Config(dict):
  __getattr__ = dict.__getitem__


def default_config(**over):
  """dreamerv3/configs.yaml:81-117 (size200m) as a flat dict."""
  cfg = Config(
      deter=8192, hidden=1024, stoch=32, classes=64, blocks=8, unimix=0.01,
      free_nats=1.0, imglayers=2, obslayers=1, dynlayers=1,
      depth=64, mults=(2, 3, 4, 4), kernel=5, units=1024, bspace=8,
      image=(64, 64, 3), actions=5, bins=255,
      rew_layers=1, con_layers=1, pol_layers=3, val_layers=3,
      imag_length=15, horizon=333, contdisc=True, lam=0.95, actent=3e-4,
      slowreg=1.0, slowrate=0.02, replay_context=1,
      retnorm_rate=0.01, retnorm_limit=1.0, perclo=5.0, perchi=95.0,
      scales=dict(image=1.0, rew=1.0, con=1.0, dyn=1.0, rep=0.1,
                  policy=1.0, value=1.0, repval=0.3),
      lr=4e-5, agc=0.3, eps=1e-20, beta1=0.9, beta2=0.999, warmup=1000,
      pmin=1e-3)
  cfg.update(over)
  return cfg


# ------------------------------------------------------------------ primitives
def silu(x):
  return x * torch.sigmoid(x)


def rms(x, scale, eps=1e-4):                                   # nets.py:374-383
  mean2 = (x * x).mean(-1, keepdim=True)
  return x * (torch.rsqrt(mean2 + eps) * scale)


def linear(p, name, x):                                         # nets.py:239-247
  return x @ p[f'{name}/kernel'] + p[f'{name}/bias']


def block_linear(p, name, x, g):                                # nets.py:267-278
  k = p[f'{name}/kernel']                                       # (g, in/g, out/g)
  x = x.reshape(*x.shape[:-1], g, x.shape[-1] // g)
  x = torch.einsum('...ki,kio->...ko', x, k)
  return x.reshape(*x.shape[:-2], -1) + p[f'{name}/bias']


def conv(p, name, x):                                           # nets.py:298-323 (NHWC, HWIO, SAME)
  k = p[f'{name}/kernel'].permute(3, 2, 0, 1)
  y = F.conv2d(x.permute(0, 3, 1, 2), k, padding=k.shape[-1] // 2)
  return y.permute(0, 2, 3, 1) + p[f'{name}/bias']


def layer(p, name, x):
  """Linear -> rms -> silu with the '<name>norm' scale (rssm.py:141-146 pattern)."""
  return silu(rms(linear(p, name, x), p[f'{name}norm/scale']))


def mlp(p, name, x, layers):                                    # nets.py:580-587
  for i in range(layers):
    x = linear(p, f'{name}/linear{i}', x)
    x = silu(rms(x, p[f'{name}/norm{i}/scale']))
  return x


def unimix_logits(logits, unimix):                              # outs.py:210-216
  probs = torch.softmax(logits, -1)
  probs = (1 - unimix) * probs + unimix / probs.shape[-1]
  return torch.log(probs)


def onehot_sample(logits, unimix, gumbel):                      # outs.py:252-270
  lg = unimix_logits(logits, unimix)
  index = torch.argmax(lg + gumbel, -1)
  value = F.one_hot(index, lg.shape[-1]).to(f32)
  probs = torch.softmax(lg, -1)
  return value + (probs - probs.detach())


def cat_kl(a, b):                                               # outs.py:236-240 (+ Agg sum :73-76)
  la, lb = torch.log_softmax(a, -1), torch.log_softmax(b, -1)
  return (torch.softmax(a, -1) * (la - lb)).sum(-1).sum(-1)


def cat_entropy(a):                                             # outs.py:230-234
  la = torch.log_softmax(a, -1)
  return -(torch.softmax(a, -1) * la).sum(-1)


def symexp(x):
  return torch.sign(x) * torch.expm1(torch.abs(x))


def symlog(x):                                                  # nets.py:59-60
  return torch.sign(x) * torch.log1p(torch.abs(x))


def dict_concat(specs, values, squish=None):                    # nets.py:467-500 (fdims = 1)
  """specs: sorted list of (key, kind 'disc'|'cont', shape, classes).  Integers -> one-hot,
  floats -> squish; -1 / -inf entries are `unavailable` and masked to zero (nets.py:76-94)."""
  ys = []
  for key, kind, shape, classes in specs:
    x = values[key]
    lead = x.shape[:x.dim() - len(shape)]
    if kind == 'disc':
      x = x.long()
      m = x != -1
      y = F.one_hot(torch.where(m, x, torch.zeros_like(x)), classes).to(f32) * m[..., None]
    else:
      x = x.to(f32)
      m = x != float('-inf')
      x = torch.where(m, x, torch.zeros_like(x))
      y = (squish(x) if squish else x) * m
    ys.append(y.reshape(*lead, -1))
  return torch.cat(ys, -1)


def space_specs(cfg):
  """(imgkeys [(key, channels)], vecspec, actspec) of a config; the defaults describe config 2:
  one `image` key and one scalar discrete `action` with cfg.actions classes."""
  img = cfg.get('imgkeys')
  if img is None:
    img = [('image', cfg.image[2])] if cfg.get('image') else []
  vec = cfg.get('vecspec') or []
  act = cfg.get('actspec') or [('action', 'disc', (), cfg.actions)]
  return img, vec, act


def spec_width(spec):
  n = int(np.prod(spec[2], dtype=np.int64))
  return n * (spec[3] if spec[1] == 'disc' else 1)


def twohot_bins(n=255):                                         # heads.py:132-144
  assert n % 2 == 1
  half = symexp(torch.linspace(-20, 0, (n - 1) // 2 + 1, dtype=f32))
  return torch.cat([half, -half[:-1].flip(0)], 0)


def twohot_pred(logits, bins):                                  # outs.py:285-302
  probs = torch.softmax(logits, -1)
  n = logits.shape[-1]
  m = (n - 1) // 2
  p1, p2, p3 = probs[..., :m], probs[..., m: m + 1], probs[..., m + 1:]
  b1, b2, b3 = bins[:m], bins[m: m + 1], bins[m + 1:]
  return (p2 * b2).sum(-1) + ((p1 * b1).flip(-1) + (p3 * b3)).sum(-1)


def twohot_loss(logits, bins, target):                          # outs.py:311-330
  target = target.detach()
  n = len(bins)
  below = (bins <= target[..., None]).to(torch.int32).sum(-1) - 1
  above = n - (bins > target[..., None]).to(torch.int32).sum(-1)
  below = below.clamp(0, n - 1).long()
  above = above.clamp(0, n - 1).long()
  equal = below == above
  one = torch.ones_like(target)
  to_below = torch.where(equal, one, (bins[below] - target).abs())
  to_above = torch.where(equal, one, (bins[above] - target).abs())
  total = to_below + to_above
  wb, wa = to_above / total, to_below / total
  tgt = F.one_hot(below, n) * wb[..., None] + F.one_hot(above, n) * wa[..., None]
  logp = logits - torch.logsumexp(logits, -1, keepdim=True)
  return -(tgt * logp).sum(-1)


def binary_logp(logit, event):                                  # outs.py:197-201
  return event * F.logsigmoid(logit) + (1 - event) * F.logsigmoid(-logit)


def lambda_return(last, term, rew, val, boot, disc, lam):       # agent.py:482-490
  rets = [boot[:, -1]]
  live = (1 - term.to(f32))[:, 1:] * disc
  cont = (1 - last.to(f32))[:, 1:] * lam
  interm = rew[:, 1:] + (1 - cont) * live * boot[:, 1:]
  for t in reversed(range(live.shape[1])):
    rets.append(interm[:, t] + live[:, t] * cont[:, t] * rets[-1])
  return torch.stack(list(reversed(rets))[:-1], 1)


# --------------------------------------------------------------------- model
class Dreamer:
  """`p`: dict name -> float32 torch CPU tensor in the reference's parameter
  naming (ninjax paths: 'dyn/dynin0/kernel', 'enc/cnn0norm/scale', ...)."""

  def __init__(self, cfg, params, slow=None, state=None):
    self.cfg = cfg
    self.p = params
    self.slow = slow if slow is not None else {
        k.replace('val/', 'slowval/', 1): v.detach().clone()
        for k, v in params.items() if k.startswith('val/')}
    self.bins = twohot_bins(cfg.bins)
    self.state = state if state is not None else dict(
        ret_lo=torch.zeros((), dtype=f32), ret_hi=torch.zeros((), dtype=f32),
        step=0, nu={k: torch.zeros_like(v) for k, v in params.items()},
        mu={k: torch.zeros_like(v) for k, v in params.items()})

  # -- encoder (rssm.py:210-250) ------------------------------------------------
  def encoder(self, obs):
    """obs: dict of the observation keys (a bare uint8 tensor = the single image key)."""
    p, cfg = self.p, self.cfg
    img, vec, _ = space_specs(cfg)
    if torch.is_tensor(obs):
      obs = {img[0][0]: obs}
    outs = []
    if vec:                                                     # :215-224 DictConcat(symlog) -> MLP
      x = dict_concat(vec, obs, symlog)
      lead = x.shape[:-1]
      x = x.reshape(-1, x.shape[-1])
      for i in range(cfg.get('enc_layers', 3)):
        x = layer(p, f'enc/mlp{i}', x)
      outs.append(x)
    if img:                                                     # :226-246
      x = torch.cat([obs[k] for k, _ in img], -1).to(f32) / 255 - 0.5
      lead = x.shape[:-3]
      x = x.reshape(-1, *x.shape[-3:])
      for i in range(len(cfg.mults)):
        x = conv(p, f'enc/cnn{i}', x)
        n, h, w, c = x.shape
        x = x.reshape(n, h // 2, 2, w // 2, 2, c).amax((2, 4))
        x = silu(rms(x, p[f'enc/cnn{i}norm/scale']))
      outs.append(x.reshape(x.shape[0], -1))
    return torch.cat(outs, -1).reshape(*lead, -1)

  # -- rssm ------------------------------------------------------------------
  def action_embed(self, action, reset=None):
    """rssm.py:76-79: mask the action dict where reset, DictConcat (nets.py:467-500), mask again.
    A bare tensor is the single action key."""
    _, _, act = space_specs(self.cfg)
    if torch.is_tensor(action):
      action = {act[0][0]: action}
    if reset is not None:
      action = {k: torch.where(reset.reshape(reset.shape + (1,) * (v.dim() - reset.dim())),
                               torch.zeros_like(v), v) for k, v in action.items()}
    a = dict_concat(act, action)
    if reset is not None:
      a = a * (~reset)[..., None]
    return a

  def core(self, deter, stoch, action):                         # rssm.py:135-159
    p, cfg, g = self.p, self.cfg, self.cfg.blocks
    stoch = stoch.reshape(stoch.shape[0], -1)
    action = action / torch.clamp(action.abs(), min=1).detach()
    x0 = layer(p, 'dyn/dynin0', deter)
    x1 = layer(p, 'dyn/dynin1', stoch)
    x2 = layer(p, 'dyn/dynin2', action)
    x = torch.cat([x0, x1, x2], -1)[:, None, :].expand(-1, g, -1)
    x = torch.cat([deter.reshape(len(deter), g, -1), x], -1).reshape(len(deter), -1)
    x = block_linear(p, 'dyn/dynhid0', x, g)
    x = silu(rms(x, p['dyn/dynhid0norm/scale']))
    x = block_linear(p, 'dyn/dyngru', x, g)
    gates = x.reshape(len(x), g, -1).chunk(3, -1)
    reset, cand, update = [y.reshape(len(x), -1) for y in gates]
    reset = torch.sigmoid(reset)
    cand = torch.tanh(reset * cand)
    update = torch.sigmoid(update - 1)
    return update * cand + (1 - update) * deter

  def prior(self, deter):                                       # rssm.py:161-171
    x = deter
    for i in range(self.cfg.imglayers):
      x = layer(self.p, f'dyn/prior{i}', x)
    x = linear(self.p, 'dyn/priorlogit', x)
    return x.reshape(*x.shape[:-1], self.cfg.stoch, self.cfg.classes)

  def observe_step(self, carry, tokens, action, reset, gumbel):  # rssm.py:75-92
    cfg = self.cfg
    keep = (~reset).to(f32)
    deter = carry['deter'] * keep[:, None]
    stoch = carry['stoch'] * keep[:, None, None]
    act = self.action_embed(action, reset)
    deter = self.core(deter, stoch, act)
    x = torch.cat([deter, tokens], -1)
    x = layer(self.p, 'dyn/obs0', x)
    logit = linear(self.p, 'dyn/obslogit', x).reshape(len(x), cfg.stoch, cfg.classes)
    stoch = onehot_sample(logit, cfg.unimix, gumbel)
    return dict(deter=deter, stoch=stoch), dict(deter=deter, stoch=stoch, logit=logit)

  def observe(self, carry, tokens, action, reset, gumbel):       # rssm.py:61-73
    feats = []
    for t in range(tokens.shape[1]):
      act_t = {k: v[:, t] for k, v in action.items()} if isinstance(action, dict) else action[:, t]
      carry, feat = self.observe_step(carry, tokens[:, t], act_t, reset[:, t], gumbel[:, t])
      feats.append(feat)
    feat = {k: torch.stack([f[k] for f in feats], 1) for k in feats[0]}
    return carry, feat

  def rssm_loss(self, feat):                                    # rssm.py:120-133
    cfg = self.cfg
    prior = unimix_logits(self.prior(feat['deter']), cfg.unimix)
    post = unimix_logits(feat['logit'], cfg.unimix)
    dyn = cat_kl(post.detach(), prior)
    rep = cat_kl(post, prior.detach())
    dyn = torch.clamp(dyn, min=cfg.free_nats)
    rep = torch.clamp(rep, min=cfg.free_nats)
    mets = dict(dyn_ent=cat_entropy(prior).sum(-1).mean(),
                rep_ent=cat_entropy(post).sum(-1).mean())
    return dict(dyn=dyn, rep=rep), mets

  def imagine_step(self, carry, action_onehot, gumbel):         # rssm.py:95-104
    deter = self.core(carry['deter'], carry['stoch'], action_onehot)
    logit = self.prior(deter)
    stoch = onehot_sample(logit, self.cfg.unimix, gumbel)
    return dict(deter=deter, stoch=stoch)

  # -- decoder (rssm.py:288-359) --------------------------------------------------
  def decoder(self, deter, stoch):
    """-> {'image': sigmoid reconstruction of the concatenated image keys, <vector key>: head
    output (`pred` in symlog space | categorical logits)}."""
    p, cfg = self.p, self.cfg
    img, vec, _ = space_specs(cfg)
    lead = deter.shape[:-1]
    recons = {}
    if vec:                                                     # :323-334
      x = torch.cat([stoch.reshape(*lead, -1), deter], -1)      # inp = [stoch, deter] (:319-321)
      x = mlp(p, 'dec/mlp', x.reshape(-1, x.shape[-1]), cfg.get('dec_layers', 3))
      for key, kind, shape, classes in vec:
        y = linear(p, f'dec/vec/{key}/' + ('logits' if kind == 'disc' else 'pred'), x)
        recons[key] = y.reshape(*lead, *shape, *((classes,) if kind == 'disc' else ()))
    if not img:
      return recons
    x0 = deter.reshape(-1, deter.shape[-1])
    x1 = stoch.reshape(x0.shape[0], -1)
    minres = cfg.image[0] // 2 ** len(cfg.mults)
    depths = [cfg.depth * m for m in cfg.mults]
    g, c = cfg.bspace, depths[-1] // cfg.bspace
    x0 = block_linear(p, 'dec/sp0', x0, g)
    x0 = x0.reshape(-1, g, minres, minres, c).permute(0, 2, 3, 1, 4).reshape(
        -1, minres, minres, g * c)                              # '(g h w c) -> h w (g c)'
    x1 = layer(p, 'dec/sp1', x1)
    x1 = linear(p, 'dec/sp2', x1).reshape(-1, minres, minres, depths[-1])
    x = silu(rms(x0 + x1, p['dec/spnorm/scale']))
    for i in reversed(range(len(depths) - 1)):
      x = x.repeat_interleave(2, 2).repeat_interleave(2, 1)
      x = conv(p, f'dec/conv{i}', x)
      x = silu(rms(x, p[f'dec/conv{i}norm/scale']))
    x = x.repeat_interleave(2, 2).repeat_interleave(2, 1)
    x = torch.sigmoid(conv(p, 'dec/imgout', x))
    recons['image'] = x.reshape(*lead, *x.shape[1:])
    return recons

  def recon_losses(self, recons, obs):                          # agent.py:176-180
    img, vec, _ = space_specs(self.cfg)
    out, c0 = {}, 0
    for key, ch in img:                                         # rssm.py:354-357: MSE, Agg(3, sum)
      target = obs[key].to(f32) / 255
      out[key] = ((recons['image'][..., c0: c0 + ch] - target) ** 2).sum((-3, -2, -1))
      c0 += ch
    for key, kind, shape, classes in vec:
      y = recons[key]
      if kind == 'disc':                                        # categorical: -logp (outs.py:23-24)
        loss = -torch.log_softmax(y, -1).gather(-1, obs[key].long()[..., None]).squeeze(-1)
      else:                                                     # symlog_mse (heads.py:126-129)
        loss = (y - symlog(obs[key].to(f32))) ** 2
      for _ in shape:                                           # Agg over the key's own axes (heads.py:87-88)
        loss = loss.sum(-1)
      out[key] = loss
    return out

  # -- heads (heads.py:16-41) ----------------------------------------------------
  def feat2tensor(self, deter, stoch):                          # agent.py:51-53
    return torch.cat([deter, stoch.reshape(*stoch.shape[:-2], -1)], -1)

  def head(self, name, x, layers, out, p=None):
    p = self.p if p is None else p
    x = mlp(p, f'{name}/mlp', x, layers)
    return linear(p, f'{name}/head/{out}', x)

  def rew_logits(self, x):
    return self.head('rew', x, self.cfg.rew_layers, 'logits')

  def con_logit(self, x):
    return self.head('con', x, self.cfg.con_layers, 'logit').squeeze(-1)

  def pol_outputs(self, x):                                     # heads.py:103-112, 146-155
    """{key: logits} (categorical) / {key: (mean, std)} (bounded_normal) per action key."""
    cfg = self.cfg
    _, _, act = space_specs(cfg)
    h = mlp(self.p, 'pol/mlp', x, cfg.pol_layers)
    outs = {}
    for key, kind, shape, classes in act:
      if kind == 'disc':
        y = linear(self.p, f'pol/head/{key}/logits', h)
        outs[key] = y.reshape(*y.shape[:-1], *shape, classes)
      else:
        mean = linear(self.p, f'pol/head/{key}/mean', h)
        std = linear(self.p, f'pol/head/{key}/stddev', h)
        lo, hi = cfg.get('minstd', 0.1), cfg.get('maxstd', 1.0)
        std = (hi - lo) * torch.sigmoid(std + 2.0) + lo
        outs[key] = (torch.tanh(mean).reshape(*mean.shape[:-1], *shape),
                     std.reshape(*std.shape[:-1], *shape))
    return outs

  def pol_sample(self, outs, noise):
    """noise: {key: Gumbel (categorical) | standard normal (Normal.sample, outs.py:156-158)};
    a bare tensor is the single action key."""
    _, _, act = space_specs(self.cfg)
    if torch.is_tensor(noise):
      noise = {act[0][0]: noise}
    res = {}
    for key, kind, shape, classes in act:
      if kind == 'disc':
        res[key] = torch.argmax(outs[key] + noise[key], -1).to(torch.int32)
      else:
        mean, std = outs[key]
        res[key] = noise[key] * std + mean
    return res

  def pol_logp_entropy(self, outs, acts):                       # agent.py:407-408
    _, _, act = space_specs(self.cfg)
    logp, ent = 0.0, 0.0
    for key, kind, shape, classes in act:
      if kind == 'disc':
        la = torch.log_softmax(outs[key], -1)
        lp = la.gather(-1, acts[key].long()[..., None]).squeeze(-1)
        en = -(torch.softmax(outs[key], -1) * la).sum(-1)
      else:
        mean, std = outs[key]
        lp = torch.distributions.Normal(mean, std).log_prob(acts[key].to(f32))   # outs.py:160-163
        en = 0.5 * torch.log(2 * math.pi * std ** 2) + 0.5                        # outs.py:165-166
      for _ in shape:
        lp, en = lp.sum(-1), en.sum(-1)
      logp, ent = logp + lp, ent + en
    return logp, ent

  def val_logits(self, x):
    return self.head('val', x, self.cfg.val_layers, 'logits')

  def slowval_logits(self, x):
    return self.head('slowval', x, self.cfg.val_layers, 'logits', self.slow)

  # -- policy (agent.py:115-135) ---------------------------------------------------
  def policy(self, carry, obs, is_first, noise):
    """carry: dict(deter, stoch, action); obs: observation dict (or the bare image);
    noise: dict(stoch=(N,S,C), action={key: ...} | tensor)."""
    _, _, act = space_specs(self.cfg)
    single = len(act) == 1
    with torch.no_grad():
      tokens = self.encoder(obs)
      dyn, feat = self.observe_step(
          carry, tokens, carry['action'], is_first, noise['stoch'])
      outs = self.pol_outputs(self.feat2tensor(feat['deter'], feat['stoch']))
      acts = self.pol_sample(outs, noise['action'])
    prev = acts[act[0][0]] if single and torch.is_tensor(carry['action']) else acts
    carry = dict(deter=dyn['deter'], stoch=dyn['stoch'], action=prev)
    out = {'dyn/deter': dyn['deter'], 'dyn/stoch': dyn['stoch']}
    return carry, acts, out

  # -- replay context (agent.py:312-340), K = replay_context -----------------------------
  def apply_replay_context(self, data, carry=None):
    """Rows whose chunk is the first of its window (consec == 0) restart from the latents
    stored in the replay (`truncate`: the last of the K context steps, rssm.py truncate);
    the other rows continue from the running `carry` (agent.py:331-339).  For K >= 1 the
    previous actions are the same in both branches: prepend(prev, act)[:, K:] == act[:, K-1:-1]."""
    K = self.cfg.replay_context
    rep = dict(deter=data['dyn/deter'][:, K - 1], stoch=data['dyn/stoch'][:, K - 1])
    if carry is not None and 'consec' in data:
      first = data['consec'][:, 0] == 0
      rep = dict(deter=torch.where(first[:, None], rep['deter'], carry['deter']),
                 stoch=torch.where(first[:, None, None], rep['stoch'], carry['stoch']))
    img, vec, act = space_specs(self.cfg)
    keys = [k for k, _ in img] + [v[0] for v in vec] + ['reward', 'is_first', 'is_last', 'is_terminal']
    obs = {k: data[k][:, K:] for k in keys}
    prevact = {a[0]: data[a[0]][:, K - 1: -1] for a in act}
    return rep, obs, prevact, data['stepid'][:, K:]

  # -- loss (agent.py:156-245) --------------------------------------------------------
  def loss(self, carry, obs, prevact, noise, update=True):
    """noise: dict(observe=(B,T,S,C), imag_stoch=(B*T,H,S,C), imag_act=(B*T,H+1,A))."""
    cfg = self.cfg
    reset = obs['is_first']
    B, T = reset.shape
    losses, metrics = {}, {}
    img, vec, act = space_specs(cfg)
    tokens = self.encoder(obs)
    carry, feat = self.observe(carry, tokens, prevact, reset, noise['observe'])
    los, mets = self.rssm_loss(feat)
    losses.update(los)
    metrics.update(mets)
    recon = self.decoder(feat['deter'], feat['stoch'])
    inp = self.feat2tensor(feat['deter'], feat['stoch'])
    losses['rew'] = twohot_loss(self.rew_logits(inp), self.bins, obs['reward'])
    con = (~obs['is_terminal']).to(f32)
    if cfg.contdisc:
      con = con * (1 - 1 / cfg.horizon)
    losses['con'] = -binary_logp(self.con_logit(inp), con)
    losses.update(self.recon_losses(recon, obs))

    # imagination: forward only (imgfeat is stop-gradient'ed, ac_grads False)
    K, H = T, cfg.imag_length
    with torch.no_grad():
      c = dict(deter=feat['deter'].reshape(B * K, -1),
               stoch=feat['stoch'].reshape(B * K, cfg.stoch, cfg.classes))
      deters, stochs, acts = [c['deter']], [c['stoch']], []
      anoise = noise['imag_act']
      if torch.is_tensor(anoise):
        anoise = {act[0][0]: anoise}
      for h in range(H + 1):
        outs = self.pol_outputs(self.feat2tensor(c['deter'], c['stoch']))
        a = self.pol_sample(outs, {k: v[:, h] for k, v in anoise.items()})
        acts.append(a)
        if h == H:
          break
        c = self.imagine_step(c, self.action_embed(a), noise['imag_stoch'][:, h])
        deters.append(c['deter'])
        stochs.append(c['stoch'])
      imgdeter, imgstoch = torch.stack(deters, 1), torch.stack(stochs, 1)
      imgact = {k: torch.stack([a[k] for a in acts], 1) for k in acts[0]}
    inp = self.feat2tensor(imgdeter, imgstoch)
    los, ret, mets = self.imag_loss(imgact, inp, update)
    losses.update({k: v.mean(1).reshape(B, K) for k, v in los.items()})
    metrics.update(mets)

    # replay value loss (agent.py:219-235, repl_loss :449-479)
    boot = ret[:, 0].reshape(B, K)
    inp = self.feat2tensor(feat['deter'], feat['stoch'])
    vlogits = self.val_logits(inp)
    val = twohot_pred(vlogits, self.bins)
    slow = twohot_pred(self.slowval_logits(inp), self.bins)
    disc = 1 - 1 / cfg.horizon
    weight = (~obs['is_last']).to(f32)
    rret = lambda_return(obs['is_last'], obs['is_terminal'], obs['reward'], val, boot, disc, cfg.lam)
    padded = torch.cat([rret, 0 * rret[:, -1:]], 1)
    losses['repval'] = weight[:, :-1] * (
        twohot_loss(vlogits, self.bins, padded) +
        cfg.slowreg * twohot_loss(vlogits, self.bins, slow))[:, :-1]

    scales = dict(cfg.scales)
    rec = scales.pop('image')                                   # agent.py:75-78
    scales.update({k: rec for k, _ in img})
    scales.update({v[0]: rec for v in vec})
    assert set(losses) == set(scales), (sorted(losses), sorted(scales))
    metrics.update({f'loss/{k}': v.mean() for k, v in losses.items()})
    total = sum(v.mean() * scales[k] for k, v in losses.items())
    entries = {'dyn/deter': feat['deter'], 'dyn/stoch': feat['stoch']}
    outs = dict(tokens=tokens, feat=feat, losses=losses, recon=recon,
                imgdeter=imgdeter, imgstoch=imgstoch, ret=ret,
                imgact=imgact[act[0][0]].long() if len(act) == 1 else imgact)
    return total, carry, entries, outs, metrics

  def imag_loss(self, act, inp, update):                        # agent.py:382-446
    cfg = self.cfg
    rew = twohot_pred(self.rew_logits(inp), self.bins)
    con = torch.exp(binary_logp(self.con_logit(inp), torch.ones(())))
    pol = self.pol_outputs(inp)
    vlogits = self.val_logits(inp)
    val = twohot_pred(vlogits, self.bins)
    slowval = twohot_pred(self.slowval_logits(inp), self.bins)
    tarval = val                                                # slowtar False
    disc = 1 if cfg.contdisc else 1 - 1 / cfg.horizon
    weight = torch.cumprod(disc * con, 1) / disc
    last = torch.zeros_like(con)
    term = 1 - con
    ret = lambda_return(last, term, rew, tarval, tarval, disc, cfg.lam)
    roffset, rscale = self.retnorm(ret, update)
    adv = (ret - tarval[:, :-1]) / rscale
    logpi, ent = self.pol_logp_entropy(pol, act)
    logpi, ent = logpi[:, :-1], ent[:, :-1]
    losses = {}
    losses['policy'] = weight[:, :-1].detach() * -(
        logpi * adv.detach() + cfg.actent * ent)
    padded = torch.cat([ret, 0 * ret[:, -1:]], 1).detach()
    losses['value'] = weight[:, :-1].detach() * (
        twohot_loss(vlogits, self.bins, padded) +
        cfg.slowreg * twohot_loss(vlogits, self.bins, slowval.detach()))[:, :-1]
    ret_normed = (ret - roffset) / rscale
    mets = dict(adv=adv.mean(), rew=rew.mean(), con=con.mean(), ret=ret_normed.mean(),
                val=val.mean(), weight=weight.mean(), ent=ent.mean())
    return losses, ret.detach(), mets

  def retnorm(self, x, update):                                 # utils.py:37-77 ('perc', debias False)
    cfg, st = self.cfg, self.state
    if update:
      x = x.detach().to(f32).flatten()
      lo = torch.quantile(x, cfg.perclo / 100)
      hi = torch.quantile(x, cfg.perchi / 100)
      st['ret_lo'] = (1 - cfg.retnorm_rate) * st['ret_lo'] + cfg.retnorm_rate * lo
      st['ret_hi'] = (1 - cfg.retnorm_rate) * st['ret_hi'] + cfg.retnorm_rate * hi
    lo, hi = st['ret_lo'], st['ret_hi']
    return lo, torch.clamp(hi - lo, min=cfg.retnorm_limit)

  # -- one train step (agent.py:137-154, opt.py:31-81) -----------------------------------
  def train(self, data, noise, carry=None):
    carry, obs, prevact, stepid = self.apply_replay_context(data, carry)
    names = list(self.p)
    leaves = [self.p[k].detach().requires_grad_(True) for k in names]
    saved = self.p
    self.p = dict(zip(names, leaves))
    total, carry, entries, outs, metrics = self.loss(carry, obs, prevact, noise, update=True)
    grads = torch.autograd.grad(total, leaves, allow_unused=True)
    self.p = saved
    grads = {k: (torch.zeros_like(self.p[k]) if g is None else g)
             for k, g in zip(names, grads)}
    self.apply_updates(grads)
    self.update_slow()
    metrics['loss'] = total.detach()
    replay = {'stepid': stepid, **{k: v.detach() for k, v in entries.items()}}
    return {k: v.detach() for k, v in carry.items()}, {'replay': replay}, metrics, grads, outs

  def learning_rate(self, count):                               # agent.py:368-378 (const + warmup)
    cfg = self.cfg
    if cfg.warmup and count < cfg.warmup:
      return cfg.lr * count / cfg.warmup
    return cfg.lr

  def apply_updates(self, grads):                               # opt.py:109-164
    cfg, st = self.cfg, self.state
    count = st['step']
    step = count + 1
    lr = self.learning_rate(count)
    for k, g in grads.items():
      w = self.p[k]
      unorm, pnorm = torch.linalg.norm(g.flatten()), torch.linalg.norm(w.flatten())
      upper = cfg.agc * torch.clamp(pnorm, min=cfg.pmin)
      u = g * (1 / torch.clamp(unorm / upper, min=1.0))
      st['nu'][k] = cfg.beta2 * st['nu'][k] + (1 - cfg.beta2) * (u * u)
      nu_hat = st['nu'][k] / (1 - cfg.beta2 ** step)
      u = u / (torch.sqrt(nu_hat) + cfg.eps)
      st['mu'][k] = (1 - cfg.beta1) * u + cfg.beta1 * st['mu'][k]
      mu_hat = st['mu'][k] / (1 - cfg.beta1 ** step)
      self.p[k] = w - lr * mu_hat
    st['step'] = step

  def update_slow(self):                                        # utils.py:113-119 (every 1)
    r = self.cfg.slowrate
    for k in self.slow:
      src = self.p[k.replace('slowval/', 'val/', 1)]
      self.slow[k] = r * src + (1 - r) * self.slow[k]


# ----------------------------------------------------------------- param shapes
def param_shapes(cfg):
  """Every optimised parameter and its shape, in the reference's naming; also the
  fan-in and outscale its initialiser uses (nets.py:144-197)."""
  D, H, S, C, g = cfg.deter, cfg.hidden, cfg.stoch, cfg.classes, cfg.blocks
  U = cfg.units
  img, vec, act = space_specs(cfg)
  A = sum(spec_width(a) for a in act)                           # DictConcat width of the action dict
  depths = [cfg.depth * m for m in cfg.mults]
  minres = cfg.image[0] // 2 ** len(cfg.mults) if img else 0
  sp = minres * minres * depths[-1]
  tokens = (U if vec else 0) + sp
  out = {}

  def lin(name, i, o, outscale=1.0):
    out[f'{name}/kernel'] = ((i, o), i, outscale)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def blk(name, i, o):
    out[f'{name}/kernel'] = ((g, i // g, o // g), i, 1.0)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def cnv(name, i, o, outscale=1.0):
    k = cfg.kernel
    out[f'{name}/kernel'] = ((k, k, i, o), k * k * i, outscale)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def nrm(name, n):
    out[f'{name}/scale'] = ((n,), None, None)

  if vec:                                                       # rssm.py:218-224
    i = sum(spec_width(v) for v in vec)
    for l in range(cfg.get('enc_layers', 3)):
      lin(f'enc/mlp{l}', i, U); nrm(f'enc/mlp{l}norm', U); i = U
  cin = cfg.image[2] if img else 0
  for i, d in enumerate(depths if img else []):
    cnv(f'enc/cnn{i}', cin, d); nrm(f'enc/cnn{i}norm', d); cin = d
  lin('dyn/dynin0', D, H); nrm('dyn/dynin0norm', H)
  lin('dyn/dynin1', S * C, H); nrm('dyn/dynin1norm', H)
  lin('dyn/dynin2', A, H); nrm('dyn/dynin2norm', H)
  blk('dyn/dynhid0', D + g * 3 * H, D); nrm('dyn/dynhid0norm', D)
  blk('dyn/dyngru', D, 3 * D)
  lin('dyn/obs0', D + tokens, H); nrm('dyn/obs0norm', H)
  lin('dyn/obslogit', H, S * C)
  lin('dyn/prior0', D, H); nrm('dyn/prior0norm', H)
  lin('dyn/prior1', H, H); nrm('dyn/prior1norm', H)
  lin('dyn/priorlogit', H, S * C)
  if vec:                                                       # rssm.py:326-334
    i = D + S * C
    for l in range(cfg.get('dec_layers', 3)):
      lin(f'dec/mlp/linear{l}', i, U); nrm(f'dec/mlp/norm{l}', U); i = U
    for v in vec:
      lin(f'dec/vec/{v[0]}/' + ('logits' if v[1] == 'disc' else 'pred'), U, spec_width(v))
  if img:
    blk('dec/sp0', D, sp)
    lin('dec/sp1', S * C, 2 * U); nrm('dec/sp1norm', 2 * U)
    lin('dec/sp2', 2 * U, sp)
    nrm('dec/spnorm', depths[-1])
    cin = depths[-1]
    for i in reversed(range(len(depths) - 1)):
      cnv(f'dec/conv{i}', cin, depths[i]); nrm(f'dec/conv{i}norm', depths[i]); cin = depths[i]
    cnv('dec/imgout', cin, cfg.image[2])
  F_ = D + S * C

  def headmlp(name, layers):
    i = F_
    for l in range(layers):
      lin(f'{name}/mlp/linear{l}', i, U); nrm(f'{name}/mlp/norm{l}', U); i = U

  headmlp('rew', cfg.rew_layers); lin('rew/head/logits', U, cfg.bins, 0.0)
  headmlp('con', cfg.con_layers); lin('con/head/logit', U, 1, 1.0)
  headmlp('pol', cfg.pol_layers)
  for a in act:                                                 # heads.py:103-112, 146-155
    n = int(np.prod(a[2], dtype=np.int64))
    if a[1] == 'disc':
      lin(f'pol/head/{a[0]}/logits', U, n * a[3], 0.01)
    else:
      lin(f'pol/head/{a[0]}/mean', U, n, 0.01); lin(f'pol/head/{a[0]}/stddev', U, n, 0.01)
  headmlp('val', cfg.val_layers); lin('val/head/logits', U, cfg.bins, 0.0)
  return out


def init_params(cfg, seed=0, outscale_override=None):
  """trunc_normal(-2,2) * 1.1368 / sqrt(fan_in) * outscale; biases 0; scales 1
  (nets.py:166-170).  `outscale_override` lets tests make the zero-initialised
  heads non-trivial."""
  gen = torch.Generator().manual_seed(seed)
  params = {}
  for name, (shape, fan, outscale) in param_shapes(cfg).items():
    if name.endswith('/scale'):
      params[name] = torch.ones(shape, dtype=f32)
    elif fan is None:
      params[name] = torch.zeros(shape, dtype=f32)
    else:
      if outscale_override is not None and outscale != 1.0:
        outscale = outscale_override
      x = torch.empty(shape, dtype=f32)
      torch.nn.init.trunc_normal_(x, 0.0, 1.0, -2.0, 2.0, generator=gen)
      params[name] = x * (1.1368 * math.sqrt(1 / fan) * outscale)
  return params


def make_noise(cfg, B, T, seed=0):
  gen = torch.Generator().manual_seed(seed)

  def gumbel(*shape):
    u = torch.rand(shape, generator=gen, dtype=f32).clamp_(1e-20, 1 - 1e-7)
    return -torch.log(-torch.log(u))
  S, C, H = cfg.stoch, cfg.classes, cfg.imag_length
  out = dict(observe=gumbel(B, T, S, C), imag_stoch=gumbel(B * T, H, S, C))
  _, _, act = space_specs(cfg)
  if len(act) == 1 and act[0][1:3] == ('disc', ()):             # the original layout: one tensor
    out['imag_act'] = gumbel(B * T, H + 1, act[0][3])
  else:
    out['imag_act'] = {
        key: gumbel(B * T, H + 1, *shape, classes) if kind == 'disc'
        else torch.randn((B * T, H + 1, *shape), generator=gen, dtype=f32)
        for key, kind, shape, classes in act}
  return out


def tiny_config(**over):
  """size1m-like (dreamerv3/configs.yaml:120-123) for tests that finish in seconds."""
  cfg = default_config(deter=512, hidden=64, classes=4, stoch=8, depth=4, units=64,
                       imag_length=5, bins=255)
  cfg.update(over)
  return cfg
