"""ppo agent behind the embodied Agent protocol (reference: ppo/agent.py:19-235, ppo/nets.py:10-71).

Same constructor, methods and replay entries as the reference: ``Agent(obs_space, act_space,
config)``, ``init_policy / init_train / init_report``, ``policy(carry, obs, mode) -> (carry, act,
out)`` with ``out = {logp/<key>, memory}`` (the entries ``ext_space`` declares, which the Driver
appends to the replay next to the transition), ``train(carry, data) -> (carry, {}, metrics)``.

Where the bytes live: observations arrive as device tensors staged by ``emb_driver_stage_obs``
(``device_obs``), actions / log-probabilities / the recurrent memory go back as device tensors and
are appended by ``emb_driver_scatter_mask_actions``; train batches are the dense device tensors of
``emb_replay_gather``.  Parameters, gradients and both Adam moments are four flat fp32 buffers, so
the optimiser chain of ``Agent._make_opt`` is ``emb_opt_clip_adam`` over one pointer and the
data-parallel gradient mean (embodied/jax/opt.py:52-54) is one NCCL call.  The advantage
recurrence of ``ppo_loss`` is ``emb_gae_advantage``.  The networks themselves (3x3 convolutions,
M = B*T GEMMs, the GRU) run as library GEMMs / convolutions under autograd, in ``compute_dtype``
(bfloat16 like the reference's default: bf16 operands, fp32 master weights, norms, heads and
losses; or float32 = the parity mode the tests compare with the oracle), and the whole update is
replayed as one CUDA graph after two eager warm-up steps.  There is no CPU path.
"""
import math
import re

import numpy as np
import torch
import torch.nn.functional as F

from .. import _lib
from .. import elements
from ..core import base
from ..dreamerv3 import params as paramlib
from . import config as configlib

f32 = torch.float32
EXCLUDE = ('is_first', 'is_last', 'is_terminal', 'reward')       # ppo/agent.py:136


def _clone(tree):
  if isinstance(tree, dict):
    return {k: _clone(v) for k, v in tree.items()}
  if isinstance(tree, (tuple, list)):
    return type(tree)(_clone(v) for v in tree)
  return tree.clone()


def _copy_into(dst, src):
  if isinstance(dst, dict):
    for k in dst:
      _copy_into(dst[k], src[k])
  elif isinstance(dst, (tuple, list)):
    for a, b in zip(dst, src):
      _copy_into(a, b)
  else:
    dst.copy_(src)


def _kind(space):
  """(discrete?, shape, classes) -- embodied/jax/nets.py:536 takes the LARGEST class count."""
  shape = tuple(int(x) for x in space.shape)
  if space.discrete:
    return True, shape, int(np.asarray(space.classes).max())
  return False, shape, 0


def _numel(shape):
  return int(np.prod(shape, dtype=np.int64))


def param_specs(cfg, obs_space, act_space):
  """name -> (shape, fan or None, outscale) in storage order; names are the reference's ninjax
  paths below `model/` (ppo/agent.py:134-159, ppo/nets.py:36-70, nets.py:503-587,634-669)."""
  enc = {k: v for k, v in obs_space.items() if k not in EXCLUDE and not k.startswith('log/')}
  vec = {k: v for k, v in enc.items() if len(v.shape) <= 2}
  img = {k: v for k, v in enc.items() if len(v.shape) == 3}
  out = {}

  def dense(name, i, o, outscale=1.0):
    out[f'{name}/kernel'] = ((i, o), i, outscale)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def norm(name, n, impl):
    if impl != 'none':
      out[f'{name}/scale'] = ((n,), None, None)
      out[f'{name}/shift'] = ((n,), None, 0.0)

  def embed(name, spaces, units):
    out[f'{name}/init'] = ((units,), units, 1.0)               # Initializer('trunc_normal', 'out')
    for key in sorted(spaces):
      disc, shape, classes = _kind(spaces[key])
      dense(f'{name}/{key}', _numel(shape) * (classes if disc else 1), units)

  width = 0
  if vec:
    embed('enc/emb', vec, cfg.enc_units)
    for i in range(cfg.enc_layers - 1):
      dense(f'enc/mlp/linear{i}', cfg.enc_units, cfg.enc_units)
      norm(f'enc/mlp/norm{i}', cfg.enc_units, cfg.enc_norm)
    width += cfg.enc_units
  if img:
    hw = {tuple(v.shape[:2]) for v in img.values()}
    assert len(hw) == 1, hw
    h, w = hw.pop()
    chans = sum(int(v.shape[2]) for v in img.values())
    for s, mult in enumerate(cfg.mults):
      d = cfg.depth * mult
      out[f'enc/s{s}in/kernel'] = ((3, 3, chans, d), 9 * chans, 1.0)
      out[f'enc/s{s}in/bias'] = ((d,), None, 0.0)
      h, w = -(-h // 2), -(-w // 2)
      for b in range(cfg.blocks):
        for c, n in (('c1', 'n1'), ('c2', 'n2')):
          norm(f'enc/s{s}b{b}{n}', d, cfg.enc_norm)
          out[f'enc/s{s}b{b}{c}/kernel'] = ((3, 3, d, d), 9 * d, 1.0)
          out[f'enc/s{s}b{b}{c}/bias'] = ((d,), None, 0.0)
      chans = d
    norm('enc/outn1', h * w * chans, cfg.enc_norm)
    dense('enc/outl', h * w * chans, cfg.outmult * cfg.depth)
    norm('enc/outn2', cfg.outmult * cfg.depth, cfg.enc_norm)
    width += cfg.outmult * cfg.depth
  feat = width
  if cfg.recurrent:
    inp = width
    if cfg.rnnact:
      embed('actemb', act_space, cfg.actemb_units)
      inp += cfg.actemb_units
    norm('rnn/norm', cfg.rnn_units + inp, cfg.rnn_norm)
    dense('rnn/linear', cfg.rnn_units + inp, 3 * cfg.rnn_units)
    feat = cfg.rnn_units

  def head(name, layers, units):
    n = feat
    for i in range(layers):
      dense(f'{name}/mlp/linear{i}', n, units)
      norm(f'{name}/mlp/norm{i}', units, cfg.head_norm)
      n = units
    return n

  n = head('policy', cfg.pol_layers, cfg.pol_units)
  for key, space in act_space.items():
    disc, shape, classes = _kind(space)
    if disc:
      dense(f'policy/head/{key}/logits', n, _numel(shape) * classes, cfg.pol_outscale)
    else:
      dense(f'policy/head/{key}/mean', n, _numel(shape), cfg.pol_outscale)
      dense(f'policy/head/{key}/stddev', n, _numel(shape), cfg.pol_outscale)
  n = head('value', cfg.val_layers, cfg.val_units)
  dense('value/head/pred', n, 1, cfg.val_outscale)
  return out


# ------------------------------------------------------------------------------ kernels
def gae(rew, val, last, term, hor, lam):
  """(rows, T) -> adv, tar (rows, T-1): ppo/agent.py:204-212 as one launch (no gradient: both
  results only enter the loss through stop_gradient)."""
  rows, T = rew.shape
  rew, val = rew.detach().to(f32).contiguous(), val.detach().to(f32).contiguous()
  last, term = last.contiguous().view(torch.uint8), term.contiguous().view(torch.uint8)
  adv = torch.empty((rows, T - 1), dtype=f32, device=rew.device)
  tar = torch.empty_like(adv)
  stream = torch.cuda.current_stream(rew.device).cuda_stream
  _lib.check(_lib.load().emb_gae_advantage(
      rew.data_ptr(), val.data_ptr(), last.data_ptr(), term.data_ptr(), adv.data_ptr(),
      tar.data_ptr(), rows, T, float(hor), float(lam), stream))
  return adv, tar


class ClipAdam:
  """optax.chain(clip_by_global_norm, scale_by_adam, add_decayed_weights, scale_by_learning_rate
  (linear_schedule(0, lr, warmup))) of ppo/agent.py:120-131 on the flat buffers."""

  def __init__(self, cfg, store):
    self.cfg, self.store = cfg, store
    dev = store.device
    self.state = torch.zeros(2, dtype=f32, device=dev)      # {updates so far, last gradient norm}
    self.scratch = torch.zeros(4 * 148 * 2, dtype=f32, device=dev)
    self.wdmask = None
    if cfg.wd:
      pattern = re.compile(cfg.wdregex)
      self.wdmask = torch.zeros(store.total, dtype=f32, device=dev)
      for name in store.specs:
        if pattern.search('/' + name):
          off = store.offsets[name]
          self.wdmask[off: off + _numel(store.specs[name][0])] = 1.0

  def launch(self):
    s, cfg = self.store, self.cfg
    stream = torch.cuda.current_stream(s.device).cuda_stream
    _lib.check(_lib.load().emb_opt_clip_adam(
        s.master.data_ptr(), s.grad.data_ptr(), s.mu.data_ptr(), s.nu.data_ptr(),
        self.wdmask.data_ptr() if self.wdmask is not None else None, s.total,
        self.scratch.data_ptr(), self.scratch.numel(), self.state.data_ptr(), float(cfg.lr),
        int(cfg.warmup), float(cfg.clip), float(cfg.eps), float(cfg.wd), 0.9, 0.999, stream))
    s.version += 1                                            # master was written through its raw pointer
    return self.state[1]


class Normalize:
  """embodied/jax/utils.py:16-91, impl 'meanstd' with debiasing; the three running scalars stay
  on the device (no host read in the train step)."""

  def __init__(self, rate, limit, device, world=1):
    self.rate, self.limit, self.world = rate, limit, world
    self.vars = torch.zeros(3, dtype=f32, device=device)      # mean, sqrs, corr

  def update(self, x):
    x = x.detach().to(f32)
    new = torch.stack([x.mean(), (x * x).mean(), torch.ones((), dtype=f32, device=x.device)])
    if self.world > 1:                                        # utils.py:76-81 pmean over the data axes
      torch.distributed.all_reduce(new, op=torch.distributed.ReduceOp.AVG)
    self.vars.mul_(1 - self.rate).add_(new, alpha=self.rate)

  def stats(self):
    mean, sqrs, corr = self.vars.unbind(0)
    corr = 1.0 / torch.clamp(corr, min=self.rate)
    mean = mean * corr
    std = torch.sqrt(torch.relu(sqrs * corr - mean * mean))
    return mean, torch.clamp(std, min=self.limit)


# ------------------------------------------------------------------------------ networks
class Model:
  """ppo/agent.py:134-183 on a ParamStore.  All tensors fp32 on the device."""

  def __init__(self, cfg, obs_space, act_space, store):
    self.cfg, self.store = cfg, store
    self.act_space = act_space
    self.cd = store.compute_dtype          # GEMMs / convolutions / GRU in it; norms, heads, losses in fp32
    enc = {k: v for k, v in obs_space.items() if k not in EXCLUDE and not k.startswith('log/')}
    assert all(len(s.shape) <= 3 for s in enc.values()), enc          # ppo/nets.py:23
    self.vec = {k: v for k, v in enc.items() if len(v.shape) <= 2}
    self.img = {k: v for k, v in enc.items() if len(v.shape) == 3}
    self.actkind = {k: _kind(v) for k, v in act_space.items()}

  def w(self, name):
    return self.store.get(name)

  # -- layers
  def linear(self, name, x):
    return torch.addmm(self.w(f'{name}/bias'), x.reshape(-1, x.shape[-1]).to(self.cd),
                       self.w(f'{name}/kernel')).reshape(*x.shape[:-1], -1)

  def norm(self, name, x, impl):                                      # nets.py:374-399
    if impl == 'none':
      return x
    assert impl == 'layer', impl
    xf = x.to(f32)                                                    # Norm computes in f32 (nets.py:376)
    mean = xf.mean(-1, keepdim=True)
    var = torch.clamp((xf * xf).mean(-1, keepdim=True) - mean * mean, min=0)
    sw = self.store.w                                                 # fp32 scale / shift
    y = (xf - mean) * (torch.rsqrt(var + self.cfg.norm_eps) * sw[f'{name}/scale']) + sw[f'{name}/shift']
    return y.to(x.dtype)

  def act(self, name, x):
    return {'relu': torch.relu, 'silu': F.silu, 'none': lambda y: y}[name](x)

  def conv(self, name, x):                                            # NCHW activations, HWIO kernel
    return F.conv2d(x, self.w(f'{name}/kernel').permute(3, 2, 0, 1), self.w(f'{name}/bias'), padding=1)

  def embed(self, name, spaces, xs, bshape, units, squish):           # nets.py:520-562
    total = self.w(f'{name}/init').expand(*bshape, units)
    lead = len(bshape)
    for key in sorted(spaces):
      disc, shape, classes = _kind(spaces[key])
      x = xs[key]
      if x.dtype.is_floating_point:                                   # nets.py:80-100 `available`
        ok = (x != float('-inf')).reshape(*bshape, -1).all(-1)
      elif x.dtype in (torch.int8, torch.int16, torch.int32, torch.int64):
        ok = (x != -1).reshape(*bshape, -1).all(-1)
      else:
        ok = None
      if ok is not None:
        x = torch.where(ok.reshape(*bshape, *([1] * (x.ndim - lead))), x, torch.zeros_like(x))
      if disc:
        x = F.one_hot(x.long(), classes).to(self.cd)
      else:
        x = squish(x.to(f32)).to(self.cd)
      x = self.linear(f'{name}/{key}', x.reshape(*bshape, -1))
      if ok is not None:
        x = torch.where(ok[..., None], x, torch.zeros_like(x))
      total = total + x
    return total

  def encoder(self, obs, bshape):                                     # ppo/nets.py:29-71
    cfg = self.cfg
    outs = []
    if self.vec:
      squish = (lambda y: torch.sign(y) * torch.log1p(torch.abs(y))) if cfg.symlog else (lambda y: y)
      x = self.embed('enc/emb', self.vec, obs, bshape, cfg.enc_units, squish)
      x = x.reshape(-1, x.shape[-1])
      for i in range(cfg.enc_layers - 1):
        x = self.linear(f'enc/mlp/linear{i}', x)
        x = self.act(cfg.enc_act, self.norm(f'enc/mlp/norm{i}', x, cfg.enc_norm))
      outs.append(x)
    if self.img:
      x = torch.cat([obs[k] for k in sorted(self.img)], -1)
      assert x.dtype == torch.uint8, x.dtype
      x = x.reshape(-1, *x.shape[-3:]).permute(0, 3, 1, 2).to(self.cd) * 255 - 0.5   # sic, ppo/nets.py:46
      for s in range(len(cfg.mults)):
        x = self.conv(f'enc/s{s}in', x)
        # reduce_window(max, 3x3, stride 2, 'same', init -inf): pad like XLA's SAME (low = total // 2)
        pads = []
        for n in (x.shape[3], x.shape[2]):
          total = max((-(-n // 2) - 1) * 2 + 3 - n, 0)
          pads += [total // 2, total - total // 2]
        x = F.max_pool2d(F.pad(x, pads, value=float('-inf')), 3, 2)
        for b in range(cfg.blocks):
          skip = x
          x = self.act(cfg.enc_act, self._cnorm(f'enc/s{s}b{b}n1', x))
          x = self.conv(f'enc/s{s}b{b}c1', x)
          x = self.act(cfg.enc_act, self._cnorm(f'enc/s{s}b{b}n2', x))
          x = self.conv(f'enc/s{s}b{b}c2', x)
          x = x + skip
      x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)               # the reference flattens NHWC
      x = self.act(cfg.enc_act, self.norm('enc/outn1', x, cfg.enc_norm))
      x = self.linear('enc/outl', x)
      x = self.act(cfg.enc_act, self.norm('enc/outn2', x, cfg.enc_norm))
      outs.append(x)
    return torch.cat(outs, -1).reshape(*bshape, -1)

  def _cnorm(self, name, x):
    if self.cfg.enc_norm == 'none':
      return x
    return self.norm(name, x.permute(0, 2, 3, 1), self.cfg.enc_norm).permute(0, 3, 1, 2)

  def gru_step(self, carry, inp, reset):                              # nets.py:657-669
    U = self.cfg.rnn_units
    carry = carry * (~reset)[:, None].to(carry.dtype)
    x = self.norm('rnn/norm', torch.cat([carry, inp], -1), self.cfg.rnn_norm)
    x = self.linear('rnn/linear', x)
    res, cand, update = x[:, :U], x[:, U:2 * U], x[:, 2 * U:]
    cand = torch.tanh(torch.sigmoid(res) * cand)
    update = torch.sigmoid(update - 1.0)
    return update * cand + (1 - update) * carry

  def head_mlp(self, name, x, layers):
    for i in range(layers):
      x = self.linear(f'{name}/mlp/linear{i}', x)
      x = self.act(self.cfg.head_act, self.norm(f'{name}/mlp/norm{i}', x, self.cfg.head_norm))
    return x

  def policy_outputs(self, feat):                                     # heads.py:100-112,146-155
    cfg = self.cfg
    h = self.head_mlp('policy', feat, cfg.pol_layers)
    outs = {}
    for key, (disc, shape, classes) in self.actkind.items():
      if disc:
        y = self.linear(f'policy/head/{key}/logits', h).to(f32)       # outs.* work in f32
        outs[key] = y.reshape(*y.shape[:-1], *shape, classes)
      else:
        mean = self.linear(f'policy/head/{key}/mean', h).to(f32)
        std = self.linear(f'policy/head/{key}/stddev', h).to(f32)
        std = (cfg.maxstd - cfg.minstd) * torch.sigmoid(std + 2.0) + cfg.minstd
        outs[key] = (torch.tanh(mean).reshape(*mean.shape[:-1], *shape),
                     std.reshape(*std.shape[:-1], *shape))
    return outs

  def logp_entropy(self, outs, acts):
    """outs.Categorical / outs.Normal logp and entropy, event axes summed (outs.Agg)."""
    logps, ents = {}, {}
    for key, (disc, shape, classes) in self.actkind.items():
      if disc:
        la = torch.log_softmax(outs[key], -1)
        lp = la.gather(-1, acts[key].long()[..., None]).squeeze(-1)
        en = -(la.exp() * la).sum(-1)
      else:
        mean, std = outs[key]
        lp = -0.5 * ((acts[key].to(f32) - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
        en = 0.5 * torch.log(2 * math.pi * std * std) + 0.5
      if shape:
        axes = tuple(range(-len(shape), 0))
        lp, en = lp.sum(axes), en.sum(axes)
      logps[key], ents[key] = lp, en
    return logps, ents

  def sample(self, outs, noise):
    acts = {}
    for key, (disc, shape, classes) in self.actkind.items():
      if disc:
        acts[key] = torch.argmax(outs[key] + noise[key], -1).to(torch.int32)
      else:
        mean, std = outs[key]
        acts[key] = noise[key] * std + mean
    return acts

  def value(self, feat):
    h = self.head_mlp('value', feat, self.cfg.val_layers)
    return self.linear('value/head/pred', h).squeeze(-1).to(f32)

  def __call__(self, memory, obs, prevact, value=True, single=False):  # ppo/agent.py:162-183
    cfg = self.cfg
    first = obs['is_first']
    bshape = tuple(first.shape[:1 if single else 2])
    embed = self.encoder(obs, bshape)
    if cfg.recurrent:
      inputs = embed
      if cfg.rnnact:
        masked = {k: torch.where(first.reshape(*bshape, *([1] * (v.ndim - len(bshape)))),
                                 torch.zeros_like(v), v) for k, v in prevact.items()}
        clip = lambda x: x / torch.clamp(torch.abs(x), min=1.0).detach()
        inputs = torch.cat([embed, self.embed('actemb', self.act_space, masked, bshape,
                                              cfg.actemb_units, clip)], -1)
      if single:
        memory = feat = self.gru_step(memory, inputs, first)
      else:
        feats = []
        for t in range(bshape[1]):
          memory = self.gru_step(memory, inputs[:, t], first[:, t])
          feats.append(memory)
        feat = torch.stack(feats, 1)
    else:
      feat = embed
    policy = self.policy_outputs(feat)
    return memory, feat, policy, (self.value(feat) if value else None)


class Agent(base.Agent):

  device_obs = True

  def __init__(self, obs_space, act_space, config=None, device=None, values=None):
    if not torch.cuda.is_available():
      raise RuntimeError('embodied_b200.ppo.Agent runs on a CUDA device; none is visible and there '
                         'is no CPU fallback.')
    _lib.load()
    self.obs_space = dict(obs_space)
    self.act_space = {k: v for k, v in act_space.items() if k != 'reset'}
    cfg = config if isinstance(config, configlib.Config) else configlib.make(**(config or {}))
    self.cfg = cfg = configlib.Config(cfg)
    cfg.setdefault('norm_eps', 1e-4)
    self.cd = {'float32': f32, 'bfloat16': torch.bfloat16}[cfg.compute_dtype]
    self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
    self.world, self.rank = 1, 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
      self.world, self.rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
    specs = param_specs(cfg, self.obs_space, self.act_space)
    self.store = paramlib.ParamStore(cfg, self.device, self.cd, cfg.seed, values, specs=specs)
    self.model = Model(cfg, self.obs_space, self.act_space, self.store)
    self.opt = ClipAdam(cfg, self.store)
    self.advnorm = Normalize(cfg.norm_rate, cfg.norm_limit, self.device, self.world)
    self.valnorm = Normalize(cfg.norm_rate, cfg.norm_limit, self.device, self.world)
    self.gen = torch.Generator(device=self.device)
    self.gen.manual_seed(cfg.seed * 1000003 + self.rank)
    self.updates = 0
    self.obskeys = list(self.model.vec) + list(self.model.img)
    # the update as ONE CUDA graph per batch signature (after GRAPH_WARMUP eager steps): the
    # T-step GRU loop and its backward are ~2500 launches that cost one cudaGraphLaunch
    self._graph_mode = cfg.get('graph', 'auto')
    self._graphs, self._graph_seen, self._graph_ok = {}, {}, True

  # ------------------------------------------------------------------- plugin properties
  @property
  def policy_keys(self):
    return '/(enc|actemb|rnn|policy)/'

  @property
  def ext_space(self):                                                # ppo/agent.py:42-51
    S = elements.Space
    spaces = {'consec': S(np.int32), 'stepid': S(np.uint8, 20)}
    for key in self.act_space:
      spaces[f'logp/{key}'] = S(np.float32)
    if self.cfg.recurrent and self.cfg.replay_context:
      spaces['memory'] = S(np.float32, self.cfg.rnn_units)
    return spaces

  def _initial(self, batch_size):
    if self.cfg.recurrent:
      return torch.zeros((batch_size, self.cfg.rnn_units), dtype=self.cd, device=self.device)
    return ()

  def init_policy(self, batch_size):                                  # ppo/agent.py:53-58
    prevact = {k: torch.zeros((batch_size, *v.shape), device=self.device,
                              dtype=torch.int32 if v.discrete else f32)
               for k, v in self.act_space.items()}
    return self._initial(batch_size), prevact

  init_train = init_policy

  def init_report(self, batch_size):
    return ()

  def _flags(self):
    if self.cd == f32:                       # parity mode: strict IEEE GEMMs / convolutions
      torch.backends.cuda.matmul.allow_tf32 = False
      torch.backends.cudnn.allow_tf32 = False

  def _dev(self, x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), device=self.device)

  def action_noise(self, lead):
    """Gumbel noise per categorical key (arg-max of logits + noise == a categorical draw), standard
    normal per continuous key (sampling noise is an explicit input)."""
    noise = {}
    for key, (disc, shape, classes) in self.model.actkind.items():
      if disc:
        u = torch.rand((*lead, *shape, classes), device=self.device, generator=self.gen)
        noise[key] = -torch.log(-torch.log(u.clamp_(1e-6, 1 - 1e-6)))
      else:
        noise[key] = torch.randn((*lead, *shape), device=self.device, generator=self.gen)
    return noise

  # ---------------------------------------------------------------------------- policy
  @torch.no_grad()
  def policy(self, carry, obs, mode='train', noise=None):             # ppo/agent.py:70-79
    self._flags()
    assert not any(k.startswith('log/') for k in obs), list(obs)
    memory, prevact = carry
    inputs = {k: self._dev(obs[k]) for k in (*self.obskeys, 'is_first')}
    n = len(inputs['is_first'])
    if noise is None:
      noise = self.action_noise((n,))
    self.store.begin_step()
    memory, _, pol, _ = self.model(memory, inputs, prevact, value=False, single=True)
    act = self.model.sample(pol, noise)
    logps, _ = self.model.logp_entropy(pol, act)
    out = {f'logp/{k}': v for k, v in logps.items()}
    if self.cfg.recurrent:
      out['memory'] = memory.to(f32)                                  # replay rows are f32 (ext_space)
    return (memory, act), dict(act), out

  # ----------------------------------------------------------------------------- train
  def ppo_loss(self, data, policy, value, update=True):               # ppo/agent.py:177-235
    cfg, m = self.cfg, self.model
    act = {k: data[k] for k in self.act_space}
    logps, ents = m.logp_entropy(policy, act)
    logpi = sum(logps.values())
    logdata = sum(data['logp/' + k] for k in self.act_space)
    rew, last, term = data['reward'], data['is_last'], data['is_terminal']
    mask = (~last & ~term).to(f32)
    ratio = torch.exp(logpi - logdata.detach())
    voffset, vscale = self.valnorm.stats()
    val = value * vscale + voffset
    adv, tar = gae(rew, val, last, term, cfg.hor, cfg.lam)
    if update:
      self.valnorm.update(tar)
    voffset, vscale = self.valnorm.stats()
    tarnormed = (tar - voffset) / vscale
    if cfg.tarclip:
      tarnormed = torch.clamp(tarnormed, -cfg.tarclip, cfg.tarclip)
    losses = {'value': (value - F.pad(tarnormed, (0, 1))) ** 2 * mask}   # outs.MSE.loss, target padded with 0
    if update:
      self.advnorm.update(adv)
    aoffset, ascale = self.advnorm.stats()
    advnormed = (adv - aoffset) / ascale
    part = ratio[:, :-1]
    maxent = cfg.actent * sum(ents.values())[:, :-1]
    upper = (part < 1 + cfg.trclip) | (advnormed < 0)
    lower = (part > 1 - cfg.trclip) | (advnormed > 0)
    tr = (upper & lower).to(f32)
    losses['policy'] = -(part * advnormed + maxent) * mask[:, :-1] * tr
    metrics = {}
    for k, (disc, shape, classes) in m.actkind.items():
      metrics[f'ent/{k}'] = ents[k].mean()
      if not shape:                                                   # Agg outputs carry no minent / maxent
        if disc:
          lo, hi = 0.0, math.log(classes)
        else:
          ent = lambda s: 0.5 * math.log(2 * math.pi * s * s) + 0.5
          lo, hi = ent(cfg.minstd), ent(cfg.maxstd)
        metrics[f'rand/{k}'] = (ents[k].mean() - lo) / (hi - lo)
    metrics.update(rew=rew.mean(), val=val.mean(), tar=tar.mean(), adv=adv.mean(),
                   advmag=adv.abs().mean(), ratio=ratio.mean(), clipfrac=(1 - tr).mean(),
                   td=(value[:, :-1] - tarnormed).abs().mean())
    return losses, metrics

  def loss(self, memory, data, prevact, update=True):                 # ppo/agent.py:102-112
    memory, _, policy, value = self.model(memory, data, prevact)
    losses, metrics = self.ppo_loss(data, policy, value, update)
    for k, v in losses.items():
      metrics[f'{k}_loss'] = v.mean()
      metrics[f'{k}_loss_std'] = v.std(unbiased=False)
    total = sum(v.mean() * self.cfg.scales[k] for k, v in losses.items())
    return total, memory, metrics, losses

  def _context(self, carry, data):                                    # ppo/agent.py:82-91
    cfg = self.cfg
    memory, prevact = carry
    data = {k: v for k, v in data.items() if k not in ('stepid', 'consec')}
    if cfg.replay_context:
      K = cfg.replay_context
      prevact = {k: data[k][:, K - 1: -1] for k in self.act_space}
      data = {k: v[:, K:] for k, v in data.items()}
      if cfg.recurrent:
        memory = data.pop('memory').to(self.cd)[:, K - 1]             # sic: row K-1 of the sliced rows
    else:
      prevact = {k: torch.cat([prevact[k][:, None], data[k][:, :-1]], 1) for k in self.act_space}
    return memory, prevact, data

  GRAPH_WARMUP = 2

  def _device_step(self, carry, data):
    """The device work of one update (ppo/agent.py:81-97): pure stream work, no host reads."""
    memory, prevact, data = self._context(carry, data)
    self.store.begin_step()
    self.store.grad.zero_()
    total, memory, metrics, losses = self.loss(memory, data, prevact)
    total.backward()
    if self.world > 1:                                                # embodied/jax/opt.py:52-54
      torch.distributed.all_reduce(self.store.grad, op=torch.distributed.ReduceOp.AVG)
    metrics = {k: v.detach() for k, v in metrics.items()}
    metrics['loss'] = total.detach()
    metrics['opt/grad_norm'] = self.opt.launch()
    if self.cd != f32:
      self.store.refresh_low()                                        # the bf16 copy the next forward reads
    self.last_losses = {k: v.detach() for k, v in losses.items()}
    prevact = {k: data[k][:, -1].clone() for k in self.act_space}
    memory = memory.detach() if self.cfg.recurrent else memory
    return (memory, prevact), metrics

  def train(self, carry, data):                                       # ppo/agent.py:81-97
    self._flags()
    if self._graph_mode not in (False, 'off') and self._graph_ok and self.world == 1:
      key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(data.items()))
      st = self._graphs.get(key)
      if st is None:
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if seen >= self.GRAPH_WARMUP:
          st = self._capture(key, carry, data)
      if st is not None:
        return self._train_graphed(st, carry, data)
    carry, metrics = self._device_step(carry, data)
    self.updates += 1
    metrics['opt/updates'] = self.updates
    return carry, {}, metrics

  def _capture(self, key, carry, data):
    import types
    import warnings
    st = types.SimpleNamespace()
    st.data = {k: v.clone() for k, v in data.items()}
    st.carry = _clone(carry)
    st.graph = torch.cuda.CUDAGraph()
    launched = _lib.launch_count()
    try:
      torch.cuda.synchronize()
      with torch.cuda.graph(st.graph):
        st.carry_out, metrics = self._device_step(st.carry, st.data)
        st.names = list(metrics)
        st.mvec = torch.stack([metrics[k].to(f32).reshape(()) for k in st.names])
      torch.cuda.synchronize()
    except Exception as e:      # noqa: BLE001 -- capture refused: stay on eager launches (still the CUDA path)
      warnings.warn(f'ppo.Agent: CUDA graph capture of the update failed ({e!r}); continuing with eager launches')
      self._graph_ok = False
      try:
        torch.cuda.synchronize()
      except Exception:         # noqa: BLE001
        pass
      return None
    # a capture records, it does not execute: the caller replays it now
    st.launches = _lib.launch_count() - launched
    _lib.launch_count_add((1 << 64) - st.launches)
    self._graphs[key] = st
    return st

  def _train_graphed(self, st, carry, data):
    _copy_into(st.carry, carry)
    for k, v in st.data.items():
      v.copy_(data[k])
    st.graph.replay()
    _lib.launch_count_add(st.launches)
    self.updates += 1
    mv = st.mvec.clone()
    metrics = {k: mv[i] for i, k in enumerate(st.names)}
    metrics['opt/updates'] = self.updates
    return st.carry_out, {}, metrics

  def report(self, carry, data):                                      # ppo/agent.py:99-100
    return carry, {}

  def stream(self, st):
    return st

  def save(self):
    data = self.store.state_dict()
    data['opt/step'] = np.asarray(self.updates)
    data['opt/count'] = self.opt.state.cpu().numpy()
    data['advnorm'] = self.advnorm.vars.cpu().numpy()
    data['valnorm'] = self.valnorm.vars.cpu().numpy()
    return data

  def load(self, data, regex=None):
    if regex:
      pattern = re.compile(regex)
      keep = {k: v for k, v in data.items() if k in self.store.specs and pattern.match(k)}
      if not keep:
        raise KeyError(f'no parameter matches {regex!r}')
      self.store.load_params(keep)
      return
    self.store.load_state_dict(data)
    self.updates = int(data.get('opt/step', 0))
    if 'opt/count' in data:
      self.opt.state.copy_(torch.as_tensor(data['opt/count']))
      self.advnorm.vars.copy_(torch.as_tensor(data['advnorm']))
      self.valnorm.vars.copy_(torch.as_tensor(data['valnorm']))
