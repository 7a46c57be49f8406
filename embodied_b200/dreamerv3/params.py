"""Parameter storage of the dreamerv3 learner, laid out for the B200.

All optimised parameters live in ONE float32 device buffer (`master`), each
tensor at a 64-byte aligned offset; gradients (`grad`), RMS (`nu`) and momentum
(`mu`) state are three more buffers of the same layout.  That makes the
gradient all-reduce one NCCL call on one pointer and the optimiser one fused
launch over the flat buffer (embodied/jax/opt.py:31-81 applies the optax chain
tensor by tensor).  The model casts a weight to the compute dtype where it is
used, as the reference does (embodied/jax/nets.py:243,272,301); the cast's
backward adds the bf16 gradient into the float32 `grad` view.

Names and shapes follow the reference's ninjax paths ('dyn/dynin0/kernel',
'enc/cnn0norm/scale', ...) so checkpoints are interchangeable key for key.
Initialisation: trunc_normal(-2, 2) * 1.1368 / sqrt(fan_in) * outscale, zero
biases, unit norm scales (embodied/jax/nets.py:144-197, 236-251).
"""
import math

import numpy as np
import torch

ALIGN = 16   # elements (64 B in fp32)


def shapes(cfg):
  """name -> (shape, fan_in or None, outscale).  Order is the storage order."""
  from . import spaces
  D, H, S, C, g = cfg.deter, cfg.hidden, cfg.stoch, cfg.classes, cfg.blocks
  U, A, k = cfg.units, cfg.actions, cfg.kernel
  image = cfg.get('image')
  vecspec, actspec = cfg.get('vecspec') or [], cfg.get('actspec')
  if actspec is None:                       # the original single discrete `action` key
    actspec = [('action', 'disc', (), A)]
  depths = [cfg.depth * m for m in cfg.mults]
  sp = 0
  if image:
    minres = image[0] // 2 ** len(cfg.mults)
    sp = minres * minres * depths[-1]
  tokens = (U if vecspec else 0) + sp
  out = {}

  def dense(name, i, o, outscale=1.0):
    out[f'{name}/kernel'] = ((i, o), i, outscale)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def block(name, i, o):
    out[f'{name}/kernel'] = ((g, i // g, o // g), i, 1.0)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def conv(name, i, o):
    out[f'{name}/kernel'] = ((k, k, i, o), k * k * i, 1.0)
    out[f'{name}/bias'] = ((o,), None, 0.0)

  def norm(name, n):
    out[f'{name}/scale'] = ((n,), None, None)

  # dyn first: the scan kernels want these contiguous
  dense('dyn/dynin0', D, H); norm('dyn/dynin0norm', H)
  dense('dyn/dynin1', S * C, H); norm('dyn/dynin1norm', H)
  dense('dyn/dynin2', A, H); norm('dyn/dynin2norm', H)
  block('dyn/dynhid0', D + g * 3 * H, D); norm('dyn/dynhid0norm', D)
  block('dyn/dyngru', D, 3 * D)
  dense('dyn/obs0', D + tokens, H); norm('dyn/obs0norm', H)
  dense('dyn/obslogit', H, S * C)
  dense('dyn/prior0', D, H); norm('dyn/prior0norm', H)
  dense('dyn/prior1', H, H); norm('dyn/prior1norm', H)
  dense('dyn/priorlogit', H, S * C)
  feat = D + S * C
  if vecspec:                               # rssm.py:218-224: DictConcat -> MLP
    i = sum(spaces.width(v) for v in vecspec)
    for l in range(cfg.get('enc_layers', 3)):
      dense(f'enc/mlp{l}', i, U); norm(f'enc/mlp{l}norm', U); i = U
  if image:
    cin = image[2]
    for i, d in enumerate(depths):
      conv(f'enc/cnn{i}', cin, d); norm(f'enc/cnn{i}norm', d); cin = d
  if vecspec:                               # rssm.py:326-334: MLP -> one head per key
    i = feat
    for l in range(cfg.get('dec_layers', 3)):
      dense(f'dec/mlp/linear{l}', i, U); norm(f'dec/mlp/norm{l}', U); i = U
    for v in vecspec:
      dense(f'dec/vec/{v[0]}/' + ('logits' if v[1] == 'disc' else 'pred'), U, spaces.width(v))
  if image:
    block('dec/sp0', D, sp)
    dense('dec/sp1', S * C, 2 * U); norm('dec/sp1norm', 2 * U)
    dense('dec/sp2', 2 * U, sp)
    norm('dec/spnorm', depths[-1])
    cin = depths[-1]
    for i in reversed(range(len(depths) - 1)):
      conv(f'dec/conv{i}', cin, depths[i]); norm(f'dec/conv{i}norm', depths[i])
      cin = depths[i]
    conv('dec/imgout', cin, image[2])

  def head(name, layers, outs):
    i = feat
    for l in range(layers):
      dense(f'{name}/mlp/linear{l}', i, U); norm(f'{name}/mlp/norm{l}', U); i = U
    for outname, outdim, outscale in outs:
      dense(f'{name}/head/{outname}', U, outdim, outscale)

  head('rew', cfg.rew_layers, [('logits', cfg.bins, 0.0)])
  head('con', cfg.con_layers, [('logit', 1, 1.0)])
  pol = []
  for a in actspec:                         # heads.py:103-112 categorical / :146-155 bounded_normal
    if a[1] == 'disc':
      pol.append((f'{a[0]}/logits', spaces.width(a), 0.01))
    else:
      pol += [(f'{a[0]}/mean', spaces.size(a), 0.01), (f'{a[0]}/stddev', spaces.size(a), 0.01)]
  head('pol', cfg.pol_layers, pol)
  head('val', cfg.val_layers, [('logits', cfg.bins, 0.0)])
  return out


class _CastParam(torch.autograd.Function):
  """leaf (f32 master view) -> its compute-dtype copy `low`; backward accumulates
  into the leaf's slice of the flat gradient buffer in one mixed-precision add
  (what the cast's backward + AccumulateGrad would do in two kernels)."""

  @staticmethod
  def forward(ctx, leaf, low, grad, store, name):
    ctx.grad, ctx.store, ctx.name = grad, store, name
    return low.view(low.shape)

  @staticmethod
  def backward(ctx, g):
    ctx.grad.add_(g)
    if ctx.store.on_grad is not None:         # bucketed exchange: this tensor's gradient is final
      ctx.store.on_grad(ctx.name)
    return None, None, None, None, None


class ParamStore:

  def __init__(self, cfg, device, compute_dtype, seed=0, values=None, specs=None):
    self.device = torch.device(device)
    self.compute_dtype = compute_dtype
    # `specs`: another agent's name -> (shape, fan, outscale) table (ppo) in place of dreamerv3's
    self.specs = shapes(cfg) if specs is None else dict(specs)
    self.offsets, total = {}, 0
    for name, (shape, _, _) in self.specs.items():
      self.offsets[name] = total
      total += (int(np.prod(shape)) + ALIGN - 1) // ALIGN * ALIGN
    self.total = total
    self.count = sum(int(np.prod(s[0])) for s in self.specs.values())
    f32 = torch.float32
    self.master = torch.zeros(total, dtype=f32, device=self.device)
    self.grad = torch.zeros(total, dtype=f32, device=self.device)
    self.nu = torch.zeros(total, dtype=f32, device=self.device)
    self.mu = torch.zeros(total, dtype=f32, device=self.device)
    self.step = 0
    self.version = 0      # bumped whenever `master` changes (packed copies key on it)
    # segment table for the per-tensor AGC norms of the fused optimiser
    names = list(self.specs)
    self.seg_begin = torch.tensor(
        [self.offsets[n] for n in names], dtype=torch.int64, device=self.device)
    self.seg_size = torch.tensor(
        [int(np.prod(self.specs[n][0])) for n in names], dtype=torch.int64,
        device=self.device)
    self._init(seed, values)
    # autograd leaves: float32 views of `master` whose .grad are views of
    # `grad`, so backward accumulates straight into the flat gradient buffer.
    self.w = {}
    self.on_grad = None       # callback(name) when a tensor's gradient has been accumulated
    for name in names:
      leaf = self._view(self.master, name).requires_grad_(True)
      leaf.grad = self._view(self.grad, name)
      leaf.register_post_accumulate_grad_hook(
          lambda p, n=name: self.on_grad(n) if self.on_grad is not None else None)
      self.w[name] = leaf
    self._cast = {}
    self.low, self._low_seen = None, -1      # flat compute-dtype copy of master (get())
    # slow value network (utils.py:94-127): a separate small buffer
    self.slow = {n.replace('val/', 'slowval/', 1): self.view('master', n).clone()
                 for n in names if n.startswith('val/')}
    self.refresh_low()       # allocated here, never inside a stream capture

  def _view(self, buf, name):
    shape = self.specs[name][0]
    off = self.offsets[name]
    return buf[off: off + int(np.prod(shape))].view(shape)

  def view(self, which, name):
    return self._view(getattr(self, which), name)

  def _init(self, seed, values):
    gen = torch.Generator().manual_seed(seed)
    for name, (shape, fan, outscale) in self.specs.items():
      if values is not None:
        x = torch.as_tensor(np.asarray(values[name]), dtype=torch.float32)
        assert tuple(x.shape) == tuple(shape), (name, x.shape, shape)
      elif name.endswith('/scale'):
        x = torch.ones(shape)
      elif fan is None:
        x = torch.zeros(shape)
      else:
        x = torch.empty(shape)
        torch.nn.init.trunc_normal_(x, 0.0, 1.0, -2.0, 2.0, generator=gen)
        x = x * (1.1368 * math.sqrt(1 / fan) * outscale)
      self.view('master', name).copy_(x)

  def get(self, name):
    """The parameter in the compute dtype (embodied/jax/nets.py:243: parameters
    are cast to the compute dtype where they are used, gradients arrive in f32).
    The low-precision values are views of ONE flat copy of `master`, refreshed
    by a single kernel after every update; the backward of the cast adds the
    low-precision gradient straight into the flat f32 gradient buffer."""
    if self.compute_dtype == torch.float32:
      return self.w[name]
    tracked = torch.is_grad_enabled()    # a no-grad caller must not decide what a later one gets
    hit = self._cast.get((name, tracked))
    if hit is None:
      if self.master._version != self._low_seen:   # master was written through torch
        self.refresh_low()
      low = self._view(self.low, name)
      if tracked:
        low = _CastParam.apply(self.w[name], low, self._view(self.grad, name), self, name)
      hit = self._cast[(name, tracked)] = low
    return hit

  @torch.no_grad()
  def refresh_low(self):
    """master -> the flat compute-dtype copy (after the optimiser wrote master)."""
    if self.compute_dtype == torch.float32:
      return
    self.low_buffer().copy_(self.master)
    self._low_seen = self.master._version

  def low_buffer(self):
    if self.low is None:
      self.low = torch.zeros(self.total, dtype=self.compute_dtype, device=self.device)
    return self.low

  def low_is_fresh(self):
    """A raw-pointer writer (the fused optimiser kernel) updated master AND low."""
    self._low_seen = self.master._version

  def begin_step(self):
    self._cast.clear()

  def named_grads(self):
    return {n: self.view('grad', n) for n in self.specs}

  def state_dict(self):
    out = {n: self.view('master', n).cpu().numpy() for n in self.specs}
    out.update({n: v.cpu().numpy() for n, v in self.slow.items()})
    out['opt/nu'] = self.nu.cpu().numpy()
    out['opt/mu'] = self.mu.cpu().numpy()
    out['opt/step'] = np.asarray(self.step)
    return out

  def load_params(self, data):
    """Overwrite the named parameters only (partial restore); a slow copy follows its source."""
    for n, v in data.items():
      self.view('master', n).copy_(torch.as_tensor(v))
      if n.startswith('val/') and ('slow' + n) in self.slow:
        self.slow['slow' + n].copy_(torch.as_tensor(v))
    self._cast.clear()
    self.refresh_low()
    self.version += 1

  def load_state_dict(self, data):
    for n in self.specs:
      self.view('master', n).copy_(torch.as_tensor(data[n]))
    for n in self.slow:
      self.slow[n].copy_(torch.as_tensor(data[n]))
    if 'opt/nu' in data:
      self.nu.copy_(torch.as_tensor(data['opt/nu']))
      self.mu.copy_(torch.as_tensor(data['opt/mu']))
      self.step = int(data['opt/step'])
    self._cast.clear()
    self.refresh_low()
    self.version += 1
