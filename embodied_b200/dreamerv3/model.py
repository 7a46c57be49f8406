"""DreamerV3 world model + actor-critic losses on the device.

Math follows dreamerv3/rssm.py and dreamerv3/agent.py of the reference (cited
per function); structure is chosen for the B200:

* everything that does not depend on the recurrent state is hoisted out of the
  T-step scan and done as one (B*T)-row GEMM: the action branch `dynin2`
  (rssm.py:145-146) and the token half of `obs0` (rssm.py:83-85);
* imagination (rssm.py:94-118) is forward-only -- `imgfeat` is stop-gradient'ed
  (agent.py:195, ac_grads False) and actions are integer samples -- so it runs
  under no_grad and keeps no activations;
* parameters are views of one flat buffer (params.py), gradients accumulate in
  place into one flat buffer that NCCL reduces with a single call;
* sampling noise is an explicit input (Gumbel tensors), generated on the device
  by default and injectable for parity tests.

Compute dtype is bfloat16 (the reference default, embodied/jax/nets.py:12) or
float32 (parity); norms, softmaxes and losses are always float32
(nets.py:372, outs.py:209,276).
"""
import math

import torch
import torch.nn.functional as F

from . import ops
from . import spaces

f32 = torch.float32


def symlog(x):
  return torch.sign(x) * torch.log1p(torch.abs(x))


def dict_concat(specs, values, squish=None):
  """nets.DictConcat (embodied/jax/nets.py:467-500) with fdims = 1: every key flattened behind its
  batch axes -- integers one-hot, floats through `squish` -- and concatenated in the (sorted) order
  of `specs`.  Unavailable entries (-inf floats, -1 integers: nets.py:80-94) contribute zeros."""
  parts = []
  for name, kind, shape, classes in specs:
    x = values[name]
    lead = x.shape[:x.dim() - len(shape)]
    if kind == 'disc':
      idx = x.long()
      ok = idx >= 0
      y = F.one_hot(torch.where(ok, idx, torch.zeros_like(idx)), classes).to(f32) * ok[..., None]
    else:
      x = x.to(f32)
      ok = x != float('-inf')
      x = torch.where(ok, x, torch.zeros_like(x))
      y = (squish(x) if squish else x) * ok
    parts.append(y.reshape(*lead, -1))
  return torch.cat(parts, -1) if len(parts) > 1 else parts[0]


def silu(x):
  return F.silu(x)


def symexp(x):
  return torch.sign(x) * torch.expm1(torch.abs(x))


_GUMBEL_CALLS = [0]


def gumbel_(u, generator=None):
  """Fill `u` (fp32) with Gumbel(0, 1) noise in place: -log(E), E ~ Exp(1) (= -log(U)), E
  clamped to what -log(clamp(U, 1e-20, 1 - 1e-7)) can be, so the noise stays inside
  [-3.9, 16.2] like the two-log formulation.  On the device: one write-only pass of
  emb_gumbel_fill (Philox keyed by the generator's seed and a per-process call counter)."""
  if u.is_cuda and u.dtype == f32 and u.is_contiguous() and u.data_ptr() % 16 == 0:
    from .. import _lib
    _GUMBEL_CALLS[0] += 1
    base = generator.initial_seed() if generator is not None else torch.initial_seed()
    seed = (base * 0x9E3779B97F4A7C15 + _GUMBEL_CALLS[0] * 0xD1B54A32D192ED03) & ((1 << 64) - 1)
    _lib.check(_lib.load().emb_gumbel_fill(
        u.data_ptr(), u.numel(), seed, torch.cuda.current_stream(u.device).cuda_stream))
    return u
  u.exponential_(1.0, generator=generator)
  u.clamp_(1e-7, 46.0)
  return u.log_().neg_()


def gumbel_like(shape, device, generator=None):
  return gumbel_(torch.empty(shape, device=device, dtype=f32), generator)


def percentiles(x, qs):
  """torch.quantile(x, qs) with linear interpolation for a flat fp32 `x`, as
  sort + two gathers: pure stream work (CUDA-graph capturable), positions
  computed on the host from the static length."""
  n = x.numel()
  s = torch.sort(x).values
  out = []
  for q in qs:
    pos = q * (n - 1)
    lo = min(int(pos), n - 1)
    hi = min(lo + 1, n - 1)
    out.append(torch.lerp(s[lo], s[hi], pos - lo))
  return torch.stack(out)


def total_fusable(losses):
  return all(v.is_cuda for v in losses.values()) and len(losses) <= 16


class Model:

  def __init__(self, cfg, store):
    self.cfg = cfg
    self.store = store
    self.cd = store.compute_dtype
    self.device = store.device
    n = cfg.bins
    half = symexp(torch.linspace(-20, 0, (n - 1) // 2 + 1, dtype=f32))
    self.bins = torch.cat([half, -half[:-1].flip(0)], 0).to(self.device)   # heads.py:132-144
    self.ret_lo = torch.zeros((), dtype=f32, device=self.device)
    self.ret_hi = torch.zeros((), dtype=f32, device=self.device)
    self.actspec = cfg.get('actspec') or [('action', 'disc', (), cfg.actions)]
    self.vecspec = cfg.get('vecspec') or []
    self.imgkeys = cfg.get('imgkeys') or ([('image', cfg.image[2])] if cfg.get('image') else [])
    # one scalar discrete action: the one-hot matmul of dynin2 is a row lookup
    self.single_disc = len(self.actspec) == 1 and self.actspec[0][1:3] == ('disc', ())
    self.fused_norm = bool(cfg.get('fused_norm', True)) and self.device.type == 'cuda'
    self.fused_spatial = bool(cfg.get('fused_spatial', True)) and self.device.type == 'cuda'
    self.tc_conv = bool(cfg.get('tc_conv', True)) and self.device.type == 'cuda'
    self.scan = None
    if cfg.get('fused_scan', True) and self.device.type == 'cuda':
      from . import scan as scanlib
      engine = scanlib.ENG_BF16 if self.cd == torch.bfloat16 else scanlib.ENG_F32
      self.scan = scanlib.Scan(cfg, store, engine)

  def _use(self, enabled, kernel, supported):
    """Gate of every own-kernel call site: True -> launch the kernel.  A call site whose kernel
    is enabled but does not take this shape / dtype runs the library formulation instead (still
    on the device) and is COUNTED in ops.FALLBACKS -- bench.py prints the table and config 2 is
    asserted to have none (tests/test_gpu_bench.py); `strict_kernels` turns it into an error."""
    if enabled and supported:
      return True
    if enabled:
      ops.note_fallback(kernel, bool(self.cfg.get('strict_kernels', False)))
    return False

  # ---------------------------------------------------------------- primitives
  def W(self, name):
    return self.store.get(name)

  def dense(self, x, name):                                  # nets.py:239-247
    w, b = self.W(f'{name}/kernel'), self.W(f'{name}/bias')
    return torch.addmm(b, x.reshape(-1, x.shape[-1]), w).reshape(*x.shape[:-1], -1)

  def block(self, x, name):                                  # nets.py:267-278
    w, b = self.W(f'{name}/kernel'), self.W(f'{name}/bias')
    g = w.shape[0]
    lead = x.shape[:-1]
    x = x.reshape(-1, g, x.shape[-1] // g).transpose(0, 1)   # (g, M, in/g)
    y = torch.bmm(x, w).transpose(0, 1)                      # (M, g, out/g)
    return y.reshape(*lead, -1) + b

  def norm(self, x, name, act=True, bias=None):              # nets.py:369-399 'rms', eps 1e-4
    """`bias`: (fp32 leaf) bias of the convolution that produced x, added here
    instead of in a separate pass over the conv output."""
    scale = self.store.w[f'{name}/scale']
    need_grad = torch.is_grad_enabled() and (x.requires_grad or scale.requires_grad)
    if self._use(self.fused_norm, 'rmsnorm_act', ops.rmsnorm_supported(x, need_grad, bias is not None)):
      return ops.rmsnorm_act(x, scale, act, bias=bias)       # one kernel each way
    if bias is not None:
      x = x + bias.to(x.dtype)
    xf = x.to(f32)
    y = xf * (torch.rsqrt(xf.square().mean(-1, keepdim=True) + 1e-4) * scale)
    y = y.to(self.cd)
    return silu(y) if act else y

  def conv(self, x, name, bias=True):                        # nets.py:298-323; x is NHWC
    """bias=False: the caller adds the bias inside the following fused norm
    (after the 2x2 max-pool, with which a per-channel constant commutes)."""
    w = self.W(f'{name}/kernel')
    if not bias and self._use(self.tc_conv, 'conv_tc', ops.conv_tc_supported(x, w.shape[2], w.shape[3], w.shape[0])):
      return ops.ConvTC.apply(x, w)                          # tcgen05 implicit GEMM (csrc/conv_tc.cu)
    w = w.permute(3, 2, 0, 1)                                # HWIO -> OIHW
    b = self.W(f'{name}/bias') if bias else None
    y = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1)                             # NHWC view (channels_last memory)

  def conv_thin_in(self, x, name):
    """SAME conv with <= 4 input channels (the image): rows of patches times the
    (k*k*Cin, Cout) kernel -- one skinny tensor-core GEMM over all pixels instead
    of a library convolution (ops.ConvPatches).  No bias (see conv())."""
    w = self.W(f'{name}/kernel')                             # HWIO
    k, _, cin, cout = w.shape
    n, h, ww, _ = x.shape
    kp = ops.patch_columns(k, cin)
    w2 = F.pad(w.reshape(k * k * cin, cout), (0, 0, 0, kp - k * k * cin))
    patches = ops.ConvPatches.apply(x, k)
    if self._use(self.tc_conv, 'thin_matmul', ops.thin_matmul_supported(patches, w2)):
      return ops.ThinMatmul.apply(patches, w2).reshape(n, h, ww, cout)
    return (patches @ w2).reshape(n, h, ww, cout)

  def conv_thin_out(self, x, name, up=1):
    """SAME conv with <= 4 output channels (the decoder's image head), optionally
    preceded by a nearest x2 up-sampling: z = x @ W[Cin, k*k*Cout] on the input
    grid, then the tap sum (+ bias) gathers the k*k shifted columns."""
    w = self.W(f'{name}/kernel')                             # HWIO
    k, _, cin, cout = w.shape
    n, h, ww, _ = x.shape
    kp = ops.patch_columns(k, cout)
    w2 = F.pad(w.permute(2, 0, 1, 3).reshape(cin, k * k * cout), (0, kp - k * k * cout))
    x2 = x.reshape(n * h * ww, cin)
    if self._use(self.tc_conv, 'thin_matmul', ops.thin_matmul_supported(x2, w2)):
      z = ops.ThinMatmul.apply(x2, w2)
    else:
      z = x2 @ w2
    return ops.ConvTapSum.apply(z, self.store.w[f'{name}/bias'], (n, h * up, ww * up, cout), k, up)

  def mlp(self, x, name, layers, params=None):               # nets.py:580-587
    for i in range(layers):
      x = self.dense(x, f'{name}/mlp/linear{i}')
      x = self.norm(x, f'{name}/mlp/norm{i}')
    return x

  # -------------------------------------------------------------------- encoder
  def encoder(self, obs, normalized=None):                   # rssm.py:210-250
    """obs: {key: tensor (..., *shape)} (a bare image tensor is accepted for the single-image
    case); `normalized`: float32 x/255-0.5 of the (concatenated) image already produced by
    emb_driver_stage_obs.  Tokens = [vector branch | image branch] (rssm.py:218-246)."""
    cfg = self.cfg
    if isinstance(obs, torch.Tensor):
      obs = {self.imgkeys[0][0]: obs}
    outs, lead = [], None
    if self.vecspec:                                          # DictConcat(symlog) -> MLP
      x = dict_concat(self.vecspec, obs, symlog)
      lead = x.shape[:-1]
      x = x.reshape(-1, x.shape[-1]).to(self.cd)
      for i in range(cfg.get('enc_layers', 3)):
        x = self.norm(self.dense(x, f'enc/mlp{i}'), f'enc/mlp{i}norm')
      outs.append(x)
    if self.imgkeys:
      if normalized is None:
        imgs = [obs[k] for k, _ in self.imgkeys]
        x = (imgs[0] if len(imgs) == 1 else torch.cat(imgs, -1)).to(f32) / 255 - 0.5
      else:
        x = normalized
      lead = x.shape[:-3]
      x = x.reshape(-1, *x.shape[-3:]).to(self.cd)
      for i in range(len(cfg.mults)):
        if x.shape[-1] <= 4 and self._use(self.fused_spatial, 'conv_patches',
                                          ops.thin_conv_supported(x, x.shape[-1], cfg.depth * cfg.mults[i])):
          x = self.conv_thin_in(x, f'enc/cnn{i}')
        else:
          x = self.conv(x, f'enc/cnn{i}', bias=False)
        x = self.pool(x)
        x = self.norm(x, f'enc/cnn{i}norm', bias=self.store.w[f'enc/cnn{i}/bias'])
      outs.append(x.reshape(x.shape[0], -1))
    x = outs[0] if len(outs) == 1 else torch.cat(outs, -1)
    return x.reshape(*lead, -1)

  # ----------------------------------------------------------------------- rssm
  def action_embed(self, action, reset):
    """DictConcat of the action dict (nets.py:467-500; a bare tensor = the single action key) with
    the reset mask of rssm.py:76-79, then `action / max(1, |action|)` (rssm.py:137; a no-op for
    one-hot columns)."""
    if isinstance(action, torch.Tensor):
      action = {self.actspec[0][0]: action}
    a = dict_concat(self.actspec, action)
    a = a * (~reset)[..., None]
    a = a / torch.clamp(a.abs(), min=1.0)
    return a.to(self.cd)

  def core(self, deter, stoch_flat, x2, out=None):           # rssm.py:135-159
    """x2 = silu(rms(dynin2(action))) is precomputed by the caller.  dynhid0's
    input is [deter_g, x0, x1, x2] per group (rssm.py:147-148); instead of
    materialising the g-fold repeat, its kernel is applied in two batched
    GEMMs: deter_g @ W[g, :Dg] + x012 @ W[g, Dg:] (x012 broadcast over g).
    `out` (no-gradient callers): a row-strided (M, D) view that receives the result."""
    g = self.cfg.blocks
    M = len(deter)
    x0 = self.norm(self.dense(deter, 'dyn/dynin0'), 'dyn/dynin0norm')
    x1 = self.norm(self.dense(stoch_flat, 'dyn/dynin1'), 'dyn/dynin1norm')
    x012 = torch.cat([x0, x1, x2], -1)
    w, b = self.W('dyn/dynhid0/kernel'), self.W('dyn/dynhid0/bias')
    Dg = deter.shape[-1] // g
    y = torch.bmm(deter.reshape(M, g, Dg).transpose(0, 1), w[:, :Dg])
    y = torch.baddbmm(y, x012[None].expand(g, -1, -1), w[:, Dg:])
    if self.fused_norm and ops.core_fused_supported(deter, g):
      # no-gradient path (imagination, policy): stay in the (group, row, column) layout of
      # the batched GEMMs; bias + norm + silu and the GRU gate chain are one kernel each
      sw = self.store.w
      x = ops.rmsnorm_grouped(y, sw['dyn/dynhid0norm/scale'], sw['dyn/dynhid0/bias'])
      pre = torch.bmm(x, self.W('dyn/dyngru/kernel'))
      return ops.gru_gates(pre, sw['dyn/dyngru/bias'], deter, out=out)
    x = y.transpose(0, 1).reshape(M, -1) + b
    x = self.norm(x, 'dyn/dynhid0norm')
    x = self.block(x, 'dyn/dyngru')
    reset, cand, update = [
        y.reshape(len(x), -1) for y in x.reshape(len(x), g, -1).chunk(3, -1)]
    reset = torch.sigmoid(reset)
    cand = torch.tanh(reset * cand)
    update = torch.sigmoid(update - 1)
    new = update * cand + (1 - update) * deter
    if out is not None:
      out.copy_(new)
      return out
    return new

  def unimix(self, logit):                                   # outs.py:210-216
    probs = torch.softmax(logit.to(f32), -1)
    return (1 - self.cfg.unimix) * probs + self.cfg.unimix / probs.shape[-1]

  def sample_stoch(self, logit, gumbel, out=None):           # outs.py:252-270
    """`out` (no-gradient callers): a row-strided (n, S*C) view that receives the sample."""
    if (self.fused_norm and not torch.is_grad_enabled() and logit.is_cuda and logit.dim() == 3
        and logit.shape[-1] <= 128 and logit.dtype in (f32, torch.bfloat16)):
      return ops.onehot_sample(logit, gumbel, self.cfg.unimix, self.cd, out=out)   # one launch
    probs = self.unimix(logit)
    index = torch.argmax(torch.log(probs) + gumbel, -1)
    value = F.one_hot(index, probs.shape[-1]).to(f32)
    value = (value + (probs - probs.detach())).to(self.cd)
    if out is not None:
      out.copy_(value.reshape(out.shape))
      return out
    return value

  def act_branch(self, action, reset):
    if isinstance(action, dict) and self.single_disc:
      action = action[self.actspec[0][0]]
    if (not torch.is_grad_enabled() and self.single_disc and isinstance(action, torch.Tensor)
        and action.dim() == 1):
      # one-hot @ kernel is a row lookup; rows of reset steps see a zero action (bias only)
      w, b = self.W('dyn/dynin2/kernel'), self.W('dyn/dynin2/bias')
      y = w[action.long()] * (~reset)[:, None].to(w.dtype) + b
      return self.norm(y, 'dyn/dynin2norm')
    a = self.action_embed(action, reset)
    return self.norm(self.dense(a, 'dyn/dynin2'), 'dyn/dynin2norm')

  def observe(self, carry, tokens, prevact, reset, gumbel):  # rssm.py:61-92
    """carry: (deter (B,D), stoch (B,S,C)); tokens (B,T,E); prevact (B,T) int;
    reset (B,T) bool; gumbel (B,T,S,C) f32.  Returns the new carry and
    deter (B,T,D), stoch (B,T,S,C), logit (B,T,S,C)."""
    cfg = self.cfg
    B, T = reset.shape
    D = cfg.deter
    x2 = self.act_branch(prevact, reset)                     # (B,T,H) hoisted
    wobs, bobs = self.W('dyn/obs0/kernel'), self.W('dyn/obs0/bias')
    tok = torch.addmm(bobs, tokens.reshape(B * T, -1), wobs[D:]).reshape(B, T, -1)
    deter, stoch = carry
    deter, stoch = deter.to(self.cd), stoch.to(self.cd)
    # T == 1 with many rows is the policy's single step over all envs: GEMM-shaped work, not the
    # 16-row weight-streaming scan.  A training scan always takes the kernel, 16 rows per launch.
    single = T == 1 and B > 16
    if not single and self._use(self.scan is not None, 'rssm_observe',
                                self.scan is not None and self.scan.supported):
      return self.observe_fused(deter, stoch, x2, tok, reset, gumbel)
    deters, stochs, logits = [], [], []
    for t in range(T):
      keep = (~reset[:, t]).to(self.cd)
      deter = deter * keep[:, None]
      stoch = stoch * keep[:, None, None]
      deter = self.core(deter, stoch.reshape(B, -1), x2[:, t])
      x = self.norm(deter @ wobs[:D] + tok[:, t], 'dyn/obs0norm')
      logit = self.dense(x, 'dyn/obslogit').reshape(B, cfg.stoch, cfg.classes)
      stoch = self.sample_stoch(logit, gumbel[:, t])
      deters.append(deter); stochs.append(stoch); logits.append(logit)
    feat = dict(deter=torch.stack(deters, 1), stoch=torch.stack(stochs, 1),
                logit=torch.stack(logits, 1))
    return (deter, stoch), feat

  def observe_fused(self, deter0, stoch0, x2, tok, reset, gumbel):
    """The T-step scan as one kernel each way (emb_rssm_observe_fwd/bwd).  Step
    0's dynin0/dynin1 pre-activations are computed here because the carry may
    be an arbitrary (not one-hot) stoch."""
    from . import scan as scanlib
    B, T = reset.shape
    keep = (~reset).to(f32)
    k0 = keep[:, 0].to(self.cd)[:, None]
    y0 = k0 * (deter0 @ self.W('dyn/dynin0/kernel')) + self.W('dyn/dynin0/bias')
    y1 = k0 * (stoch0.reshape(B, -1) @ self.W('dyn/dynin1/kernel')) + self.W('dyn/dynin1/bias')
    weights = [self.store.w[n] for n in scanlib.PARAMS]
    y0, y1, x2, tok = y0.to(f32), y1.to(f32), x2.to(f32), tok.to(f32)
    deter0 = deter0.detach()
    parts = []
    for lo in range(0, B, scanlib.ROWS):                    # the kernels walk 16 batch rows per launch:
      rows = slice(lo, lo + scanlib.ROWS)                   # larger batches stream the weights once per 16 rows
      parts.append(scanlib.ObserveFn.apply(
          self.scan, deter0[rows], y0[rows], y1[rows], x2[rows], tok[rows], keep[rows],
          gumbel[rows], *weights)[:3])
    deter, logit, stoch = parts[0] if len(parts) == 1 else [torch.cat(p, 0) for p in zip(*parts)]
    feat = dict(deter=deter, stoch=stoch, logit=logit)
    return (deter[:, -1], stoch[:, -1]), feat

  def prior(self, deter):                                    # rssm.py:161-171
    x = deter
    for i in range(self.cfg.imglayers):
      x = self.norm(self.dense(x, f'dyn/prior{i}'), f'dyn/prior{i}norm')
    x = self.dense(x, 'dyn/priorlogit')
    return x.reshape(*x.shape[:-1], self.cfg.stoch, self.cfg.classes)

  def kl_losses(self, post_logit, prior_logit):              # rssm.py:123-132, outs.py:236-240
    """dyn = max(KL(sg(post) || prior), free), rep = max(KL(post || sg(prior)), free),
    both on unimixed distributions, summed over the S latents."""
    cfg = self.cfg
    if self._use(self.fused_norm, 'rssm_kl', post_logit.dim() == 4 and ops.kl_supported(post_logit, prior_logit)):
      dyn, rep, ent_post, ent_prior = ops.rssm_kl(
          post_logit, prior_logit, cfg.unimix, cfg.free_nats)
      return dyn, rep, dict(dyn_ent=ent_prior.mean(), rep_ent=ent_post.mean())
    post = torch.log(self.unimix(post_logit))
    prior = torch.log(self.unimix(prior_logit))

    def kl(a, b):
      la, lb = torch.log_softmax(a, -1), torch.log_softmax(b, -1)
      return (torch.softmax(a, -1) * (la - lb)).sum(-1).sum(-1)

    def ent(a):
      la = torch.log_softmax(a, -1)
      return -(torch.softmax(a, -1) * la).sum(-1).sum(-1)
    dyn = torch.clamp(kl(post.detach(), prior), min=cfg.free_nats)
    rep = torch.clamp(kl(post, prior.detach()), min=cfg.free_nats)
    mets = dict(dyn_ent=ent(prior.detach()).mean(), rep_ent=ent(post.detach()).mean())
    return dyn, rep, mets

  # -------------------------------------------------------------------- decoder
  def decoder(self, deter, stoch):                           # rssm.py:288-359
    """-> {'image': sigmoid reconstruction of the (channel-concatenated) image keys (fp32),
    <vector key>: raw head output (fp32): `pred` in symlog space or categorical logits}."""
    cfg = self.cfg
    lead = deter.shape[:-1]
    recons = {}
    if self.vecspec:                                          # rssm.py:323-334
      sflat = stoch.reshape(*lead, -1).to(self.cd)
      x = torch.cat([sflat, deter.to(self.cd)], -1)           # [stoch | deter] (rssm.py:319-321)
      x = self.mlp(x.reshape(-1, x.shape[-1]), 'dec', cfg.get('dec_layers', 3))
      for spec in self.vecspec:
        name = f'dec/vec/{spec[0]}/' + ('logits' if spec[1] == 'disc' else 'pred')
        y = self.dense(x, name).to(f32)
        shape = spec[2] + ((spec[3],) if spec[1] == 'disc' else ())
        recons[spec[0]] = y.reshape(*lead, *shape)
    if not self.imgkeys:
      return recons
    depths = [cfg.depth * m for m in cfg.mults]
    minres = cfg.image[0] // 2 ** len(cfg.mults)
    g, c = cfg.bspace, depths[-1] // cfg.bspace
    x0 = deter.reshape(-1, deter.shape[-1])
    x1 = stoch.reshape(x0.shape[0], -1)
    x0 = self.block(x0, 'dec/sp0')
    x0 = x0.reshape(-1, g, minres, minres, c).permute(0, 2, 3, 1, 4).reshape(
        -1, minres, minres, g * c)                           # '(g h w c) -> h w (g c)'
    x1 = self.norm(self.dense(x1, 'dec/sp1'), 'dec/sp1norm')
    x1 = self.dense(x1, 'dec/sp2').reshape(-1, minres, minres, depths[-1])
    x = self.norm(x0 + x1, 'dec/spnorm')
    for i in reversed(range(len(depths) - 1)):
      w = self.W(f'dec/conv{i}/kernel')
      if self._use(self.tc_conv, 'upconv_subpixel', w.shape[0] == 5 and ops.subpixel_supported(x, w.shape[2], w.shape[3])):
        # up-sample + 5x5 conv on the LOW-resolution grid: 36 instead of 100 taps per input pixel
        y = ops.upconv_subpixel(x, w)
      else:
        y = self.conv(self.upsample(x), f'dec/conv{i}', bias=False)
      x = self.norm(y, f'dec/conv{i}norm', bias=self.store.w[f'dec/conv{i}/bias'])
    if self._use(self.fused_spatial, 'conv_tapsum', ops.thin_conv_supported(x, x.shape[-1], cfg.image[2])):
      x = self.conv_thin_out(x, 'dec/imgout', up=2)          # up-sampling folded in
    else:
      x = self.conv(self.upsample(x), 'dec/imgout')
    x = torch.sigmoid(x.to(f32))
    recons['image'] = x.reshape(*lead, *x.shape[1:])
    return recons

  def recon_losses(self, recons, obs):
    """One (B, T) loss per decoded key (agent.py:176-180): images MSE against x / 255 summed over
    H, W, C (rssm.py:354-357, split per image key); float vectors MSE in symlog space, integer
    vectors categorical cross-entropy, both summed over the key's own axes (heads.py:87-88)."""
    out = {}
    if self.imgkeys:
      c0 = 0
      for key, ch in self.imgkeys:
        target = obs[key].to(f32) / 255
        out[key] = (recons['image'][..., c0: c0 + ch] - target).square().sum((-3, -2, -1))
        c0 += ch
    for name, kind, shape, classes in self.vecspec:
      y, dims = recons[name], tuple(range(-len(shape), 0)) if shape else ()
      if kind == 'disc':
        logp = torch.log_softmax(y, -1).gather(-1, obs[name].long()[..., None]).squeeze(-1)
        loss = -logp
      else:
        loss = (y - symlog(obs[name].to(f32))).square()
      out[name] = loss.sum(dims) if dims else loss
    return out

  def upsample(self, x):                                     # x.repeat(2,-2).repeat(2,-3), NHWC
    if self._use(self.fused_spatial, 'upsample2', ops.spatial_supported(x)):
      return ops.Upsample2.apply(x)
    y = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode='nearest')
    return y.permute(0, 2, 3, 1)

  def pool(self, x):                                         # rssm.py:239-240, NHWC
    if self._use(self.fused_spatial, 'maxpool2', ops.spatial_supported(x)):
      return ops.MaxPool2.apply(x)
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)

  # ---------------------------------------------------------------------- heads
  def feat2tensor(self, deter, stoch):                       # agent.py:51-53
    return torch.cat([deter.to(self.cd), stoch.reshape(*stoch.shape[:-2], -1).to(self.cd)], -1)

  def head(self, x, name, layers, out):                      # heads.py:34-41
    return self.dense(self.mlp(x, name, layers), f'{name}/head/{out}').to(f32)

  # -- policy: one distribution per action key (agent.py:61-64, heads.py:103-155) ---------
  def policy_outputs(self, x):
    """{key: logits (..., *shape, classes)} for discrete keys, {key: (mean, std)} for continuous
    ones: bounded_normal = Normal(tanh(mean), (maxstd - minstd) * sigmoid(std + 2) + minstd)."""
    cfg = self.cfg
    h = self.mlp(x, 'pol', cfg.pol_layers)
    outs = {}
    for name, kind, shape, classes in self.actspec:
      if kind == 'disc':
        y = self.dense(h, f'pol/head/{name}/logits').to(f32)
        outs[name] = y.reshape(*y.shape[:-1], *shape, classes)
      else:
        mean = self.dense(h, f'pol/head/{name}/mean').to(f32)
        std = self.dense(h, f'pol/head/{name}/stddev').to(f32)
        lo, hi = cfg.get('minstd', 0.1), cfg.get('maxstd', 1.0)
        std = (hi - lo) * torch.sigmoid(std + 2.0) + lo
        outs[name] = (torch.tanh(mean).reshape(*mean.shape[:-1], *shape),
                      std.reshape(*std.shape[:-1], *shape))
    return outs

  def policy_sample(self, outs, noise):
    """Categorical: argmax(logits + Gumbel) (jax.random.categorical); Normal: mean + std * eps."""
    acts = {}
    for name, kind, shape, classes in self.actspec:
      if kind == 'disc':
        acts[name] = torch.argmax(outs[name] + noise[name], -1).to(torch.int32)
      else:
        mean, std = outs[name]
        acts[name] = mean + std * noise[name]
    return acts

  def policy_logp_entropy(self, outs, acts):
    """sum_k logp_k(act_k), sum_k entropy_k (agent.py:407-408), each summed over the key's own
    axes (outs.Agg).  Normal: outs.py:160-167."""
    logp, ent = 0.0, 0.0
    for name, kind, shape, classes in self.actspec:
      dims = tuple(range(-len(shape), 0)) if shape else ()
      if kind == 'disc':
        la = torch.log_softmax(outs[name], -1)
        lp = la.gather(-1, acts[name].long()[..., None]).squeeze(-1)
        en = -(torch.softmax(outs[name], -1) * la).sum(-1)
      else:
        mean, std = outs[name]
        a = acts[name].to(f32)
        lp = -0.5 * ((a - mean) / std).square() - torch.log(std) - 0.5 * math.log(2 * math.pi)
        en = 0.5 * torch.log(2 * math.pi * std.square()) + 0.5
      logp = logp + (lp.sum(dims) if dims else lp)
      ent = ent + (en.sum(dims) if dims else en)
    return logp, ent

  def action_noise(self, lead, generator=None):
    """{key: Gumbel noise (*lead, *shape, classes) | normal noise (*lead, *shape)}."""
    out = {}
    for name, kind, shape, classes in self.actspec:
      if kind == 'disc':
        out[name] = gumbel_like((*lead, *shape, classes), self.device, generator)
      else:
        out[name] = torch.randn((*lead, *shape), device=self.device, generator=generator)
    return out

  def slow_value_logits(self, x):
    """The slow critic (utils.py:94-127): val's architecture, slow parameters."""
    slow = self.store.slow
    cache = self.store._cast           # cleared at every step boundary (ParamStore.begin_step)

    def low(name):                     # the compute-dtype copy, made once per update (two call sites)
      hit = cache.get(('slow', name))
      if hit is None:
        hit = cache[('slow', name)] = slow[name].to(self.cd)
      return hit
    with torch.no_grad():
      for i in range(self.cfg.val_layers):
        w, b = low(f'slowval/mlp/linear{i}/kernel'), low(f'slowval/mlp/linear{i}/bias')
        x = torch.addmm(b, x.reshape(-1, x.shape[-1]), w).reshape(*x.shape[:-1], -1)
        scale = slow[f'slowval/mlp/norm{i}/scale']
        if self.fused_norm and ops.rmsnorm_supported(x, False):
          x = ops.rmsnorm_act(x, scale, True)                # one kernel instead of eight
        else:
          xf = x.to(f32)
          x = silu((xf * (torch.rsqrt(xf.square().mean(-1, keepdim=True) + 1e-4) * scale)).to(self.cd))
      w, b = low('slowval/head/logits/kernel'), low('slowval/head/logits/bias')
      return torch.addmm(b, x.reshape(-1, x.shape[-1]), w).reshape(
          *x.shape[:-1], -1).to(f32)

  def twohot_pred(self, logits):                             # outs.py:285-302 (symmetric sum)
    if not logits.requires_grad and self._use(self.fused_norm, 'twohot_pred', ops.twohot_supported(logits)):
      return ops.twohot_pred(logits, self.bins)              # one launch
    probs = torch.softmax(logits, -1)
    bins = self.bins
    m = (logits.shape[-1] - 1) // 2
    p1, p2, p3 = probs[..., :m], probs[..., m: m + 1], probs[..., m + 1:]
    b1, b2, b3 = bins[:m], bins[m: m + 1], bins[m + 1:]
    return (p2 * b2).sum(-1) + ((p1 * b1).flip(-1) + (p3 * b3)).sum(-1)

  def twohot_loss(self, logits, target, target2=None, w2=0.0):   # outs.py:311-330
    """CE against twohot(target) (+ w2 * CE against twohot(target2): the critic's two terms)."""
    if self._use(self.fused_norm, 'twohot_loss', ops.twohot_supported(logits)):
      return ops.twohot_loss(logits, self.bins, target, target2, w2)     # one launch each way
    if target2 is not None:
      return self.twohot_loss(logits, target) + w2 * self.twohot_loss(logits, target2)
    bins, n = self.bins, len(self.bins)
    target = target.detach().to(f32)
    below = (bins <= target[..., None]).sum(-1) - 1
    above = n - (bins > target[..., None]).sum(-1)
    below = below.clamp(0, n - 1)
    above = above.clamp(0, n - 1)
    equal = below == above
    one = torch.ones_like(target)
    to_below = torch.where(equal, one, (bins[below] - target).abs())
    to_above = torch.where(equal, one, (bins[above] - target).abs())
    total = to_below + to_above
    logp = logits - torch.logsumexp(logits, -1, keepdim=True)
    lb = logp.gather(-1, below[..., None]).squeeze(-1)
    la = logp.gather(-1, above[..., None]).squeeze(-1)
    return -(lb * (to_above / total) + la * (to_below / total))

  @staticmethod
  def lambda_return(last, term, rew, val, boot, disc, lam):  # agent.py:482-490
    if rew.is_cuda and rew.dim() == 2:                        # one launch; feeds detached targets only
      return ops.lambda_return(last, term, rew, boot, disc, lam)
    live = (1 - term.to(f32))[:, 1:] * disc
    cont = (1 - last.to(f32))[:, 1:] * lam
    interm = rew[:, 1:] + (1 - cont) * live * boot[:, 1:]
    rets = [boot[:, -1]]
    for t in reversed(range(live.shape[1])):
      rets.append(interm[:, t] + live[:, t] * cont[:, t] * rets[-1])
    return torch.stack(list(reversed(rets))[:-1], 1)

  def retnorm(self, ret, update):                            # utils.py:37-77 'perc', debias False
    cfg = self.cfg
    if update:
      x = ret.detach().to(f32).flatten()
      q = percentiles(self.gather_returns(x), [cfg.perclo / 100, cfg.perchi / 100])
      r = cfg.retnorm_rate
      # in place: the EMA state keeps its address across CUDA-graph replays
      self.ret_lo.mul_(1 - r).add_(q[0], alpha=r)
      self.ret_hi.mul_(1 - r).add_(q[1], alpha=r)
    lo, hi = self.ret_lo.clone(), self.ret_hi.clone()
    return lo, torch.clamp(hi - lo, min=cfg.retnorm_limit)

  def gather_returns(self, x):
    """utils.py:83-88: with data-parallel ranks the returns of ALL ranks are
    gathered before the percentile, so every rank normalises identically."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
      return x
    parts = torch.empty(dist.get_world_size() * x.numel(), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(parts, x.contiguous())
    return parts

  # ----------------------------------------------------------------- imagination
  @torch.no_grad()
  def imagine(self, deter, stoch, noise_stoch, noise_act):   # rssm.py:94-118, agent.py:188-200
    """From B*K start states roll H steps with the policy in the loop.  Returns
    the features (BK, H+1, D + S*C) -- deter | flat stoch, written in place step
    by step -- and the actions {key: (BK, H+1, *shape)}.  noise_act: {key: (BK, H+1, ...)}."""
    cfg = self.cfg
    H, D = cfg.imag_length, cfg.deter
    n = len(deter)
    never = torch.zeros(n, dtype=torch.bool, device=deter.device)
    feat = torch.empty((n, H + 1, D + cfg.stoch * cfg.classes), dtype=self.cd, device=deter.device)
    feat[:, 0, :D] = deter
    feat[:, 0, D:] = stoch.reshape(n, -1)
    acts = {spec[0]: [] for spec in self.actspec}
    for h in range(H + 1):
      cur = feat[:, h]                  # (n, D + S*C) rows of the buffer: no per-step copies,
      a = self.policy_sample(self.policy_outputs(cur), {k: v[:, h] for k, v in noise_act.items()})
      for k, v in a.items():
        acts[k].append(v)
      if h == H:
        break
      x2 = self.act_branch(a, never)
      nxt = feat[:, h + 1]              # the next state is written where it is kept
      self.core(cur[:, :D], cur[:, D:], x2, out=nxt[:, :D])
      self.sample_stoch(self.prior(nxt[:, :D]), noise_stoch[:, h], out=nxt[:, D:])
    return feat, {k: torch.stack(v, 1) for k, v in acts.items()}

  # ------------------------------------------------------------------------ loss
  def loss(self, carry, obs, prevact, noise, update=True):   # agent.py:156-245
    cfg = self.cfg
    reset = obs['is_first']
    B, T = reset.shape
    losses, metrics = {}, {}
    tokens = self.encoder(obs)
    carry, feat = self.observe(carry, tokens, prevact, reset, noise['observe'])
    dyn, rep, mets = self.kl_losses(feat['logit'], self.prior(feat['deter']))
    losses['dyn'], losses['rep'] = dyn, rep
    metrics.update(mets)
    recon = self.decoder(feat['deter'], feat['stoch'])
    inp = self.feat2tensor(feat['deter'], feat['stoch'])
    losses['rew'] = self.twohot_loss(
        self.head(inp, 'rew', cfg.rew_layers, 'logits'), obs['reward'])
    con = (~obs['is_terminal']).to(f32)
    if cfg.contdisc:
      con = con * (1 - 1 / cfg.horizon)
    clogit = self.head(inp, 'con', cfg.con_layers, 'logit').squeeze(-1)
    losses['con'] = -(con * F.logsigmoid(clogit) + (1 - con) * F.logsigmoid(-clogit))
    losses.update(self.recon_losses(recon, obs))

    K, H = T, cfg.imag_length
    noise_act = noise['imag_act']
    if not isinstance(noise_act, dict):
      noise_act = {self.actspec[0][0]: noise_act}
    imgfeat, imgact = self.imagine(
        feat['deter'].detach().reshape(B * K, -1),
        feat['stoch'].detach().reshape(B * K, cfg.stoch, cfg.classes),
        noise['imag_stoch'], noise_act)
    imgdeter = imgfeat[..., :cfg.deter]
    imgstoch = imgfeat[..., cfg.deter:].reshape(B * K, H + 1, cfg.stoch, cfg.classes)
    los, ret, mets = self.imag_loss(imgact, imgfeat, update)
    losses.update({k: v.mean(1).reshape(B, K) for k, v in los.items()})
    metrics.update(mets)

    # replay value loss (agent.py:219-235, repl_loss :449-479)
    boot = ret[:, 0].reshape(B, K)
    vlogits = self.head(inp, 'val', cfg.val_layers, 'logits')
    val = self.twohot_pred(vlogits.detach())       # only feeds stop-gradient targets (agent.py:462-470)
    slow = self.twohot_pred(self.slow_value_logits(inp.detach()))
    rret = self.lambda_return(
        obs['is_last'], obs['is_terminal'], obs['reward'].to(f32), val, boot,
        1 - 1 / cfg.horizon, cfg.lam)
    padded = torch.cat([rret, 0 * rret[:, -1:]], 1)
    weight = (~obs['is_last']).to(f32)
    losses['repval'] = weight[:, :-1] * self.twohot_loss(vlogits, padded, slow, cfg.slowreg)[:, :-1]

    scales = dict(cfg.scales)
    rec = scales.pop('image')                                # agent.py:75-78: `rec` for every decoded key
    scales.update({k: rec for k, _ in self.imgkeys})
    scales.update({spec[0]: rec for spec in self.vecspec})
    assert set(losses) == set(scales), (sorted(losses), sorted(scales))
    if self._use(self.fused_norm, 'loss_reduce', total_fusable(losses)):
      total, means = ops.loss_sum(losses, scales)            # one launch (agent.py:237-240)
      metrics.update({f'loss/{k}': v for k, v in means.items()})
    else:
      metrics.update({f'loss/{k}': v.detach().mean() for k, v in losses.items()})
      total = sum(v.mean() * scales[k] for k, v in losses.items())
    outs = dict(tokens=tokens, feat=feat, losses=losses, recon=recon,
                imgdeter=imgdeter, imgstoch=imgstoch, ret=ret,
                imgact=imgact[self.actspec[0][0]] if len(self.actspec) == 1 else imgact)
    return total, carry, outs, metrics

  def imag_loss(self, act, inp, update):                     # agent.py:382-446
    cfg = self.cfg
    with torch.no_grad():
      rew = self.twohot_pred(self.head(inp, 'rew', cfg.rew_layers, 'logits'))
      con = torch.sigmoid(self.head(inp, 'con', cfg.con_layers, 'logit').squeeze(-1))
      slowval = self.twohot_pred(self.slow_value_logits(inp))
    pol = self.policy_outputs(inp)
    vlogits = self.head(inp, 'val', cfg.val_layers, 'logits')
    val = self.twohot_pred(vlogits.detach())
    disc = 1 if cfg.contdisc else 1 - 1 / cfg.horizon
    weight = torch.cumprod(disc * con, 1) / disc
    ret = self.lambda_return(torch.zeros_like(con), 1 - con, rew, val, val, disc, cfg.lam)
    roffset, rscale = self.retnorm(ret, update)
    adv = (ret - val[:, :-1]) / rscale
    logpi, ent = self.policy_logp_entropy(pol, act)
    logpi, ent = logpi[:, :-1], ent[:, :-1]
    losses = {}
    losses['policy'] = weight[:, :-1] * -(logpi * adv + cfg.actent * ent)
    padded = torch.cat([ret, 0 * ret[:, -1:]], 1)
    losses['value'] = weight[:, :-1] * self.twohot_loss(vlogits, padded, slowval, cfg.slowreg)[:, :-1]
    ret_normed = (ret - roffset) / rscale
    mets = dict(adv=adv.mean(), rew=rew.mean(), con=con.mean(), ret=ret_normed.mean(),
                val=val.mean(), weight=weight.mean(), ent=ent.detach().mean())
    return losses, ret, mets
