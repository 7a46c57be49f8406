"""The reference's optimiser chain on the flat parameter buffer
(dreamerv3/agent.py:342-379; embodied/jax/opt.py:109-164):

  clip_by_agc(0.3)      per tensor: g *= 1 / max(1, |g| / (clip * max(pmin, |w|)))
  scale_by_rms(b2, eps) nu = b2 nu + (1-b2) g^2 ; g /= sqrt(nu / (1-b2^t)) + eps
  scale_by_momentum(b1) mu = b1 mu + (1-b1) g   ; g  = mu / (1-b1^t)
  learning rate         linear warm-up 0 -> lr over `warmup` steps, then const;
                        optax evaluates the schedule at the count BEFORE the
                        increment, so the very first update has lr = 0
  w -= lr * g

and SlowModel.update (embodied/jax/utils.py:113-119): slow = r*val + (1-r)*slow.
"""
import ctypes

import numpy as np
import torch

CHUNK = 4096


class Optimizer:

  def __init__(self, cfg, store):
    self.cfg = cfg
    self.store = store
    names = list(store.specs)
    self.fused = bool(cfg.get('fused_opt', True)) and store.device.type == 'cuda'
    if self.fused:
      self._init_fused(names)
    self._grads = [store.view('grad', n) for n in names]
    self._params = [store.view('master', n) for n in names]
    self._slow_pairs = [
        (store.slow[n.replace('val/', 'slowval/', 1)], store.view('master', n))
        for n in names if n.startswith('val/')]

  def learning_rate(self, count):
    cfg = self.cfg
    if cfg.warmup and count < cfg.warmup:
      return cfg.lr * count / cfg.warmup
    return cfg.lr

  def _init_fused(self, names):
    from .. import _lib
    st = self.store
    self._lib = _lib
    self.lib = _lib.load()
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    self.lib.emb_opt_agc_rms_momentum_cast.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, vp]
    self.lib.emb_opt_agc_rms_momentum_cast.restype = ctypes.c_int
    rows = []
    for ti, n in enumerate(names):
      off, size = st.offsets[n], int(np.prod(st.specs[n][0]))
      for b in range(0, size, CHUNK):
        rows.append((off + b, min(CHUNK, size - b), ti))
    table = np.zeros(len(rows), dtype=[('begin', '<i8'), ('count', '<i4'), ('tensor', '<i4')])
    table['begin'], table['count'], table['tensor'] = zip(*rows)
    self.nchunks, self.ntensors = len(rows), len(names)
    self.chunks = torch.from_numpy(table.view(np.uint8).copy()).to(st.device)
    self.norms = torch.zeros(2 * len(names), dtype=torch.float32, device=st.device)
    # deterministic per-tensor norms: per-chunk partial sums + the first chunk of every tensor
    self.partials = torch.zeros(2 * len(rows), dtype=torch.float32, device=st.device)
    first = np.searchsorted(table['tensor'], np.arange(len(names) + 1)).astype(np.int32)
    self.tensor_first = torch.from_numpy(first).to(st.device)
    cfg = self.cfg
    self.hyper = torch.tensor(
        [0, 0, 0, cfg.beta1, cfg.beta2, cfg.eps, cfg.agc, cfg.pmin],
        dtype=torch.float32, device=st.device)
    self.count = torch.full((), float(st.step), dtype=torch.float64, device=st.device)

  @torch.no_grad()
  def device_hyper(self):
    """Step-dependent scalars from the DEVICE update counter (float64), so the
    whole update is stream work and can live inside a CUDA graph.  optax
    evaluates the schedule at the count before the increment."""
    cfg = self.cfg
    count = self.count
    t = count + 1
    if cfg.warmup:
      lr = cfg.lr * torch.clamp(count / cfg.warmup, max=1.0)
    else:
      lr = torch.full_like(count, cfg.lr)
    h = torch.stack([lr, 1 / (1 - cfg.beta1 ** t), 1 / (1 - cfg.beta2 ** t)])
    self.hyper[:3].copy_(h)
    self.count.add_(1)

  def begin_update(self):
    """Host-side bookkeeping of one update; returns the scalar metrics."""
    count = self.store.step
    self._lr, self._t = self.learning_rate(count), count + 1
    return {'opt/updates': self._t, 'opt/lr': self._lr}

  def end_update(self):
    self.store.step += 1
    self.store.version += 1

  @torch.no_grad()
  def launch(self):
    """The optimiser chain on the current stream; returns the gradient norm.
    Fused: hyper-parameter scalars + the two kernels (graph-capturable)."""
    cfg, st = self.cfg, self.store
    if self.fused:
      self.device_hyper()
      stream = torch.cuda.current_stream(st.device).cuda_stream
      # the kernel also writes the bf16 copy of the new parameters (ParamStore.get)
      low = st.low_buffer() if st.compute_dtype == torch.bfloat16 else None
      self._lib.check(self.lib.emb_opt_agc_rms_momentum_cast(
          st.grad.data_ptr(), st.master.data_ptr(), st.nu.data_ptr(), st.mu.data_ptr(),
          None if low is None else low.data_ptr(),
          self.chunks.data_ptr(), self.nchunks, self.norms.data_ptr(), self.ntensors,
          self.hyper.data_ptr(), self.partials.data_ptr(), self.tensor_first.data_ptr(), stream))
      if low is not None:
        st.low_is_fresh()
      return self.norms[0::2].sum().sqrt()
    t, lr = self._t, self._lr
    gn = torch.stack(torch._foreach_norm(self._grads))
    pn = torch.stack(torch._foreach_norm(self._params))
    upper = cfg.agc * torch.clamp(pn, min=cfg.pmin)
    scale = 1 / torch.clamp(gn / upper, min=1.0)
    torch._foreach_mul_(self._grads, list(scale.unbind()))
    g = st.grad
    st.nu.mul_(cfg.beta2).addcmul_(g, g, value=1 - cfg.beta2)
    u = g / ((st.nu / (1 - cfg.beta2 ** t)).sqrt_() + cfg.eps)
    st.mu.mul_(cfg.beta1).add_(u, alpha=1 - cfg.beta1)
    st.master.add_(st.mu, alpha=-lr / (1 - cfg.beta1 ** t))
    st.refresh_low()
    return torch.linalg.vector_norm(gn)

  @torch.no_grad()
  def step(self):
    """One eager update (begin_update + launch + end_update)."""
    mets = self.begin_update()
    mets['opt/grad_norm'] = self.launch()
    self.end_update()
    return mets

  def sync_count(self):
    """After loading a checkpoint: device counter <- host counter."""
    if self.fused:
      self.count.fill_(float(self.store.step))

  @torch.no_grad()
  def update_slow(self):
    r = self.cfg.slowrate
    torch._foreach_lerp_([s for s, _ in self._slow_pairs], [v for _, v in self._slow_pairs], r)
