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
import torch


class Optimizer:

  def __init__(self, cfg, store):
    self.cfg = cfg
    self.store = store
    names = list(store.specs)
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

  @torch.no_grad()
  def step(self):
    cfg, st = self.cfg, self.store
    count = st.step
    t = count + 1
    lr = self.learning_rate(count)
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
    st.step = t
    st.version += 1
    return {'opt/grad_norm': torch.linalg.vector_norm(gn), 'opt/updates': t,
            'opt/lr': lr}

  @torch.no_grad()
  def update_slow(self):
    r = self.cfg.slowrate
    for slow, src in self._slow_pairs:
      slow.mul_(1 - r).add_(src, alpha=r)
