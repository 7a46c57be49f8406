"""ParamStore.get in a low-precision compute dtype (embodied/jax/nets.py:243: parameters
are cast where they are used, gradients arrive in f32): views of one flat bf16 copy of
master + a backward that adds into the flat f32 gradient buffer must give exactly the
gradients of the plain formulation -- per-tensor `.to(bf16)` inside the autograd graph."""
import pytest
import torch

from oracle import dreamer_oracle as do
import dreamer_cases as cases
from embodied_b200.dreamerv3 import model as M, params as P


def _grads(plain):
  ocfg = do.tiny_config()
  vals = do.init_params(ocfg, 0, outscale_override=1.0)
  cfg = cases.product_config(ocfg)
  cfg['fused_scan'] = False
  cfg['fused_norm'] = False
  cfg['compute_dtype'] = 'bfloat16'
  store = P.ParamStore(cfg, 'cpu', torch.bfloat16, 0, {k: v.numpy() for k, v in vals.items()})
  if plain:
    def get(name):
      key = (name, torch.is_grad_enabled())
      hit = store._cast.get(key)
      if hit is None:
        hit = store._cast[key] = store.w[name].to(torch.bfloat16)
      return hit
    store.get = get
  model = M.Model(cfg, store)
  B, T = 2, 4
  data, noise = cases.batch(ocfg, B, T, seed=3), do.make_noise(ocfg, B, T, seed=4)
  carry, obs, prevact, _ = do.Dreamer(ocfg, vals).apply_replay_context(data)
  out = []
  for _ in range(2):                     # second pass: the per-step cache is rebuilt
    store.begin_step()
    store.grad.zero_()
    total, *_ = model.loss((carry['deter'], carry['stoch']), obs, prevact, noise)
    total.backward()
    out.append((float(total.detach()), store.grad.clone()))
  return store, out


def test_flat_low_precision_copy_gives_the_gradients_of_per_tensor_casts():
  _, want = _grads(plain=True)
  store, got = _grads(plain=False)
  for (lw, gw), (lg, gg) in zip(want, got):
    assert lw == lg
    assert float(gw.abs().max()) > 0
    assert torch.equal(gw, gg)
  assert store.low.dtype == torch.bfloat16 and store.low.numel() == store.master.numel()


def test_low_precision_copy_follows_master():
  store, _ = _grads(plain=False)
  name = 'dyn/obslogit/kernel'
  store.view('master', name).mul_(2.0)            # written through torch: picked up lazily
  store.begin_step()
  with torch.no_grad():
    assert torch.equal(store.get(name), store.view('master', name).to(torch.bfloat16))
  store.refresh_low()
  assert torch.equal(store.low, store.master.to(torch.bfloat16))


def test_a_no_grad_use_does_not_hide_the_parameter_from_autograd():
  store, _ = _grads(plain=False)
  name = 'dyn/obslogit/kernel'
  store.begin_step()
  store.grad.zero_()
  with torch.no_grad():
    a = store.get(name)
  b = store.get(name)
  assert not a.requires_grad and b.requires_grad
  b.float().sum().backward()
  assert float(store.view('grad', name).min()) == 1.0
