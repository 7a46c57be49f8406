"""Self-checks of the ppo oracle restatement (parity unpinned by the reference: no reference test
touches ppo numerics and JAX is absent).  Structural facts of ppo/agent.py and optax, and the
committed golden of the oracle's own numbers (oracle drift shows up as a diff)."""
import math
import pathlib

import numpy as np
import torch

from oracle import ppo_oracle as po
import ppo_cases as cases

GOLDEN = pathlib.Path(__file__).parent / 'golden' / 'ppo_tiny.npz'


def _tiny(spaces=cases.dummy_spaces, **over):
  obs, act = spaces()
  cfg = po.tiny_config(**over)
  model, vals = cases.oracle_for(cfg, obs, act)
  return cfg, obs, act, model


def test_advantage_closed_form():
  """No episode ends, constant reward r and value v: adv_t = sum_k (keep*lam*keep)^k * delta with
  delta = r + keep*v - v (ppo/agent.py:204-212)."""
  cfg, obs, act, m = _tiny()
  B, T = 2, 9
  data = cases.batch(cfg, obs, act, B, T - 1, flags=False)
  data = {k: v[:, :T] for k, v in data.items()}
  data['reward'] = torch.full((B, T), 0.5)
  value = torch.full((B, T), 2.0)
  pol = {'act_disc': torch.zeros(B, T, 5), 'act_cont': (torch.zeros(B, T, 6), torch.ones(B, T, 6))}
  losses, mets = m.ppo_loss(data, pol, value, update=False)
  keep = 1 - 1 / cfg.hor
  # before any update the value normaliser is (0, max(limit, 0)): val = value * limit
  v = 2.0 * cfg.norm_limit
  delta = 0.5 + keep * v - v
  expect = [sum((keep * cfg.lam) ** k * delta for k in range(T - 1 - t)) for t in range(T - 1)]
  np.testing.assert_allclose(float(mets['adv']), np.mean(expect), rtol=1e-5)
  assert tuple(losses['value'].shape) == (B, T) and tuple(losses['policy'].shape) == (B, T - 1)


def test_first_update_has_zero_learning_rate_and_counts():
  cfg, obs, act, m = _tiny()
  before = {k: v.clone() for k, v in m.p.items()}
  data = cases.batch(cfg, obs, act, 2, 6)
  carry = (m.initial(2), {k: torch.zeros(2, *v.shape) for k, v in act.items()})
  _, _, mets, grads, _ = m.train(carry, data)
  assert mets['opt/updates'] == 1
  assert all(torch.equal(before[k], m.p[k]) for k in before)          # linear_schedule(0, lr, warmup)(0) == 0
  assert any(float(g.abs().max()) > 0 for g in grads.values())
  _, _, mets, _, _ = m.train(carry, data)
  assert not all(torch.equal(before[k], m.p[k]) for k in before)


def test_clip_by_global_norm_and_adam_unit_step():
  """With eps -> 0 the first Adam step is sign(g) (m_hat / sqrt(v_hat) = g / |g|)."""
  cfg, obs, act, m = _tiny(warmup=0, eps=0.0, lr=1e-2, clip=1e-3)
  before = {k: v.clone() for k, v in m.p.items()}
  grads = {k: torch.randn_like(v) for k, v in m.p.items()}
  mets = m.apply_updates(grads)
  gnorm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
  np.testing.assert_allclose(mets['opt/grad_norm'], gnorm, rtol=1e-6)
  for k in before:
    np.testing.assert_allclose((before[k] - m.p[k]).numpy(), 1e-2 * torch.sign(grads[k]).numpy(), rtol=1e-4, atol=1e-7)


def test_normalizer_debias():
  """After one update the debiased mean equals the batch mean exactly (utils.py:60-75)."""
  cfg, obs, act, m = _tiny()
  x = torch.randn(4, 7) * 3 + 2
  m.norm_update('valnorm', x)
  mean, std = m.norm_stats('valnorm')
  np.testing.assert_allclose(mean, float(x.mean()), rtol=1e-5)
  np.testing.assert_allclose(std, float(x.std(unbiased=False)), rtol=1e-4)


def test_policy_outputs_and_memory_reset():
  cfg, obs, act, m = _tiny()
  g = torch.Generator().manual_seed(3)
  o = cases.obs_batch(obs, (3,), g)
  o['is_first'][:] = True
  carry = (torch.randn(3, cfg.rnn_units), cases.act_batch(act, (3,), g))
  noise = po.make_noise(act, (3,), 0)
  (mem, prev), acts, out = m.policy(carry, o, noise)
  carry2 = (torch.zeros(3, cfg.rnn_units), {k: torch.zeros_like(v) for k, v in carry[1].items()})
  (mem2, _), acts2, out2 = m.policy(carry2, o, noise)
  assert torch.equal(mem, mem2)                                       # is_first wipes memory and prevact
  assert set(out) == {'logp/act_disc', 'logp/act_cont', 'memory'}
  assert acts['act_disc'].dtype == torch.int32 and tuple(acts['act_cont'].shape) == (3, 6)
  assert float(out['logp/act_disc'].max()) <= 0


def test_golden_oracle_numbers():
  import sys
  sys.path.insert(0, str(pathlib.Path(__file__).parent.parent / 'oracle'))
  from oracle import gen_ppo_golden as gen
  got = gen.run()
  want = np.load(GOLDEN)
  assert set(got) == set(want.files)
  for k in want.files:
    np.testing.assert_allclose(got[k], want[k], rtol=2e-5, atol=1e-6, err_msg=k)
