"""Self-checks of the dreamerv3 oracle restatement (parity is unpinned by the
reference: SURVEY 8c).  Structural facts the reference's own code asserts or
documents: all world-model losses shaped (B, T) (dreamerv3/agent.py:184-186),
loss keys == scale keys (:237-238), TwoHot.pred()==0 at init by the symmetric
sum (embodied/jax/outs.py:286-290 with outscale 0.0 heads), KL >= 0 and the
free-nats floor (rssm.py:127-129), first optimiser step has lr 0 (warm-up)."""
import numpy as np
import torch

from oracle import dreamer_oracle as do
import dreamer_cases as cases


def test_losses_shapes_and_floors():
  cfg = do.tiny_config()
  m = do.Dreamer(cfg, do.init_params(cfg, 0))
  B, T = 2, 6
  data = cases.batch(cfg, B, T)
  carry, obs, prevact, stepid = m.apply_replay_context(data)
  total, carry, entries, outs, mets = m.loss(carry, obs, prevact, do.make_noise(cfg, B, T, 0))
  for k in ('dyn', 'rep', 'rew', 'con', 'image', 'policy', 'value'):
    assert tuple(outs['losses'][k].shape) == (B, T), k
  assert tuple(outs['losses']['repval'].shape) == (B, T - 1)
  assert float(outs['losses']['dyn'].min()) >= cfg.free_nats
  assert float(outs['losses']['rep'].min()) >= cfg.free_nats
  assert tuple(stepid.shape) == (B, T, 20)
  assert torch.isfinite(total)


def test_zero_init_heads_predict_zero():
  cfg = do.tiny_config()
  m = do.Dreamer(cfg, do.init_params(cfg, 0))
  x = torch.randn(7, cfg.deter + cfg.stoch * cfg.classes)
  assert float(do.twohot_pred(m.rew_logits(x), m.bins).abs().max()) == 0.0
  assert float(do.twohot_pred(m.val_logits(x), m.bins).abs().max()) == 0.0
  np.testing.assert_allclose(
      do.twohot_loss(m.rew_logits(x), m.bins, torch.zeros(7)).numpy(),
      np.log(cfg.bins), rtol=1e-6)


def test_twohot_loss_is_expectation_preserving():
  bins = do.twohot_bins(255)
  target = torch.tensor([-3.3, 0.0, 0.2, 17.0, 1e9, -1e9])
  logits = torch.zeros(6, 255, requires_grad=True)
  loss = do.twohot_loss(logits, bins, target).sum()
  grad, = torch.autograd.grad(loss, logits)
  tgt = torch.softmax(logits, -1) - grad              # d/dlogits CE = p - target
  expect = (tgt * bins).sum(-1)
  clipped = target.clamp(bins[0], bins[-1])
  np.testing.assert_allclose(expect.detach().numpy(), clipped.numpy(), rtol=1e-5, atol=1e-5)


def test_kl_nonnegative_and_zero_on_equal():
  a = torch.randn(3, 5, 8, 4)
  assert float(do.cat_kl(a, a).abs().max()) < 1e-6
  assert float(do.cat_kl(a, torch.randn(3, 5, 8, 4)).min()) >= 0


def test_gate_identities():
  """update gate sigmoid(u - 1) (rssm.py:157): with all-zero dyngru the new
  deter is (1 - sigmoid(-1)) * deter."""
  cfg = do.tiny_config()
  p = do.init_params(cfg, 0)
  p['dyn/dyngru/kernel'].zero_()
  m = do.Dreamer(cfg, p)
  deter = torch.randn(3, cfg.deter)
  stoch = torch.zeros(3, cfg.stoch, cfg.classes)
  out = m.core(deter, stoch, torch.zeros(3, cfg.actions))
  np.testing.assert_allclose(
      out.numpy(), ((1 - torch.sigmoid(torch.tensor(-1.0))) * deter).numpy(), rtol=1e-6, atol=1e-7)


def test_first_update_has_zero_learning_rate():
  cfg = do.tiny_config()
  p = do.init_params(cfg, 0, outscale_override=1.0)
  m = do.Dreamer(cfg, {k: v.clone() for k, v in p.items()})
  B, T = 2, 5
  data, noise = cases.batch(cfg, B, T), do.make_noise(cfg, B, T, 0)
  m.train(data, noise)
  assert all(torch.equal(m.p[k], p[k]) for k in p)
  m.train(data, noise)
  assert any(not torch.equal(m.p[k], p[k]) for k in p)
  assert m.state['step'] == 2


def test_lambda_return_matches_recursion():
  B, T = 3, 7
  g = torch.Generator().manual_seed(0)
  rew, val = torch.randn(B, T, generator=g), torch.randn(B, T, generator=g)
  last = torch.rand(B, T, generator=g) < 0.2
  term = last & (torch.rand(B, T, generator=g) < 0.5)
  ret = do.lambda_return(last, term, rew, val, val, 0.99, 0.95)
  assert tuple(ret.shape) == (B, T - 1)
  want = torch.zeros(B, T)
  want[:, -1] = val[:, -1]
  for t in reversed(range(T - 1)):
    live = (1 - term[:, t + 1].float()) * 0.99
    cont = (1 - last[:, t + 1].float()) * 0.95
    want[:, t] = rew[:, t + 1] + live * ((1 - cont) * val[:, t + 1] + cont * want[:, t + 1])
  np.testing.assert_allclose(ret.numpy(), want[:, :-1].numpy(), rtol=1e-5, atol=1e-6)


def test_oracle_reproduces_its_committed_golden():
  """tests/golden/dreamer_tiny.npz (oracle/gen_dreamer_golden.py): an edit that changes the
  oracle's numbers must be deliberate.  fp32 CPU arithmetic differs slightly between BLAS
  builds / thread counts, hence 1e-5 on floats; sampled indices and actions are exact."""
  import numpy as np
  import pathlib
  from oracle import gen_dreamer_golden as gg
  want = np.load(pathlib.Path(__file__).parent / 'golden' / 'dreamer_tiny.npz')
  got = gg.run()
  assert sorted(got) == sorted(want.files)
  for k in want.files:
    a, b = np.asarray(got[k]), want[k]
    if b.dtype == np.int8:
      assert (a == b).all(), k
    else:
      assert np.allclose(a, b, rtol=1e-5, atol=1e-6 * float(np.abs(b).max() + 1e-30)), k
