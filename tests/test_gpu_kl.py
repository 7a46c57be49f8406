"""emb_rssm_kl_fwd/bwd (the fused KL / free-nats / entropy reduction of
RSSM.loss, dreamerv3/rssm.py:120-133) against the oracle's restatement
(oracle/dreamer_oracle.py unimix_logits / cat_kl / cat_entropy) with torch
autograd for the gradients.  fp32 tolerance 1e-5 (max-abs relative)."""
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
from embodied_b200.dreamerv3 import ops      # noqa: E402
from oracle import dreamer_oracle as do      # noqa: E402


def oracle_kl(post, prior, unimix, free):
  q = do.unimix_logits(prior.float(), unimix)
  p = do.unimix_logits(post.float(), unimix)
  dyn = torch.clamp(do.cat_kl(p.detach(), q), min=free)
  rep = torch.clamp(do.cat_kl(p, q.detach()), min=free)
  return dyn, rep, do.cat_entropy(p).sum(-1), do.cat_entropy(q).sum(-1)


def rel(a, b):
  return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize('B,T,S,C', [(1, 1, 1, 4), (3, 5, 6, 8), (16, 64, 32, 64), (2, 3, 40, 96),
                                     (2, 2, 4, 128)])
@pytest.mark.parametrize('free', [0.0, 1.0])
def test_matches_oracle_fp32(B, T, S, C, free):
  g = torch.Generator().manual_seed(B * T + S * C)
  # a spread of scales so that some rows fall below and some above free_nats
  scale = torch.rand(B, T, 1, 1, generator=g) * 3
  post = (torch.randn(B, T, S, C, generator=g) * scale)
  prior = (torch.randn(B, T, S, C, generator=g) * scale)
  gd, gr = torch.randn(B, T, generator=g), torch.randn(B, T, generator=g)
  p0, q0 = post.clone().requires_grad_(True), prior.clone().requires_grad_(True)
  dyn0, rep0, ep0, eq0 = oracle_kl(p0, q0, 0.01, free)
  ((dyn0 * gd).sum() + (rep0 * gr).sum()).backward()
  p1 = post.cuda().requires_grad_(True)
  q1 = prior.cuda().requires_grad_(True)
  assert ops.kl_supported(p1, q1)
  dyn1, rep1, ep1, eq1 = ops.rssm_kl(p1, q1, 0.01, free)
  ((dyn1 * gd.cuda()).sum() + (rep1 * gr.cuda()).sum()).backward()
  for a, b in ((dyn1, dyn0), (rep1, rep0), (ep1, ep0), (eq1, eq0)):
    assert rel(a.cpu(), b) < 1e-5
  assert rel(p1.grad.cpu(), p0.grad) < 2e-5
  assert rel(q1.grad.cpu(), q0.grad) < 2e-5


def test_strided_views_and_bf16_prior():
  """The posterior logits arrive as a (B, T) transposed view of the scan's
  time-major buffer, the prior logits in the compute dtype."""
  T, R, B, S, C = 7, 16, 5, 32, 64
  g = torch.Generator().manual_seed(3)
  buf = torch.randn(T, R, S * C, generator=g).cuda()
  post = buf[:, :B].transpose(0, 1).reshape(B, T, S, C)
  prior = torch.randn(B, T, S, C, generator=g).cuda().to(torch.bfloat16)
  assert not post.is_contiguous()
  dyn, rep, ep, eq = ops.rssm_kl(post, prior, 0.01, 1.0)
  dyn0, rep0, ep0, eq0 = oracle_kl(post.cpu().contiguous(), prior.cpu().float(), 0.01, 1.0)
  assert rel(dyn.cpu(), dyn0) < 1e-5 and rel(eq.cpu(), eq0) < 1e-5 and rel(ep.cpu(), ep0) < 1e-5


def test_argument_validation():
  lib = ops._lib_bound()
  args = ops.KlArgs(None, None, 0, 0, 1, 1, 1, 200, 0, 0, 0, 0, 0.01, 1.0)
  import ctypes
  assert lib.emb_rssm_kl_fwd(ctypes.byref(args), None, None, None, None, None, None) == -1
  assert b'classes' in lib.emb_last_error()


@pytest.mark.parametrize('rows,L', [(1, 2), (16, 64), (1024, 16), (5, 1)])
def test_lambda_return_kernel_matches_recurrence(rows, L):
  """emb_lambda_return against the reference's reversed Python recurrence
  (dreamerv3/agent.py:482-490)."""
  g = torch.Generator().manual_seed(rows + L)
  last = (torch.rand(rows, L, generator=g) < 0.1).float()
  term = (torch.rand(rows, L, generator=g) < 0.1).float()
  rew, boot = torch.randn(rows, L, generator=g), torch.randn(rows, L, generator=g)
  disc, lam = 1 - 1 / 333, 0.95
  live = (1 - term)[:, 1:] * disc
  cont = (1 - last)[:, 1:] * lam
  interm = rew[:, 1:] + (1 - cont) * live * boot[:, 1:]
  rets = [boot[:, -1]]
  for t in reversed(range(live.shape[1])):
    rets.append(interm[:, t] + live[:, t] * cont[:, t] * rets[-1])
  want = torch.stack(list(reversed(rets))[:-1], 1) if L > 1 else torch.zeros(rows, 0)
  got = ops.lambda_return(last.cuda(), term.cuda(), rew.cuda(), boot.cuda(), disc, lam).cpu()
  assert got.shape == want.shape
  if L > 1:
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize('n,S,C', [(3, 8, 4), (256, 32, 64), (5, 16, 96)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_onehot_sample_kernel_matches_formula(n, S, C, dtype):
  """emb_onehot_sample against argmax(log(unimix(softmax(logit))) + gumbel)
  (embodied/jax/outs.py:210-216,252-270)."""
  g = torch.Generator().manual_seed(n + S + C)
  logit = (torch.randn(n, S, C, generator=g) * 2).to(dtype)
  gumbel = do.make_noise(do.tiny_config(stoch=S, classes=C), n, 1, seed=3)['observe'][:, 0]
  lg = do.unimix_logits(logit.float(), 0.01)
  want = torch.nn.functional.one_hot(torch.argmax(lg + gumbel, -1), C).float()
  got = ops.onehot_sample(logit.cuda(), gumbel.cuda(), 0.01, dtype)
  assert got.dtype == dtype and got.shape == (n, S, C)
  assert torch.equal(got.float().cpu(), want)
  # row-strided noise (one step of a (n, H, S, C) buffer) and output (columns of a wider row)
  noise = torch.zeros(n, 3, S, C, device='cuda')
  noise[:, 1] = gumbel.cuda()
  wide = torch.full((n, 5 + S * C), 7.0, device='cuda').to(dtype)
  res = ops.onehot_sample(logit.cuda(), noise[:, 1], 0.01, dtype, out=wide[:, 5:])
  assert res.data_ptr() == wide[:, 5:].data_ptr()
  assert torch.equal(wide[:, 5:].float().cpu().reshape(n, S, C), want)
  assert float((wide[:, :5] - 7).abs().max()) == 0
