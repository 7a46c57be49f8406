"""emb_twohot_loss_fwd/_bwd, emb_twohot_pred, emb_loss_reduce (csrc/losses.cu) against the oracle's
restatement of embodied/jax/outs.py:273-330 (TwoHot) and dreamerv3/agent.py:237-240: fp32, 1e-5
relative (BASELINE.json north_star), gradients through torch autograd of the oracle formulation."""
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
from embodied_b200.dreamerv3 import ops          # noqa: E402
from oracle import dreamer_oracle as do           # noqa: E402


def bins(n=255):
  half = do.symexp(torch.linspace(-20, 0, (n - 1) // 2 + 1))
  return torch.cat([half, -half[:-1].flip(0)], 0)


def oracle_loss(logits, target, b):
  """outs.py:311-330, written as in the reference."""
  n = len(b)
  below = (b <= target[..., None]).sum(-1) - 1
  above = n - (b > target[..., None]).sum(-1)
  below, above = below.clamp(0, n - 1), above.clamp(0, n - 1)
  equal = below == above
  one = torch.ones_like(target)
  db = torch.where(equal, one, (b[below] - target).abs())
  da = torch.where(equal, one, (b[above] - target).abs())
  total = db + da
  tgt = (torch.nn.functional.one_hot(below, n) * (da / total)[..., None] +
         torch.nn.functional.one_hot(above, n) * (db / total)[..., None])
  return -(tgt * torch.log_softmax(logits, -1)).sum(-1)


@pytest.mark.parametrize('shape,n', [((16, 64), 255), ((1024, 16), 255), ((7,), 5), ((3, 5), 64)])
def test_twohot_loss_and_gradient(shape, n):
  g = torch.Generator().manual_seed(n + len(shape))
  b = bins(n) if n % 2 else torch.sort(torch.randn(n, generator=g) * 3).values
  logits = torch.randn(*shape, n, generator=g) * 2
  # targets: inside, exactly on a bin, beyond both ends
  t1 = torch.randn(shape, generator=g) * 30
  t2 = torch.randn(shape, generator=g) * 0.1
  t1.view(-1)[0] = b[n // 2]
  t1.view(-1)[1] = b[0] - 5
  t1.view(-1)[2] = b[-1] + 5
  t1.view(-1)[3] = b[1]
  w2 = 0.7
  gl = torch.randn(shape, generator=g)
  lo = logits.clone().requires_grad_(True)
  want = oracle_loss(lo, t1, b) + w2 * oracle_loss(lo, t2, b)
  (want * gl).sum().backward()
  ld = logits.cuda().requires_grad_(True)
  got = ops.twohot_loss(ld, b.cuda(), t1.cuda(), t2.cuda(), w2)
  (got * gl.cuda()).sum().backward()
  assert torch.allclose(got.cpu(), want.detach(), rtol=1e-5, atol=1e-5)
  assert torch.allclose(ld.grad.cpu(), lo.grad, rtol=1e-5, atol=1e-6)
  single = ops.twohot_loss(logits.cuda(), b.cuda(), t1.cuda())
  assert torch.allclose(single.cpu(), oracle_loss(logits, t1, b), rtol=1e-5, atol=1e-5)


def test_twohot_pred_matches_oracle_and_is_zero_at_symmetric_init():
  b = bins(255)
  g = torch.Generator().manual_seed(0)
  logits = torch.randn(33, 9, 255, generator=g) * 3
  probs = torch.softmax(logits, -1)
  m = 127
  want = (probs[..., m: m + 1] * b[m: m + 1]).sum(-1) + (
      (probs[..., :m] * b[:m]).flip(-1) + probs[..., m + 1:] * b[m + 1:]).sum(-1)
  got = ops.twohot_pred(logits.cuda(), b.cuda()).cpu()
  assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))
  zero = ops.twohot_pred(torch.zeros(4, 255, device='cuda'), b.cuda())
  assert bool((zero == 0).all())          # outs.py:286-290: exactly zero, not 1e-9


def test_loss_reduce_total_means_and_gradients():
  g = torch.Generator().manual_seed(1)
  terms = {k: torch.randn(s, generator=g) for k, s in
           {'a': (16, 64), 'b': (16, 64), 'c': (16, 63), 'd': (3,)}.items()}
  scales = {'a': 1.0, 'b': 0.1, 'c': 0.3, 'd': 2.0}
  leaves = {k: v.clone().cuda().requires_grad_(True) for k, v in terms.items()}
  for _ in range(2):                      # second call: the ticket word was reset by the first
    total, means = ops.loss_sum(leaves, scales)
  want = sum(v.mean() * scales[k] for k, v in terms.items())
  assert abs(float(total) - float(want)) <= 1e-5 * abs(float(want))
  for k, v in terms.items():
    assert abs(float(means[k]) - float(v.mean())) <= 1e-5 * abs(float(v.mean())) + 1e-7
  total.backward()
  for k, v in leaves.items():
    assert torch.allclose(v.grad.cpu(), torch.full(v.shape, scales[k] / v.numel()), rtol=1e-6)
