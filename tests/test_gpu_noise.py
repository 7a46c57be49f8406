"""emb_gumbel_fill: the sampling noise of the categorical draws (noise is an explicit
input; what is checked is the distribution, the bounds and reproducibility under a seed)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fill(n, seed):
  from embodied_b200 import _lib
  out = torch.full((n + 8,), 123.0, device='cuda')
  _lib.check(_lib.load().emb_gumbel_fill(out.data_ptr(), n, seed, torch.cuda.current_stream().cuda_stream))
  torch.cuda.synchronize()
  assert (out[n:] == 123.0).all()                     # nothing written past n
  return out[:n]


def test_gumbel_moments_bounds_and_seeds():
  n = 4_000_003                                        # not a multiple of 4: scalar tail
  a, b, c = _fill(n, 7), _fill(n, 7), _fill(n, 8)
  assert torch.equal(a, b) and not torch.equal(a, c)
  x = a.double().cpu().numpy()
  assert np.isfinite(x).all() and x.min() >= -3.84 and x.max() <= 16.2
  np.testing.assert_allclose(x.mean(), 0.5772156649, atol=4e-3)           # Euler-Mascheroni
  np.testing.assert_allclose(x.var(), math.pi ** 2 / 6, rtol=1e-2)
  # arg-max of logits + Gumbel noise is a categorical draw with softmax probabilities
  logits = torch.tensor([0.0, 1.0, -1.0, 2.0], device='cuda')
  g = _fill(4 * 500_000, 11).reshape(-1, 4)
  freq = torch.bincount(torch.argmax(logits + g, -1), minlength=4).double() / len(g)
  np.testing.assert_allclose(freq.cpu().numpy(), torch.softmax(logits, 0).double().cpu().numpy(), atol=3e-3)
  # neighbouring values are not correlated (subsequences per thread, four values per counter)
  assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 3e-3


def test_model_gumbel_uses_the_kernel_and_advances():
  from embodied_b200 import _lib
  from embodied_b200.dreamerv3 import model
  gen = torch.Generator(device='cuda').manual_seed(5)
  before = _lib.launch_count()
  a = model.gumbel_like((16, 32, 64), 'cuda', gen)
  b = model.gumbel_like((16, 32, 64), 'cuda', gen)
  assert _lib.launch_count() - before == 2
  assert not torch.equal(a, b) and torch.isfinite(a).all()
  c = torch.empty(16, 33, dtype=torch.float32, device='cuda')[:, 1:]       # not contiguous: torch path
  model.gumbel_(c, gen)
  assert torch.isfinite(c).all()
