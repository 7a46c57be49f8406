"""emb_maxpool2_nhwc_* / emb_upsample2_nhwc_* against PyTorch on NHWC tensors
(dreamerv3/rssm.py:239-240, :336,349).  Both are data movement plus a max:
bit-exact for fp32 and bf16, forward and backward (ties: first maximum in
(dy, dx) order, as torch.max_pool2d)."""
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import torch.nn.functional as F               # noqa: E402
from embodied_b200.dreamerv3 import ops      # noqa: E402

SHAPES = [(1, 2, 2, 8), (3, 4, 6, 16), (5, 16, 16, 24), (2, 64, 64, 128)]


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_maxpool(shape, dtype):
  g = torch.Generator(device='cuda').manual_seed(sum(shape))
  x = torch.randn(*shape, generator=g, device='cuda').to(dtype)
  x[0, 0, 0, :] = 1.0
  x[0, 0, 1, :] = 1.0                                    # a tie: first wins
  x1 = x.clone().requires_grad_(True)
  y1 = ops.MaxPool2.apply(x1)
  x2 = x.clone().requires_grad_(True)
  y2 = F.max_pool2d(x2.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
  assert torch.equal(y1, y2)
  gy = torch.randn(y1.shape, generator=g, device='cuda').to(dtype)
  y1.backward(gy)
  y2.backward(gy)
  assert torch.equal(x1.grad, x2.grad)


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_upsample(shape, dtype):
  g = torch.Generator(device='cuda').manual_seed(sum(shape) + 1)
  x = torch.randn(*shape, generator=g, device='cuda').to(dtype)
  x1 = x.clone().requires_grad_(True)
  y1 = ops.Upsample2.apply(x1)
  want = x.repeat_interleave(2, 2).repeat_interleave(2, 1)
  assert torch.equal(y1, want)
  gy = torch.randn(y1.shape, generator=g, device='cuda').to(dtype)
  y1.backward(gy)
  n, h, w, c = shape
  gref = gy.float().reshape(n, h, 2, w, 2, c)
  gref = ((gref[:, :, 0, :, 0] + gref[:, :, 0, :, 1]) + (gref[:, :, 1, :, 0] + gref[:, :, 1, :, 1])).to(dtype)
  assert torch.equal(x1.grad, gref)


def test_odd_channels_not_supported():
  assert not ops.spatial_supported(torch.zeros(1, 4, 4, 12, device='cuda', dtype=torch.bfloat16))
  assert ops.spatial_supported(torch.zeros(1, 4, 4, 12, device='cuda'))
