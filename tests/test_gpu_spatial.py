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


# ------------------------------------------------------------ thin convolutions
def _conv_ref(x, w, b=None):
  """NHWC x, HWIO w, SAME -- the reference's Conv2D (embodied/jax/nets.py:298-323)."""
  torch.backends.cudnn.allow_tf32 = False
  y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=w.shape[0] // 2)
  return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize('n,h,w,cin,cout,k', [(2, 8, 8, 3, 16, 5), (3, 6, 10, 1, 8, 3), (1, 64, 64, 3, 128, 5)])
def test_thin_input_conv_matches_conv2d(n, h, w, cin, cout, k):
  g = torch.Generator(device='cuda').manual_seed(n * h + cout)
  x = torch.randn(n, h, w, cin, generator=g, device='cuda')
  wt = (torch.randn(k, k, cin, cout, generator=g, device='cuda') * 0.2).requires_grad_(True)
  kp = ops.patch_columns(k, cin)
  w2 = torch.nn.functional.pad(wt.reshape(k * k * cin, cout), (0, 0, 0, kp - k * k * cin))
  y = (ops.ConvPatches.apply(x, k) @ w2).reshape(n, h, w, cout)
  wr = wt.detach().clone().requires_grad_(True)
  yr = _conv_ref(x, wr)
  assert float((y - yr).abs().max()) < 1e-4 * float(yr.abs().max())
  gy = torch.randn_like(y)
  y.backward(gy)
  yr.backward(gy)
  assert float((wt.grad - wr.grad).abs().max()) < 1e-4 * float(wr.grad.abs().max())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('n,h,w,cin,cout,k,up', [(2, 8, 8, 16, 3, 5, 1), (2, 4, 6, 8, 3, 5, 2),
                                                   (1, 32, 32, 128, 3, 5, 2), (3, 6, 6, 8, 1, 3, 2)])
def test_thin_output_conv_matches_conv2d(n, h, w, cin, cout, k, up, dtype):
  """x (n, h, w, cin) -> [nearest x`up`] -> SAME conv to cout channels + bias,
  values and all three gradients against upsample + conv2d in fp32."""
  g = torch.Generator(device='cuda').manual_seed(h * w + cin)
  x = torch.randn(n, h, w, cin, generator=g, device='cuda').to(dtype)
  wt = (torch.randn(k, k, cin, cout, generator=g, device='cuda') * 0.2).to(dtype)
  b = torch.randn(cout, generator=g, device='cuda')
  x1, w1, b1 = (t.clone().requires_grad_(True) for t in (x, wt, b))
  kp = ops.patch_columns(k, cout)
  w2 = torch.nn.functional.pad(w1.permute(2, 0, 1, 3).reshape(cin, k * k * cout), (0, kp - k * k * cout))
  z = x1.reshape(n * h * w, cin) @ w2
  y = ops.ConvTapSum.apply(z, b1, (n, h * up, w * up, cout), k, up)
  # float64 reference: cuDNN's fp32 wgrad for 3 output channels is only ~4e-3 accurate
  x2, w3, b2 = (t.double().clone().requires_grad_(True) for t in (x, wt, b))
  xu = x2.repeat_interleave(up, 1).repeat_interleave(up, 2) if up > 1 else x2
  yr = _conv_ref(xu, w3, b2)
  tol = 1e-4 if dtype == torch.float32 else 3e-2
  assert float((y.double() - yr).abs().max()) < tol * float(yr.abs().max())
  gy = torch.randn(*yr.shape, generator=g, device='cuda')
  y.backward(gy.to(dtype))
  yr.backward(gy.double())
  for name, a, r in (('x', x1.grad, x2.grad), ('w', w1.grad, w3.grad), ('b', b1.grad, b2.grad)):
    err = float((a.double() - r).abs().max()) / float(r.abs().max())
    assert err < tol, (name, err)
