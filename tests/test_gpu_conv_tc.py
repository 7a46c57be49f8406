"""emb_conv5x5_nhwc_tc (tcgen05 implicit-GEMM convolution) against an fp32
convolution of the same bf16-rounded operands.  Tolerance: the kernel accumulates
bf16 products in fp32 (exact products, different summation order) and rounds the
result to bf16 once: |err| <= 2^-8 |y| + a few fp32 ulps of the accumulated magnitude."""
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import torch.nn.functional as F                      # noqa: E402
from embodied_b200.dreamerv3 import ops              # noqa: E402

@pytest.fixture(autouse=True)
def _strict_fp32_reference():
  """The fp32 reference convolutions must not run on TF32 tensor cores (other tests in the same
  process switch the library's algorithm search on, which picks such kernels)."""
  saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
           torch.backends.cudnn.benchmark)
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  torch.backends.cudnn.benchmark = False
  yield
  (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
   torch.backends.cudnn.benchmark) = saved


SHAPES = [  # n, h, w, cin, cout, k  -- the dreamerv3 size200m layers (rssm.py:233-240, 336-352) + edge cases
    (4, 32, 32, 128, 192, 5), (4, 16, 16, 192, 256, 5), (4, 8, 8, 256, 256, 5),
    (2, 16, 16, 256, 192, 5), (2, 32, 32, 192, 128, 5),
    (2, 8, 8, 64, 32, 5), (3, 32, 32, 64, 64, 3), (8, 4, 4, 64, 96, 5), (8, 4, 8, 128, 64, 1),
    (150, 16, 16, 64, 64, 5),       # more tiles than SMs: every CTA loops, both accumulator buffers reused
    # widths that do not divide 128 (the reference's Atari config is 96x96, dreamerv3/configs.yaml:30):
    # partially filled tiles of whole rows / whole images -- 96, 96, 96, 72, 72 (3 x 5: 15 pixels x 8 images = 120) pixels
    (3, 96, 96, 64, 128, 5), (2, 48, 48, 128, 192, 5), (4, 24, 24, 192, 256, 5), (6, 12, 12, 256, 256, 5),
    (10, 6, 6, 256, 64, 3), (16, 3, 5, 64, 32, 5), (5, 7, 9, 64, 64, 3), (1, 2, 100, 64, 96, 5),
]


def reference(x, w, bias=None):
  y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1),
               None if bias is None else bias, padding=w.shape[0] // 2)
  return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize('n,h,w,cin,cout,k', SHAPES)
def test_forward_matches_fp32_convolution(n, h, w, cin, cout, k):
  g = torch.Generator(device='cuda').manual_seed(n * 1000 + cin + cout + k)
  x = torch.randn((n, h, w, cin), generator=g, device='cuda').to(torch.bfloat16)
  wt = (torch.randn((k, k, cin, cout), generator=g, device='cuda') / (k * cin ** 0.5)).to(torch.bfloat16)
  assert ops.conv_tc_supported(x, cin, cout, k)
  y = ops.conv_tc(x, ops.pack_conv_weight(wt), k=k)
  ref = reference(x, wt)
  err = (y.float() - ref).abs()
  tol = 2.0 ** -8 * ref.abs() + 1e-3 * ref.abs().max()
  assert bool((err <= tol).all()), float((err - tol).max())
  # the borders are where the TMA's zero fill stands in for the SAME padding
  for sl in (y[:, 0], y[:, -1], y[:, :, 0], y[:, :, -1]):
    assert float(sl.float().abs().max()) > 0


def test_bias_and_data_gradient_packing():
  n, h, w, cin, cout, k = 2, 16, 16, 128, 192, 5
  g = torch.Generator(device='cuda').manual_seed(0)
  x = torch.randn((n, h, w, cin), generator=g, device='cuda').to(torch.bfloat16)
  wt = (torch.randn((k, k, cin, cout), generator=g, device='cuda') / 50).to(torch.bfloat16)
  bias = torch.randn(cout, generator=g, device='cuda')
  y = ops.conv_tc(x, ops.pack_conv_weight(wt), bias=bias, k=k)
  ref = reference(x, wt, bias)
  assert float((y.float() - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())
  # input gradient of sum(conv(x) * gy) = the same launch on gy with the flipped, transposed kernel
  gy = torch.randn((n, h, w, cout), generator=g, device='cuda').to(torch.bfloat16)
  xf = x.float().requires_grad_(True)
  (reference(xf, wt) * gy.float()).sum().backward()
  gx = ops.conv_tc(gy, ops.pack_conv_weight(wt, data_grad=True), k=k)
  assert gx.shape == x.shape
  assert float((gx.float() - xf.grad).abs().max()) <= 2.0 ** -7 * float(xf.grad.abs().max())


@pytest.mark.parametrize('n,h,w,cin,cout,k', [(2, 48, 48, 128, 192, 5), (3, 12, 12, 256, 128, 3), (4, 6, 6, 64, 64, 5)])
def test_autograd_function_on_partial_tiles(n, h, w, cin, cout, k):
  """ConvTC on a width that does not divide 128: forward and input gradient from the tcgen05 kernel
  (partially filled tiles), weight gradient from whichever path takes the shape."""
  g = torch.Generator(device='cuda').manual_seed(7)
  x = torch.randn((n, h, w, cin), generator=g, device='cuda').to(torch.bfloat16).requires_grad_(True)
  wt = (torch.randn((k, k, cin, cout), generator=g, device='cuda') / (k * cin ** 0.5)).requires_grad_(True)
  gy = torch.randn((n, h, w, cout), generator=g, device='cuda').to(torch.bfloat16)
  y = ops.ConvTC.apply(x, wt.to(torch.bfloat16))
  (y.float() * gy.float()).sum().backward()
  xf = x.detach().float().requires_grad_(True)
  wf = wt.detach().to(torch.bfloat16).float().requires_grad_(True)
  (reference(xf, wf) * gy.float()).sum().backward()
  assert float((y.float() - reference(xf, wf)).abs().max()) <= 2.0 ** -7 * float(reference(xf, wf).abs().max())
  assert float((x.grad.float() - xf.grad).abs().max()) <= 2.0 ** -6 * float(xf.grad.abs().max())
  assert float((wt.grad - wf.grad).abs().max()) <= 2.0 ** -6 * float(wf.grad.abs().max())


def test_rejects_unsupported_shapes():
  x = torch.zeros((2, 8, 8, 48), dtype=torch.bfloat16, device='cuda')
  assert not ops.conv_tc_supported(x, 48, 64)
  wp = torch.zeros((25, 64, 48), dtype=torch.bfloat16, device='cuda')
  with pytest.raises(RuntimeError, match='multiple of 64'):
    ops.conv_tc(x, wp)


def test_autograd_function_matches_library_convolution():
  """ops.ConvTC (forward + both gradients) against torch's bf16 convolution of the same
  operands: outputs to bf16 rounding, gradients to 1e-2 relative L2."""
  n, h, w, cin, cout, k = 8, 16, 16, 192, 256, 5
  g = torch.Generator(device='cuda').manual_seed(3)
  x0 = torch.randn((n, h, w, cin), generator=g, device='cuda').to(torch.bfloat16)
  w0 = (torch.randn((k, k, cin, cout), generator=g, device='cuda') / 70).to(torch.bfloat16)
  gy = torch.randn((n, h, w, cout), generator=g, device='cuda').to(torch.bfloat16)
  res = []
  for own in (True, False):
    x, wt = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    if own:
      y = ops.ConvTC.apply(x, wt)
    else:
      y = F.conv2d(x.permute(0, 3, 1, 2), wt.permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
    (y.float() * gy.float()).sum().backward()
    res.append((y.detach().float(), x.grad.float(), wt.grad.float()))
  rel2 = lambda a, b: float((a - b).norm() / b.norm())
  assert rel2(res[0][0], res[1][0]) < 1e-2
  assert rel2(res[0][1], res[1][1]) < 1e-2
  assert rel2(res[0][2], res[1][2]) < 1e-2


WSHAPES = [  # n, h, w, cin, cout, k
    (8, 32, 32, 128, 192, 5), (8, 16, 16, 192, 256, 5), (8, 8, 8, 256, 256, 5),
    (8, 16, 16, 256, 192, 5), (8, 32, 32, 192, 128, 5), (4, 16, 16, 128, 64, 3), (300, 8, 8, 128, 128, 5),
    # image sizes whose rows do not pack into 64-pixel chunks (96 x 96 and its down-sampled maps, odd sizes):
    # partially filled chunks on a zero-initialised ring, row pieces where a row is wider than 64 pixels
    (3, 96, 96, 128, 64, 5), (2, 48, 48, 128, 192, 5), (4, 24, 24, 192, 256, 5), (6, 12, 12, 256, 256, 5),
    (10, 6, 6, 256, 128, 3), (9, 3, 5, 128, 64, 5), (5, 7, 9, 64, 128, 3), (2, 4, 100, 128, 128, 5),
]


@pytest.mark.parametrize('n,h,w,cin,cout,k', WSHAPES)
def test_weight_gradient_matches_fp32(n, h, w, cin, cout, k):
  """emb_conv5x5_wgrad_tc: exact bf16 products accumulated in fp32 (tensor memory, then
  red.global.add across the pixel splits) against fp32 autograd of the same operands."""
  g = torch.Generator(device='cuda').manual_seed(n + cin + cout)
  x = torch.randn((n, h, w, cin), generator=g, device='cuda').to(torch.bfloat16)
  gy = torch.randn((n, h, w, cout), generator=g, device='cuda').to(torch.bfloat16)
  assert ops.conv_wgrad_tc_supported(x, gy, k)
  dw = ops.conv_wgrad(x, gy, k)
  assert dw.shape == (k, k, cin, cout) and dw.dtype == torch.float32
  # float64 reference without any library convolution: one GEMM over the pixels per tap
  pad = k // 2
  xp = F.pad(x.double(), (0, 0, pad, pad, pad, pad))
  g2 = gy.double().reshape(-1, cout)
  want = torch.stack([
      torch.stack([xp[:, ky: ky + h, kx: kx + w].reshape(-1, cin).t() @ g2 for kx in range(k)])
      for ky in range(k)])
  err = float((dw.double() - want).abs().max())
  wt = type('G', (), {'grad': want})
  assert err <= 1e-4 * float(wt.grad.abs().max()), err


@pytest.mark.parametrize('n,h,w,cin,cout', [(8, 4, 4, 256, 256), (4, 8, 8, 256, 192), (4, 16, 16, 192, 128),
                                            (3, 6, 6, 256, 256), (2, 12, 12, 256, 192), (2, 24, 24, 192, 128), (3, 5, 7, 128, 64)])
def test_subpixel_upconv_matches_upsample_then_conv(n, h, w, cin, cout):
  """ops.upconv_subpixel (four 3x3 phase convolutions on the low-resolution grid) against
  nearest x2 up-sampling + 5x5 SAME convolution in fp32 on the same bf16 operands: output, input
  gradient and kernel gradient.  The folded kernels are sums of up to four bf16 taps rounded to
  bf16 once, hence relative-L2 tolerances at bf16 resolution."""
  g = torch.Generator(device='cuda').manual_seed(h + cin)
  x0 = torch.randn((n, h, w, cin), generator=g, device='cuda').to(torch.bfloat16)
  w0 = (torch.randn((5, 5, cin, cout), generator=g, device='cuda') / 60).to(torch.bfloat16)
  gy = torch.randn((n, 2 * h, 2 * w, cout), generator=g, device='cuda').to(torch.bfloat16)
  x, wt = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
  assert ops.subpixel_supported(x, cin, cout)
  y = ops.upconv_subpixel(x, wt)
  (y.float() * gy.float()).sum().backward()
  xr, wr = x0.float().requires_grad_(True), w0.float().requires_grad_(True)
  up = F.interpolate(xr.permute(0, 3, 1, 2), scale_factor=2, mode='nearest')
  yr = F.conv2d(up, wr.permute(3, 2, 0, 1), padding=2).permute(0, 2, 3, 1)
  (yr * gy.float()).sum().backward()
  rel2 = lambda a, b: float((a.float() - b).norm() / b.norm())
  assert y.shape == yr.shape
  assert rel2(y, yr) < 6e-3, rel2(y, yr)
  assert rel2(x.grad, xr.grad) < 6e-3, rel2(x.grad, xr.grad)
  assert rel2(wt.grad, wr.grad) < 6e-3, rel2(wt.grad, wr.grad)


@pytest.mark.parametrize('P,K,N', [(64 * 200, 80, 128), (64 * 150, 128, 80), (64 * 3, 256, 8), (64 * 40, 24, 256)])
def test_thin_matmul_weight_gradient(P, K, N):
  """ops.ThinMatmul: y = a @ w with millions of pixel rows and a small matrix; the weight gradient
  a^T @ gy comes from the tcgen05 weight-gradient kernel with ksize = 1 (the N side padded to 64
  channels by TMA zero fill)."""
  g = torch.Generator(device='cuda').manual_seed(P + K + N)
  a = torch.randn((P, K), generator=g, device='cuda').to(torch.bfloat16).requires_grad_(K in (128, 256))
  w = (torch.randn((K, N), generator=g, device='cuda') / K ** 0.5).to(torch.bfloat16).requires_grad_(True)
  gy = torch.randn((P, N), generator=g, device='cuda').to(torch.bfloat16)
  assert ops.thin_matmul_supported(a, w)
  y = ops.ThinMatmul.apply(a, w)
  (y.float() * gy.float()).sum().backward()
  want_w = a.detach().double().t() @ gy.double()
  err = float((w.grad.double() - want_w).abs().max())
  assert err <= 2.0 ** -7 * float(want_w.abs().max()), err          # the result is rounded to bf16 once
  assert float((y.float() - a.detach().float() @ w.detach().float()).abs().max()) <= 2.0 ** -6 * float(y.float().abs().max())
  if a.requires_grad:
    want_a = gy.float() @ w.detach().float().t()
    assert float((a.grad.float() - want_a).abs().max()) <= 2.0 ** -6 * float(want_a.abs().max())
