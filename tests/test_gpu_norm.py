"""emb_rmsnorm_act_fwd/bwd against the plain PyTorch fp32 formula of the
reference's Norm('rms') + silu (embodied/jax/nets.py:361-399).  fp32: 1e-5
relative (max-abs / max-abs); bf16: one bf16 ulp (2^-8) of the fp32 result."""
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
from embodied_b200.dreamerv3 import ops      # noqa: E402


def reference(x, scale, act, dtype):
  xf = x.float()
  y = xf * (torch.rsqrt(xf.square().mean(-1, keepdim=True) + 1e-4) * scale)
  y = y.to(dtype)
  return torch.nn.functional.silu(y) if act else y


def rel(a, b):
  return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize('rows,cols', [(1, 8), (7, 64), (1000, 128), (33, 1024), (4, 2048), (5, 8192)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('act', [True, False])
def test_forward(rows, cols, dtype, act):
  g = torch.Generator(device='cuda').manual_seed(rows * cols)
  x = (torch.randn(rows, cols, generator=g, device='cuda') * 3).to(dtype)
  scale = torch.rand(cols, generator=g, device='cuda') + 0.5
  got = ops.rmsnorm_act(x, scale, act)
  want = reference(x, scale, act, dtype)
  assert got.dtype == dtype
  assert rel(got, want) < (1e-5 if dtype == torch.float32 else 2 ** -7)


@pytest.mark.parametrize('shape', [(3, 8), (2, 5, 64), (1000, 128), (9, 4, 4, 256), (300, 512), (37, 640), (2000, 1024), (4, 2048)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('act', [True, False])
def test_backward(shape, dtype, act):
  g = torch.Generator(device='cuda').manual_seed(len(shape) * shape[-1])
  x = (torch.randn(*shape, generator=g, device='cuda') * 2).to(dtype)
  scale = torch.rand(shape[-1], generator=g, device='cuda') + 0.5
  gy = torch.randn(*shape, generator=g, device='cuda').to(dtype)
  x1, s1 = x.clone().requires_grad_(True), scale.clone().requires_grad_(True)
  ops.rmsnorm_act(x1, s1, act).backward(gy)
  # fp32 reference of the same function (no bf16 rounding of intermediates)
  x2, s2 = x.float().clone().requires_grad_(True), scale.clone().requires_grad_(True)
  reference(x2, s2, act, torch.float32).backward(gy.float())
  tol = 2e-5 if dtype == torch.float32 else 3e-2
  assert rel(x1.grad, x2.grad) < tol
  assert rel(s1.grad, s2.grad) < tol


def test_unsupported_shapes_fall_back():
  x = torch.randn(3, 12, device='cuda', dtype=torch.bfloat16)
  assert not ops.rmsnorm_supported(x, False)
  assert ops.rmsnorm_supported(torch.randn(3, 8192, device='cuda'), False)
  assert not ops.rmsnorm_supported(torch.randn(3, 8192, device='cuda'), True)


@pytest.mark.parametrize('shape', [(5, 8), (1000, 128), (77, 192), (9, 4, 4, 256)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_bias_folded_into_the_norm(shape, dtype):
  """y = silu(rms(x + bias) * scale): the convolution bias added inside the
  norm kernel (values and all three gradients) against the fp32 formula."""
  g = torch.Generator(device='cuda').manual_seed(shape[-1])
  x = (torch.randn(*shape, generator=g, device='cuda') * 2).to(dtype)
  scale = torch.rand(shape[-1], generator=g, device='cuda') + 0.5
  bias = torch.randn(shape[-1], generator=g, device='cuda')
  gy = torch.randn(*shape, generator=g, device='cuda').to(dtype)
  x1 = x.clone().requires_grad_(True)
  s1, b1 = scale.clone().requires_grad_(True), bias.clone().requires_grad_(True)
  assert ops.rmsnorm_supported(x1, True, True)
  y1 = ops.rmsnorm_act(x1, s1, True, bias=b1)
  y1.backward(gy)
  x2 = x.float().clone().requires_grad_(True)
  s2, b2 = scale.clone().requires_grad_(True), bias.clone().requires_grad_(True)
  y2 = reference(x2 + b2, s2, True, torch.float32)
  y2.backward(gy.float())
  tol = 2e-5 if dtype == torch.float32 else 3e-2
  assert rel(y1, y2) < (1e-5 if dtype == torch.float32 else 2 ** -6)
  assert rel(x1.grad, x2.grad) < tol
  assert rel(s1.grad, s2.grad) < tol
  assert rel(b1.grad, b2.grad) < tol
  assert not ops.rmsnorm_supported(torch.randn(3, 512, device='cuda'), True, True)


@pytest.mark.parametrize('g,M,Dg', [(8, 5, 64), (4, 33, 16), (8, 1024, 1024), (1, 3, 8)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_grouped_norm_and_gru_gates_forward(g, M, Dg, dtype):
  """emb_rmsnorm_grouped_fwd / emb_gru_gates_fwd (the no-gradient block-GRU path)
  against the reference's op-by-op formulation (dreamerv3/rssm.py:147-158)."""
  gen = torch.Generator(device='cuda').manual_seed(g * M + Dg)
  D = g * Dg
  y = torch.randn(g, M, Dg, generator=gen, device='cuda').to(dtype)
  scale = torch.rand(D, generator=gen, device='cuda') + 0.5
  bias = torch.randn(D, generator=gen, device='cuda') * 0.3
  with torch.no_grad():
    got = ops.rmsnorm_grouped(y, scale, bias)
    flat = (y.transpose(0, 1).reshape(M, D) + bias.to(dtype))
    want = reference(flat, scale, True, dtype).reshape(M, g, Dg).transpose(0, 1)
    tol = 1e-5 if dtype == torch.float32 else 2 ** -6
    assert rel(got, want) < tol
    pre = torch.randn(g, M, 3 * Dg, generator=gen, device='cuda').to(dtype)
    b3 = torch.randn(3 * D, generator=gen, device='cuda') * 0.3
    deter = torch.randn(M, D, generator=gen, device='cuda').to(dtype)
    out = ops.gru_gates(pre, b3, deter)
    x = pre.transpose(0, 1).reshape(M, -1) + b3.to(dtype)
    r, c, u = [t.reshape(M, -1) for t in x.reshape(M, g, -1).chunk(3, -1)]
    r = torch.sigmoid(r)
    c = torch.tanh(r * c)
    u = torch.sigmoid(u - 1)
    ref = u * c + (1 - u) * deter
    assert rel(out, ref) < (1e-5 if dtype == torch.float32 else 3e-2)
    # row-strided operands: deter read from, and the result written into, the deter
    # columns of a wider (rows, steps, deter | stoch) buffer, as the roll-out does
    buf = torch.full((M, 3, D + 24), 7.0, device='cuda').to(dtype)
    buf[:, 0, :D] = deter
    res = ops.gru_gates(pre, b3, buf[:, 0, :D], out=buf[:, 1, :D])
    assert res.data_ptr() == buf[:, 1, :D].data_ptr()
    assert torch.equal(buf[:, 1, :D], out)
    assert float((buf[:, 1, D:] - 7).abs().max()) == 0 and float((buf[:, 2] - 7).abs().max()) == 0
  assert not ops.core_fused_supported(deter, g)      # grad mode: the differentiable path runs
