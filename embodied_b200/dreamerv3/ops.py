"""torch.autograd wrappers of the element-wise / reduction kernels in
libembodied_b200.so (include/embodied_b200.h)."""
import ctypes
import os

import torch

from .. import _lib

f32 = torch.float32
_vp, _i32, _i64, _fl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
_bound = False


def _lib_bound():
  global _bound
  lib = _lib.load()
  if not _bound:
    lib.emb_rmsnorm_act_fwd.argtypes = [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _fl, _vp]
    lib.emb_rmsnorm_act_fwd.restype = ctypes.c_int
    lib.emb_rmsnorm_act_bwd.argtypes = [
        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _fl, _vp]
    lib.emb_rmsnorm_act_bwd.restype = ctypes.c_int
    kp = ctypes.POINTER(KlArgs)
    lib.emb_rssm_kl_fwd.argtypes = [kp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.emb_rssm_kl_fwd.restype = ctypes.c_int
    lib.emb_rssm_kl_bwd.argtypes = [kp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.emb_rssm_kl_bwd.restype = ctypes.c_int
    for name in ('emb_maxpool2_nhwc_fwd', 'emb_maxpool2_nhwc_bwd'):
      getattr(lib, name).argtypes = [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]
      getattr(lib, name).restype = ctypes.c_int
    for name in ('emb_upsample2_nhwc_fwd', 'emb_upsample2_nhwc_bwd'):
      getattr(lib, name).argtypes = [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]
      getattr(lib, name).restype = ctypes.c_int
    lib.emb_conv_patches_nhwc.argtypes = [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
    lib.emb_conv_patches_nhwc.restype = ctypes.c_int
    lib.emb_conv_tapsum_nhwc.argtypes = [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
    lib.emb_conv_tapsum_nhwc.restype = ctypes.c_int
    lib.emb_rmsnorm_grouped_fwd.argtypes = [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _fl, _vp]
    lib.emb_rmsnorm_grouped_fwd.restype = ctypes.c_int
    lib.emb_gru_gates_fwd.argtypes = [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i64, _i64, _vp]
    lib.emb_gru_gates_fwd.restype = ctypes.c_int
    lib.emb_onehot_sample.argtypes = [_vp, _i32, _i64, _vp, _i64, _i64, _i32, _i32, _fl, _vp, _i32, _i64, _vp, _vp]
    lib.emb_onehot_sample.restype = ctypes.c_int
    lib.emb_lambda_return.argtypes = [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _fl, _fl, _vp]
    lib.emb_lambda_return.restype = ctypes.c_int
    _bound = True
  return lib


# kernel name -> how many call sites ran the library formulation instead because the kernel does
# not take their shape / dtype (Model._use).  Empty on the benchmarked configuration.
FALLBACKS = {}


def note_fallback(kernel, strict=False):
  if strict:
    raise RuntimeError(f'strict_kernels: no {kernel} kernel for this shape / dtype')
  FALLBACKS[kernel] = FALLBACKS.get(kernel, 0) + 1


def _dtype_code(t):
  return {torch.float32: 0, torch.bfloat16: 1}[t.dtype]


def rmsnorm_supported(x, need_grad, bias=False):
  if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
    return False
  cols = x.shape[-1]
  per = 4 if x.dtype == torch.float32 else 8
  if cols % per or (need_grad and cols > 2048) or (bias and cols > 256):
    return False
  return True


class RmsNormAct(torch.autograd.Function):
  """y = act(rms_norm(x + bias) * scale) over the last axis (nets.py:361-399 +
  act; `bias` = the bias of the preceding convolution, or None)."""

  @staticmethod
  def forward(ctx, x, scale, bias, act, eps):
    lib = _lib_bound()
    x = x.contiguous()
    y = torch.empty_like(x)
    cols = x.shape[-1]
    rows = x.numel() // cols
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_rmsnorm_act_fwd(
        x.data_ptr(), scale.data_ptr(), None if bias is None else bias.data_ptr(), y.data_ptr(),
        rows, cols, _dtype_code(x), int(act), eps, stream))
    ctx.save_for_backward(x, scale, bias)
    ctx.act, ctx.eps = int(act), eps
    return y

  @staticmethod
  def backward(ctx, gy):
    lib = _lib_bound()
    x, scale, bias = ctx.saved_tensors
    gy = gy.contiguous()
    gx = torch.empty_like(x)
    gscale = torch.zeros_like(scale)
    gbias = None if bias is None else torch.zeros_like(bias)
    cols = x.shape[-1]
    rows = x.numel() // cols
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_rmsnorm_act_bwd(
        x.data_ptr(), scale.data_ptr(), None if bias is None else bias.data_ptr(), gy.data_ptr(),
        gx.data_ptr(), gscale.data_ptr(), None if bias is None else gbias.data_ptr(),
        rows, cols, _dtype_code(x), ctx.act, ctx.eps, stream))
    return gx, gscale, gbias, None, None


def rmsnorm_act(x, scale, act=True, eps=1e-4, bias=None):
  if bias is not None:
    bias = bias.to(torch.float32)
  return RmsNormAct.apply(x, scale, bias, act, eps)


# ------------------------------------------------------------------ KL reduction
class KlArgs(ctypes.Structure):
  _fields_ = [('post', _vp), ('prior', _vp), ('dtype_post', _i32), ('dtype_prior', _i32),
              ('B', _i32), ('T', _i32), ('S', _i32), ('C', _i32),
              ('post_stride_b', _i64), ('post_stride_t', _i64),
              ('prior_stride_b', _i64), ('prior_stride_t', _i64),
              ('unimix', _fl), ('free_nats', _fl)]


def _kl_view(x):
  """(B, T, S, C) with the last two axes dense -> (tensor, stride_b, stride_t)."""
  S, C = x.shape[-2:]
  if x.stride(-1) != 1 or x.stride(-2) != C:
    x = x.contiguous()
  return x, x.stride(0), x.stride(1)


def kl_supported(post, prior):
  ok = lambda x: x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16)
  return ok(post) and ok(prior) and post.shape == prior.shape and post.shape[-1] <= 128


class RssmKl(torch.autograd.Function):
  """dyn, rep, ent_post, ent_prior = RSSM.loss's KL block (rssm.py:123-132) for
  logits (B, T, S, C): one kernel each way (emb_rssm_kl_fwd / _bwd)."""

  @staticmethod
  def forward(ctx, post, prior, unimix, free_nats):
    lib = _lib_bound()
    post, psb, pst = _kl_view(post)
    prior, qsb, qst = _kl_view(prior)
    B, T, S, C = post.shape
    out = torch.empty((5, B, T), dtype=f32, device=post.device)
    args = KlArgs(post.data_ptr(), prior.data_ptr(), _dtype_code(post), _dtype_code(prior),
                  B, T, S, C, psb, pst, qsb, qst, unimix, free_nats)
    stream = torch.cuda.current_stream(post.device).cuda_stream
    p = [out[i].data_ptr() for i in range(5)]
    _lib.check(lib.emb_rssm_kl_fwd(ctypes.byref(args), p[0], p[1], p[2], p[3], p[4], stream))
    ctx.save_for_backward(post, prior, out[2])
    ctx.consts = (unimix, free_nats)
    ctx.mark_non_differentiable(out[3], out[4])
    return out[0], out[1], out[3], out[4]

  @staticmethod
  def backward(ctx, g_dyn, g_rep, _a, _b):
    lib = _lib_bound()
    post, prior, kl_raw = ctx.saved_tensors
    post, psb, pst = _kl_view(post)
    prior, qsb, qst = _kl_view(prior)
    B, T, S, C = post.shape
    unimix, free_nats = ctx.consts
    zero = lambda g: torch.zeros((B, T), dtype=f32, device=post.device) if g is None \
        else g.to(f32).contiguous()
    g_dyn, g_rep = zero(g_dyn), zero(g_rep)
    g_post = torch.empty((B, T, S, C), dtype=f32, device=post.device)
    g_prior = torch.empty((B, T, S, C), dtype=f32, device=post.device)
    args = KlArgs(post.data_ptr(), prior.data_ptr(), _dtype_code(post), _dtype_code(prior),
                  B, T, S, C, psb, pst, qsb, qst, unimix, free_nats)
    stream = torch.cuda.current_stream(post.device).cuda_stream
    _lib.check(lib.emb_rssm_kl_bwd(
        ctypes.byref(args), kl_raw.data_ptr(), g_dyn.data_ptr(), g_rep.data_ptr(),
        g_post.data_ptr(), g_prior.data_ptr(), stream))
    return g_post.to(post.dtype), g_prior.to(prior.dtype), None, None


def rssm_kl(post, prior, unimix, free_nats):
  return RssmKl.apply(post, prior, float(unimix), float(free_nats))


def spatial_supported(x):
  """x: NHWC, contiguous, fp32 / bf16, channels a multiple of one 16-byte vector."""
  if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 4:
    return False
  per = 4 if x.dtype == torch.float32 else 8
  return x.shape[-1] % per == 0 and x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0


class MaxPool2(torch.autograd.Function):
  """(N, 2H, 2W, C) -> (N, H, W, C), NHWC (dreamerv3/rssm.py:239-240)."""

  @staticmethod
  def forward(ctx, x):
    lib = _lib_bound()
    x = x.contiguous()
    n, h2, w2, c = x.shape
    h, w = h2 // 2, w2 // 2
    y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
    per = 4 if x.dtype == torch.float32 else 8
    idx = torch.empty(n * h * w * (c // per),
                      dtype=torch.uint8 if per == 4 else torch.int16, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_maxpool2_nhwc_fwd(
        x.data_ptr(), y.data_ptr(), idx.data_ptr(), n, h, w, c, _dtype_code(x), stream))
    ctx.save_for_backward(idx)
    ctx.shape = (n, h, w, c)
    return y

  @staticmethod
  def backward(ctx, gy):
    lib = _lib_bound()
    idx, = ctx.saved_tensors
    n, h, w, c = ctx.shape
    gy = gy.contiguous()
    gx = torch.empty((n, 2 * h, 2 * w, c), dtype=gy.dtype, device=gy.device)
    stream = torch.cuda.current_stream(gy.device).cuda_stream
    _lib.check(lib.emb_maxpool2_nhwc_bwd(
        gy.data_ptr(), idx.data_ptr(), gx.data_ptr(), n, h, w, c, _dtype_code(gy), stream))
    return gx


class Upsample2(torch.autograd.Function):
  """(N, H, W, C) -> (N, 2H, 2W, C) nearest, NHWC (dreamerv3/rssm.py:336,349)."""

  @staticmethod
  def forward(ctx, x):
    lib = _lib_bound()
    x = x.contiguous()
    n, h, w, c = x.shape
    y = torch.empty((n, 2 * h, 2 * w, c), dtype=x.dtype, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_upsample2_nhwc_fwd(
        x.data_ptr(), y.data_ptr(), n, h, w, c, _dtype_code(x), stream))
    ctx.shape = (n, h, w, c)
    return y

  @staticmethod
  def backward(ctx, gy):
    lib = _lib_bound()
    n, h, w, c = ctx.shape
    gy = gy.contiguous()
    gx = torch.empty((n, h, w, c), dtype=gy.dtype, device=gy.device)
    stream = torch.cuda.current_stream(gy.device).cuda_stream
    _lib.check(lib.emb_upsample2_nhwc_bwd(
        gy.data_ptr(), gx.data_ptr(), n, h, w, c, _dtype_code(gy), stream))
    return gx


# ------------------------------------------------------------ thin convolutions
def patch_columns(k, c):
  """Row width of the patch / tap matrices: k*k*c rounded up to 8."""
  return (k * k * c + 7) // 8 * 8


def _patches(x, shape, k, kp, sign, up):
  """x NHWC on the (h*up, w*up) grid -> (n*h*w, kp); see emb_conv_patches_nhwc."""
  lib = _lib_bound()
  n, h, w, c = shape
  x = x.contiguous()
  out = torch.empty((n * h * w, kp), dtype=x.dtype, device=x.device)
  stream = torch.cuda.current_stream(x.device).cuda_stream
  _lib.check(lib.emb_conv_patches_nhwc(
      x.data_ptr(), out.data_ptr(), n, h, w, c, k, kp, sign, up, _dtype_code(x), stream))
  return out


class ConvPatches(torch.autograd.Function):
  """(N, H, W, C) image -> (N*H*W, kp) rows of its k x k SAME patches (zero
  padded).  The input is a constant of the graph (the observation)."""

  @staticmethod
  def forward(ctx, x, k):
    n, h, w, c = x.shape
    return _patches(x, (n, h, w, c), k, patch_columns(k, c), 1, 1)

  @staticmethod
  def backward(ctx, g):
    return None, None


def thin_matmul_supported(a, w):
  """a (P, K) @ w (K, N) bf16 with one side of 128 / 256 columns and the other a multiple of 8:
  the weight gradient is emb_conv_wgrad_tc with ksize = 1 over P / 64 chunks of 64 rows."""
  if not (a.is_cuda and a.dtype == w.dtype == torch.bfloat16 and a.dim() == 2 and a.is_contiguous()):
    return False
  if os.environ.get('EMB_THIN_MATMUL', '1') == '0':          # A/B switch (profiles/)
    return False
  P, K = a.shape
  N = w.shape[1]
  return P % 64 == 0 and ((K in (128, 256) and N % 8 == 0 and N <= 256) or
                          (N in (128, 256) and K % 8 == 0 and K <= 256))


class ThinMatmul(torch.autograd.Function):
  """y = a @ w for the two thin convolutions (a = rows of image patches, or w = the tap matrix of
  the image head): millions of pixel rows times a small matrix.  The forward product is a library
  GEMM; the WEIGHT gradient a^T @ gy -- a reduction over all pixels into an (80 x 128)-sized
  result, HBM-bound -- streams both operands once through the tcgen05 weight-gradient kernel
  (ksize = 1) instead of a split-K library GEMM."""

  @staticmethod
  def forward(ctx, a, w):
    ctx.save_for_backward(a, w)
    return a @ w

  @staticmethod
  def backward(ctx, gy):
    a, w = ctx.saved_tensors
    ga = gw = None
    if ctx.needs_input_grad[0]:
      ga = gy @ w.t()
    if ctx.needs_input_grad[1]:
      P, K = a.shape
      N = w.shape[1]
      gy = gy.contiguous()
      m_is_in = K in (128, 256)
      m, n = (K, N) if m_is_in else (N, K)
      n_pad = (n + 63) // 64 * 64
      dw = torch.zeros((1, m, n_pad), dtype=f32, device=a.device)
      _conv_general(x=a.data_ptr(), gy=gy.data_ptr(), dw=dw.data_ptr(), n=P // 64, h=1, w=64, cin=K,
                    cout=N, ksize=1, m_is_in=int(m_is_in), gy_up=1, gy_phase=0)
      gw = dw[0, :, :n] if m_is_in else dw[0, :, :n].t()
      gw = gw.to(w.dtype)
    return ga, gw


class ConvTapSum(torch.autograd.Function):
  """y[p, ch] = bias[ch] + sum over the k x k taps of z[p + offset, tap*C + ch]:
  the second half of a SAME convolution with few output channels computed as
  z = x @ W[Cin, k*k*C].  up = 2: z lives on the half-resolution grid (nearest
  up-sampling folded in)."""

  @staticmethod
  def forward(ctx, z, bias, shape, k, up):
    lib = _lib_bound()
    n, h, w, c = shape
    z = z.contiguous()
    y = torch.empty(shape, dtype=z.dtype, device=z.device)
    stream = torch.cuda.current_stream(z.device).cuda_stream
    _lib.check(lib.emb_conv_tapsum_nhwc(
        z.data_ptr(), None if bias is None else bias.data_ptr(), y.data_ptr(), n, h, w, c, k,
        z.shape[-1], up, _dtype_code(z), stream))
    ctx.cfg = (shape, k, up, z.shape[-1], bias is not None)
    return y

  @staticmethod
  def backward(ctx, gy):
    (n, h, w, c), k, up, kp, has_bias = ctx.cfg
    gz = _patches(gy, (n, h // up, w // up, c), k, kp, -1, up)
    gb = gy.reshape(-1, c).sum(0, dtype=f32) if has_bias else None
    return gz, gb, None, None, None


def thin_conv_supported(x, cin, cout):
  if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 4:
    return False
  return min(cin, cout) <= 4


# ------------------------------------------- forward-only block-GRU fusions
def core_fused_supported(deter, groups):
  if torch.is_grad_enabled() or not deter.is_cuda or deter.dtype not in (torch.float32, torch.bfloat16):
    return False
  per = 4 if deter.dtype == torch.float32 else 8
  D = deter.shape[-1]
  return D % groups == 0 and (D // groups) % per == 0 and D // per <= 16 * 256


@torch.no_grad()
def rmsnorm_grouped(x, scale, bias, act=True, eps=1e-4):
  """x (g, M, Dg) contiguous -> same layout, norm over the full row (g*Dg)."""
  lib = _lib_bound()
  g, M, Dg = x.shape
  y = torch.empty_like(x)
  stream = torch.cuda.current_stream(x.device).cuda_stream
  _lib.check(lib.emb_rmsnorm_grouped_fwd(
      x.data_ptr(), scale.data_ptr(), None if bias is None else bias.data_ptr(), y.data_ptr(),
      M, g, Dg, _dtype_code(x), int(act), eps, stream))
  return y


@torch.no_grad()
def _rows_ok(x):
  """2-D view whose rows are contiguous and 16-byte aligned (any row stride)."""
  return (x.dim() == 2 and x.stride(1) == 1 and x.data_ptr() % 16 == 0 and
          (x.stride(0) * x.element_size()) % 16 == 0)


def gru_gates(pre, bias, deter, out=None):
  """pre (g, M, 3*Dg) contiguous, bias fp32 (3*D,), deter (M, D) -> new deter (M, D).
  deter and `out` may be row-strided views (columns of a wider buffer)."""
  lib = _lib_bound()
  g, M, Dg3 = pre.shape
  if not _rows_ok(deter):
    deter = deter.contiguous()
  dst = out if out is not None and _rows_ok(out) else torch.empty(
      (M, g * Dg3 // 3), dtype=deter.dtype, device=deter.device)
  assert dst.shape == deter.shape and dst.dtype == deter.dtype == pre.dtype
  stream = torch.cuda.current_stream(pre.device).cuda_stream
  _lib.check(lib.emb_gru_gates_fwd(
      pre.data_ptr(), bias.data_ptr(), deter.data_ptr(), dst.data_ptr(), M, g, Dg3 // 3,
      _dtype_code(pre), deter.stride(0), dst.stride(0), stream))
  if out is not None and dst is not out:
    out.copy_(dst)
    return out
  return dst


@torch.no_grad()
def lambda_return(last, term, rew, boot, disc, lam):
  """dreamerv3/agent.py:482-490 for (rows, L) inputs -> (rows, L-1); one launch.
  The result only ever feeds stop-gradient targets, so there is no backward."""
  lib = _lib_bound()
  f = lambda x: x.detach().to(f32).contiguous()
  last, term, rew, boot = f(last), f(term), f(rew), f(boot)
  rows, L = rew.shape
  ret = torch.empty((rows, L - 1), dtype=f32, device=rew.device)
  stream = torch.cuda.current_stream(rew.device).cuda_stream
  _lib.check(lib.emb_lambda_return(
      last.data_ptr(), term.data_ptr(), rew.data_ptr(), boot.data_ptr(), ret.data_ptr(), rows, L,
      float(disc), float(lam), stream))
  return ret


@torch.no_grad()
def onehot_sample(logit, gumbel, unimix, out_dtype, out=None):
  """logit (n, S, C) fp32 / bf16, gumbel (n, S, C) fp32 -> one-hot (n, S, C) in
  `out_dtype`: the sampled value of outs.OneHot (no straight-through term).
  gumbel may be a row-strided view (one step of a (n, H, S, C) buffer); `out`, if
  given, is a (n, S*C) row-strided view that receives the sample."""
  lib = _lib_bound()
  n, S, C = logit.shape
  logit = logit.contiguous()
  gumbel = gumbel.to(f32)
  if not (gumbel.stride(2) == 1 and gumbel.stride(1) == C):
    gumbel = gumbel.contiguous()
  if out is None:
    res = torch.empty((n, S, C), dtype=out_dtype, device=logit.device)
    flat = res.view(n, S * C)
  else:
    assert out.shape == (n, S * C) and out.stride(1) == 1 and out.dtype == out_dtype
    res = flat = out
  stream = torch.cuda.current_stream(logit.device).cuda_stream
  _lib.check(lib.emb_onehot_sample(
      logit.data_ptr(), _dtype_code(logit), S * C, gumbel.data_ptr(), gumbel.stride(0), n, S, C,
      float(unimix), flat.data_ptr(), _dtype_code(flat), flat.stride(0), None, stream))
  return res


# ------------------------------------------- tcgen05 implicit-GEMM convolutions
def conv_tc_supported(x, cin, cout, k=5):
  """emb_conv5x5_nhwc_tc: bf16 NHWC, 64-channel K blocks, whole-row tiles of at most 128 pixels."""
  if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4 or k not in (1, 3, 5):
    return False
  n, h, w, c = x.shape
  # any width up to 128: tiles are whole rows / whole images of at most 128 pixels (csrc/conv_tc.cu tile_box)
  return not (c != cin or cin % 64 or cout % 32 or not 32 <= cout <= 256 or w > 128)


def pack_conv_weight(w, data_grad=False):
  """HWIO kernel (k, k, cin, cout) -> the bf16 [k*k][N][K] layout of emb_conv5x5_nhwc_tc.
  Forward: N = cout, K = cin.  data_grad: the spatially flipped kernel with the channel
  roles exchanged (N = cin, K = cout), so that the same launch computes the input gradient."""
  k, _, cin, cout = w.shape
  if data_grad:
    return w.flip(0, 1).reshape(k * k, cin, cout).to(torch.bfloat16).contiguous()
  return w.permute(0, 1, 3, 2).reshape(k * k, cout, cin).to(torch.bfloat16).contiguous()


def conv_tc(x, wp, bias=None, k=5):
  """x (N, H, W, Cin) bf16 NHWC, wp = pack_conv_weight(...) -> (N, H, W, Cout) bf16."""
  lib = _lib.load()
  if not getattr(lib, '_conv_tc_bound', False):
    lib.emb_conv5x5_nhwc_tc.argtypes = [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]
    lib.emb_conv5x5_nhwc_tc.restype = ctypes.c_int
    lib._conv_tc_bound = True
  x = x.contiguous()
  n, h, w, cin = x.shape
  taps, cout, kk = wp.shape
  assert taps == k * k and kk == cin and wp.dtype == torch.bfloat16 and wp.is_contiguous(), (wp.shape, x.shape)
  out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=x.device)
  stream = torch.cuda.current_stream(x.device).cuda_stream
  _lib.check(lib.emb_conv5x5_nhwc_tc(
      x.data_ptr(), wp.data_ptr(), None if bias is None else bias.data_ptr(), out.data_ptr(),
      n, h, w, cin, cout, k, stream))
  return out


class ConvTC(torch.autograd.Function):
  """y = conv_same(x, w): x (N, H, W, Cin) bf16 NHWC, w (k, k, Cin, Cout) HWIO
  (nets.py:298-323).  Forward and input gradient are the same tcgen05 launch with two
  weight packings; the weight gradient is `wgrad` below."""

  @staticmethod
  def forward(ctx, x, w):
    k = w.shape[0]
    x = x.contiguous()
    y = conv_tc(x, pack_conv_weight(w), k=k)
    ctx.save_for_backward(x, w)
    return y

  @staticmethod
  def backward(ctx, gy):
    x, w = ctx.saved_tensors
    k = w.shape[0]
    gy = gy.contiguous()
    gx = gw = None
    if ctx.needs_input_grad[0]:
      gx = conv_tc(gy, pack_conv_weight(w, data_grad=True), k=k)
    if ctx.needs_input_grad[1]:
      gw = conv_wgrad(x, gy, k).to(w.dtype)
    return gx, gw


def conv_wgrad_tc_supported(x, gy, k):
  if not (x.is_cuda and x.dtype == gy.dtype == torch.bfloat16 and x.dim() == 4 and k in (1, 3, 5)):
    return False
  n, h, w, cin = x.shape
  cout = gy.shape[-1]
  if not ((cin in (128, 256) and cout % 64 == 0 and 64 <= cout <= 256) or
          (cout in (128, 256) and cin % 64 == 0 and 64 <= cin <= 256)):
    return False
  return w <= 256            # any image size: chunks of whole rows / images (or row pieces) of <= 64 pixels


def conv_wgrad(x, gy, k):
  """dL/dw (k, k, Cin, Cout), fp32, of a SAME convolution from x (N, H, W, Cin) and
  gy (N, H, W, Cout), both NHWC: emb_conv5x5_wgrad_tc where the shape fits its tiles."""
  cin, cout = x.shape[-1], gy.shape[-1]
  if conv_wgrad_tc_supported(x, gy, k):
    lib = _lib.load()
    if not getattr(lib, '_conv_wgrad_bound', False):
      lib.emb_conv5x5_wgrad_tc.argtypes = [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
      lib.emb_conv5x5_wgrad_tc.restype = ctypes.c_int
      lib._conv_wgrad_bound = True
    x, gy = x.contiguous(), gy.contiguous()
    n, h, w, _ = x.shape
    m_is_in = cin in (128, 256)
    dw = torch.zeros((k * k, cin, cout) if m_is_in else (k * k, cout, cin), dtype=f32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_conv5x5_wgrad_tc(
        x.data_ptr(), gy.data_ptr(), dw.data_ptr(), n, h, w, cin, cout, k, int(m_is_in), stream))
    if not m_is_in:
      dw = dw.transpose(1, 2)
    return dw.reshape(k, k, cin, cout)
  g = torch.nn.grad.conv2d_weight(
      x.permute(0, 3, 1, 2), (cout, cin, k, k), gy.permute(0, 3, 1, 2), padding=k // 2)
  return g.permute(2, 3, 1, 0)


# ----------------------------- sub-pixel form of `nearest x2 up-sampling -> 5x5 conv`
class _ConvArgs(ctypes.Structure):
  _fields_ = [('inp', _vp), ('w_packed', _vp), ('bias', _vp), ('out', _vp), ('n', _i64),
              ('h', _i32), ('w', _i32), ('cin', _i32), ('cout', _i32), ('ksize', _i32),
              ('in_up', _i32), ('out_up', _i32), ('out_phase', _i32)]


class _WgradArgs(ctypes.Structure):
  _fields_ = [('x', _vp), ('gy', _vp), ('dw', _vp), ('n', _i64),
              ('h', _i32), ('w', _i32), ('cin', _i32), ('cout', _i32), ('ksize', _i32),
              ('m_is_in', _i32), ('gy_up', _i32), ('gy_phase', _i32)]


def _conv_general(**kw):
  lib = _lib.load()
  if not getattr(lib, '_conv_general_bound', False):
    lib.emb_conv_nhwc_tc.argtypes = [ctypes.POINTER(_ConvArgs), _vp]
    lib.emb_conv_nhwc_tc.restype = ctypes.c_int
    lib.emb_conv_wgrad_tc.argtypes = [ctypes.POINTER(_WgradArgs), _vp]
    lib.emb_conv_wgrad_tc.restype = ctypes.c_int
    lib._conv_general_bound = True
  stream = torch.cuda.current_stream().cuda_stream
  if 'dw' in kw:
    _lib.check(lib.emb_conv_wgrad_tc(ctypes.byref(_WgradArgs(**kw)), stream))
  else:
    _lib.check(lib.emb_conv_nhwc_tc(ctypes.byref(_ConvArgs(**kw)), stream))


_SUBPIXEL_FOLD = {}


def subpixel_fold(device):
  """(36, 25) 0/1 matrix F: Weff[(py, px, ty, tx)] = sum_{ky, kx} F[.., (ky, kx)] W[ky, kx].
  A 5x5 tap dy = ky - 2 applied at output row 2y + py of the up-sampled image reads
  low-resolution row y + floor((py + dy) / 2): ty = floor((py + ky - 2) / 2) + 1."""
  key = str(device)
  if key not in _SUBPIXEL_FOLD:
    F = torch.zeros((2, 2, 3, 3, 5, 5), dtype=f32)
    for py in range(2):
      for px in range(2):
        for ky in range(5):
          for kx in range(5):
            F[py, px, (py + ky - 2) // 2 + 1, (px + kx - 2) // 2 + 1, ky, kx] = 1.0
    _SUBPIXEL_FOLD[key] = F.reshape(36, 25).to(device)
  return _SUBPIXEL_FOLD[key]


def subpixel_supported(x, cin, cout):
  """x: the LOW-resolution input (N, h, w, cin) of an up-sample + 5x5 conv stage."""
  if not conv_tc_supported(x, cin, cout, 3) or cout % 64:       # cout is K of the data gradient
    return False
  n, h, w, _ = x.shape
  if not ((cin in (128, 256) and 64 <= cout <= 256) or (cout in (128, 256) and cin % 64 == 0 and 64 <= cin <= 256)):
    return False
  return w <= 128            # any map size: partially filled tiles / chunks (csrc/conv_tc.cu)


class SubpixelConv(torch.autograd.Function):
  """y (N, 2h, 2w, cout) = conv5x5_same(nearest_up2(x), w) from the low-resolution x (N, h, w, cin)
  and the folded kernels weff (4, 9, cin, cout) = subpixel_fold @ w: four 3x3 tcgen05 convolutions
  that write the four output phases, one launch for the input gradient (the reduction runs over
  the phases as well), four launches for the gradient of weff."""

  @staticmethod
  def forward(ctx, x, weff):
    x = x.contiguous()
    n, h, w, cin = x.shape
    cout = weff.shape[-1]
    wp = weff.permute(0, 1, 3, 2).to(torch.bfloat16).contiguous()          # (4, 9, cout, cin)
    y = torch.empty((n, 2 * h, 2 * w, cout), dtype=torch.bfloat16, device=x.device)
    for phase in range(4):
      _conv_general(inp=x.data_ptr(), w_packed=wp[phase].data_ptr(), bias=None, out=y.data_ptr(), n=n,
                    h=h, w=w, cin=cin, cout=cout, ksize=3, in_up=1, out_up=2, out_phase=phase)
    ctx.save_for_backward(x, weff)
    return y

  @staticmethod
  def backward(ctx, gy):
    x, weff = ctx.saved_tensors
    n, h, w, cin = x.shape
    cout = weff.shape[-1]
    gy = gy.contiguous()
    gx = gw = None
    if ctx.needs_input_grad[0]:
      wd = weff.flip(1).to(torch.bfloat16).contiguous()                     # (4, 9, cin, cout): N = cin, K = cout
      gx = torch.empty_like(x)
      _conv_general(inp=gy.data_ptr(), w_packed=wd.data_ptr(), bias=None, out=gx.data_ptr(), n=n, h=h,
                    w=w, cin=cout, cout=cin, ksize=3, in_up=2, out_up=1, out_phase=0)
    if ctx.needs_input_grad[1]:
      m_is_in = cin in (128, 256)
      gw = torch.zeros((4, 9, cin, cout) if m_is_in else (4, 9, cout, cin), dtype=f32, device=x.device)
      for phase in range(4):
        _conv_general(x=x.data_ptr(), gy=gy.data_ptr(), dw=gw[phase].data_ptr(), n=n, h=h, w=w, cin=cin,
                      cout=cout, ksize=3, m_is_in=int(m_is_in), gy_up=2, gy_phase=phase)
      if not m_is_in:
        gw = gw.transpose(2, 3)
      gw = gw.to(weff.dtype)
    return gx, gw


def upconv_subpixel(x, w):
  """nearest x2 up-sampling followed by the SAME 5x5 convolution with HWIO kernel w, computed on
  the low-resolution grid (dreamerv3/rssm.py:336-340)."""
  k, _, cin, cout = w.shape
  assert k == 5
  weff = (subpixel_fold(w.device) @ w.reshape(25, cin * cout).to(f32)).reshape(4, 9, cin, cout)
  return SubpixelConv.apply(x, weff)


# ------------------------------------------------------------------ head losses
def _losses_lib():
  lib = _lib.load()
  if not getattr(lib, '_losses_bound', False):
    lib.emb_twohot_loss_fwd.argtypes = [_vp, _vp, _vp, _fl, _vp, _vp, _vp, _i64, _i32, _vp]
    lib.emb_twohot_loss_fwd.restype = ctypes.c_int
    lib.emb_twohot_loss_bwd.argtypes = [_vp, _vp, _vp, _fl, _vp, _vp, _vp, _vp, _i64, _i32, _vp]
    lib.emb_twohot_loss_bwd.restype = ctypes.c_int
    lib.emb_twohot_pred.argtypes = [_vp, _vp, _vp, _i64, _i32, _vp]
    lib.emb_twohot_pred.restype = ctypes.c_int
    lib.emb_loss_reduce.argtypes = [ctypes.POINTER(_vp), ctypes.POINTER(_i64), ctypes.POINTER(_fl), _i32,
                                    _vp, _vp, _vp, _vp]
    lib.emb_loss_reduce.restype = ctypes.c_int
    lib._losses_bound = True
  return lib


def twohot_supported(logits):
  return logits.is_cuda and logits.dtype == f32 and 2 <= logits.shape[-1] <= 1024


class TwoHotLoss(torch.autograd.Function):
  """outs.TwoHot.loss (embodied/jax/outs.py:311-330) for fp32 logits (..., nbins) and one or two
  stop-gradient targets (...): CE(twohot(t1)) + w2 * CE(twohot(t2)), one launch each way."""

  @staticmethod
  def forward(ctx, logits, bins, target, target2, w2):
    lib = _losses_lib()
    logits = logits.contiguous()
    lead, nb = logits.shape[:-1], logits.shape[-1]
    rows = logits.numel() // nb
    t1 = target.detach().to(f32).contiguous()
    t2 = None if target2 is None else target2.detach().to(f32).contiguous()
    assert t1.numel() == rows and (t2 is None or t2.numel() == rows), (logits.shape, target.shape)
    loss = torch.empty(lead, dtype=f32, device=logits.device)
    lse = torch.empty(lead, dtype=f32, device=logits.device)
    stream = torch.cuda.current_stream(logits.device).cuda_stream
    _lib.check(lib.emb_twohot_loss_fwd(
        logits.data_ptr(), t1.data_ptr(), None if t2 is None else t2.data_ptr(), float(w2),
        bins.data_ptr(), loss.data_ptr(), lse.data_ptr(), rows, nb, stream))
    ctx.save_for_backward(logits, bins, t1, t2, lse)
    ctx.w2 = float(w2)
    return loss

  @staticmethod
  def backward(ctx, gloss):
    lib = _losses_lib()
    logits, bins, t1, t2, lse = ctx.saved_tensors
    nb = logits.shape[-1]
    rows = logits.numel() // nb
    gloss = gloss.to(f32).contiguous()
    glogits = torch.empty_like(logits)
    stream = torch.cuda.current_stream(logits.device).cuda_stream
    _lib.check(lib.emb_twohot_loss_bwd(
        logits.data_ptr(), t1.data_ptr(), None if t2 is None else t2.data_ptr(), ctx.w2, bins.data_ptr(),
        lse.data_ptr(), gloss.data_ptr(), glogits.data_ptr(), rows, nb, stream))
    return glogits, None, None, None, None


def twohot_loss(logits, bins, target, target2=None, w2=0.0):
  return TwoHotLoss.apply(logits, bins, target, target2, w2)


@torch.no_grad()
def twohot_pred(logits, bins):
  """outs.TwoHot.pred (outs.py:285-309); only ever feeds stop-gradient quantities."""
  lib = _losses_lib()
  logits = logits.contiguous()
  nb = logits.shape[-1]
  pred = torch.empty(logits.shape[:-1], dtype=f32, device=logits.device)
  stream = torch.cuda.current_stream(logits.device).cuda_stream
  _lib.check(lib.emb_twohot_pred(logits.data_ptr(), bins.data_ptr(), pred.data_ptr(),
                                 logits.numel() // nb, nb, stream))
  return pred


_TICKETS = {}


class LossSum(torch.autograd.Function):
  """total = sum_i scale_i * mean(term_i) and the per-term means (dreamerv3/agent.py:237-240) in
  one launch; the backward pass is one constant fill per term."""

  @staticmethod
  def forward(ctx, scales, *terms):
    lib = _losses_lib()
    dev = terms[0].device
    terms = [t.to(f32).contiguous() for t in terms]
    n = len(terms)
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    if key not in _TICKETS:
      _TICKETS[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty(n + 1, dtype=f32, device=dev)
    ptrs = (_vp * n)(*[t.data_ptr() for t in terms])
    counts = (_i64 * n)(*[t.numel() for t in terms])
    sc = (_fl * n)(*[float(s) for s in scales])
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.emb_loss_reduce(ptrs, counts, sc, n, out.data_ptr(), out[n:].data_ptr(),
                                   _TICKETS[key].data_ptr(), stream))
    ctx.shapes = [t.shape for t in terms]
    ctx.scales = [float(s) for s in scales]
    ctx.mark_non_differentiable(out[:n])
    return out[n], out[:n]

  @staticmethod
  def backward(ctx, gtotal, _gmeans):
    grads = []
    for shape, scale in zip(ctx.shapes, ctx.scales):
      numel = 1
      for s in shape:
        numel *= s
      grads.append((gtotal * (scale / numel)).expand(shape))
    return (None, *grads)


def loss_sum(losses, scales):
  """losses: {name: tensor}; returns (total, {name: mean})."""
  names = list(losses)
  total, means = LossSum.apply([scales[k] for k in names], *[losses[k] for k in names])
  return total, {k: means[i] for i, k in enumerate(names)}
