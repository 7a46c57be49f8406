"""torch.autograd wrappers of the element-wise / reduction kernels in
libembodied_b200.so (include/embodied_b200.h)."""
import ctypes

import torch

from .. import _lib

f32 = torch.float32
_vp, _i32, _i64, _fl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
_bound = False


def _lib_bound():
  global _bound
  lib = _lib.load()
  if not _bound:
    lib.emb_rmsnorm_act_fwd.argtypes = [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _fl, _vp]
    lib.emb_rmsnorm_act_fwd.restype = ctypes.c_int
    lib.emb_rmsnorm_act_bwd.argtypes = [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _fl, _vp]
    lib.emb_rmsnorm_act_bwd.restype = ctypes.c_int
    for name in ('emb_maxpool2_nhwc_fwd', 'emb_maxpool2_nhwc_bwd'):
      getattr(lib, name).argtypes = [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]
      getattr(lib, name).restype = ctypes.c_int
    for name in ('emb_upsample2_nhwc_fwd', 'emb_upsample2_nhwc_bwd'):
      getattr(lib, name).argtypes = [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]
      getattr(lib, name).restype = ctypes.c_int
    _bound = True
  return lib


def _dtype_code(t):
  return {torch.float32: 0, torch.bfloat16: 1}[t.dtype]


def rmsnorm_supported(x, need_grad):
  if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
    return False
  cols = x.shape[-1]
  per = 4 if x.dtype == torch.float32 else 8
  if cols % per or (need_grad and cols > 2048):
    return False
  return True


class RmsNormAct(torch.autograd.Function):
  """y = act(rms_norm(x) * scale) over the last axis (nets.py:361-399 + act)."""

  @staticmethod
  def forward(ctx, x, scale, act, eps):
    lib = _lib_bound()
    x = x.contiguous()
    y = torch.empty_like(x)
    cols = x.shape[-1]
    rows = x.numel() // cols
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_rmsnorm_act_fwd(
        x.data_ptr(), scale.data_ptr(), y.data_ptr(), rows, cols, _dtype_code(x), int(act),
        eps, stream))
    ctx.save_for_backward(x, scale)
    ctx.act, ctx.eps = int(act), eps
    return y

  @staticmethod
  def backward(ctx, gy):
    lib = _lib_bound()
    x, scale = ctx.saved_tensors
    gy = gy.contiguous()
    gx = torch.empty_like(x)
    gscale = torch.zeros_like(scale)
    cols = x.shape[-1]
    rows = x.numel() // cols
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.emb_rmsnorm_act_bwd(
        x.data_ptr(), scale.data_ptr(), gy.data_ptr(), gx.data_ptr(), gscale.data_ptr(),
        rows, cols, _dtype_code(x), ctx.act, ctx.eps, stream))
    return gx, gscale, None, None


def rmsnorm_act(x, scale, act=True, eps=1e-4):
  return RmsNormAct.apply(x, scale, act, eps)


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
