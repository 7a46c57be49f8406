"""Host side of the fused RSSM scan kernels (include/embodied_b200.h
`emb_rssm_observe_fwd` / `emb_rssm_observe_bwd`; embodied_b200/csrc/rssm_*.cu).

`pack(store, engine)` lays the six in-scan weight matrices out in the order the
kernel streams them (mma B-fragment order for the bf16 engine) -- once per
optimiser step.  `observe(...)` is the autograd function the model calls in
place of the per-step torch loop; everything that does not depend on the
recurrent state (action branch, token half of obs0, step 0's dynin0/dynin1) is
computed by the caller with ordinary (B*T)-row GEMMs and passed in.
"""
import ctypes

import torch

from .. import _lib

f32 = torch.float32
ROWS = 16
ENG_F32, ENG_BF16 = 0, 1

_vp, _i32, _fl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float


class FwdArgs(ctypes.Structure):
  _fields_ = (
      [(n, _i32) for n in ('B', 'T', 'D', 'H', 'S', 'C', 'G', 'engine')] +
      [('unimix', _fl), ('eps', _fl)] +
      [(n, _vp) for n in (
          'w_ph1', 'w_logit', 'w_hid', 'w_gru', 'w_in1',
          'b0', 'b1', 'b_hid', 'b_gru', 'b_logit', 's0', 's1', 's_hid', 's_obs',
          'deter0', 'x2', 'pre_tok', 'keep', 'gumbel',
          'deter', 'logit', 'index',
          'y0', 'y1', 'yhid', 'gates', 'yobs', 'sumsq', 'deterA', 'barrier')])


def _bind(lib):
  if getattr(lib, '_rssm_bound', False):
    return
  lib.emb_rssm_observe_fwd.argtypes = [ctypes.POINTER(FwdArgs), _vp]
  lib.emb_rssm_observe_fwd.restype = ctypes.c_int
  lib._rssm_bound = True


def pack_matrix(w, engine):
  """w: (..., K, N) fp32 -> the engine's streaming layout (see rssm_common.cuh)."""
  *lead, K, N = w.shape
  assert K % 16 == 0 and N % 8 == 0, (K, N)
  if engine == ENG_BF16:
    n = len(lead)
    x = w.reshape(*lead, K // 16, 2, 4, 2, N // 8, 8)      # kstep, reg, kq, half, tile, nn
    x = x.permute(*range(n), n, n + 4, n + 5, n + 2, n + 1, n + 3)
    return x.to(torch.bfloat16).contiguous()
  x = w.reshape(*lead, K, N // 8, 8).transpose(-3, -2)
  return x.contiguous()


@torch.no_grad()
def pack(store, cfg, engine):
  """The packed copies of the in-scan weights (dreamerv3/rssm.py:135-159, 81-86)."""
  D, G = cfg.deter, cfg.blocks
  Dg = D // G
  m = lambda n: store.view('master', n)
  wobs = m('dyn/obs0/kernel')
  gru = m('dyn/dyngru/kernel')                              # (G, Dg, 3*Dg), columns (gate, j)
  gru = gru.reshape(G, Dg, 3, Dg // 8, 8).permute(0, 1, 3, 2, 4).reshape(G, Dg, 3 * Dg)
  cd = torch.bfloat16 if engine == ENG_BF16 else f32
  return dict(
      w_ph1=pack_matrix(torch.cat([wobs[:D], m('dyn/dynin0/kernel')], 1), engine),
      w_logit=pack_matrix(m('dyn/obslogit/kernel'), engine),
      w_hid=pack_matrix(m('dyn/dynhid0/kernel'), engine),
      w_gru=pack_matrix(gru, engine),
      w_in1=m('dyn/dynin1/kernel').to(cd).contiguous())


def time_major(x, B):
  """(B, T, ...) -> contiguous fp32 (T, 16, ...), rows >= B zero."""
  T = x.shape[1]
  out = torch.zeros((T, ROWS, *x.shape[2:]), dtype=f32, device=x.device)
  out[:, :B] = x.transpose(0, 1)
  return out


def rows16(x, B):
  out = torch.zeros((ROWS, *x.shape[1:]), dtype=f32, device=x.device)
  out[:B] = x
  return out


class Scan:
  """Launches the scan kernels for one model (weights from `store`)."""

  def __init__(self, cfg, store, engine):
    self.cfg, self.store, self.engine = cfg, store, engine
    self.lib = _lib.load()
    _bind(self.lib)
    self.packed = None
    self.packed_step = -1

  def weights(self):
    if self.packed is None or self.packed_step != self.store.version:
      self.packed = pack(self.store, self.cfg, self.engine)
      self.packed_step = self.store.version
    return self.packed

  @torch.no_grad()
  def forward(self, deter0, y0, y1, x2, pre_tok, keep, gumbel):
    """All inputs batch-major: deter0 (B,D), y0/y1 (B,H) [step 0, pre-norm],
    x2/pre_tok (B,T,H), keep (B,T) float, gumbel (B,T,S,C).  Returns
    deter (B,T,D), logit (B,T,S,C), index (B,T,S) and the saved activations."""
    cfg, dev = self.cfg, deter0.device
    B, T = keep.shape
    D, H, S, C, G = cfg.deter, cfg.hidden, cfg.stoch, cfg.classes, cfg.blocks
    w = self.weights()
    m = lambda n: self.store.view('master', n)
    z = lambda *s: torch.empty(s, dtype=f32, device=dev)
    keep_tm = torch.ones((T + 1, ROWS), dtype=f32, device=dev)
    keep_tm[:T, :B] = keep.transpose(0, 1)
    sv = dict(
        deter0=rows16(deter0.to(f32), B), x2=time_major(x2, B),
        pre_tok=time_major(pre_tok, B), keep=keep_tm,
        gumbel=time_major(gumbel.reshape(B, T, S * C), B),
        deter=z(T, ROWS, D), logit=z(T, ROWS, S * C),
        index=torch.empty((T, ROWS, S), dtype=torch.int32, device=dev),
        y0=z(T + 1, ROWS, H), y1=z(T + 1, ROWS, H), yhid=z(T, ROWS, D),
        gates=z(T, 3, ROWS, D), yobs=z(T, ROWS, H),
        sumsq=torch.zeros((T, ROWS), dtype=f32, device=dev),
        deterA=torch.empty((2, ROWS * D), dtype=torch.bfloat16, device=dev),
        barrier=torch.zeros(4, dtype=torch.int32, device=dev))
    sv['y0'][0] = rows16(y0.to(f32), B)
    sv['y1'][0] = rows16(y1.to(f32), B)
    vec = dict(
        b0=m('dyn/dynin0/bias'), b1=m('dyn/dynin1/bias'), b_hid=m('dyn/dynhid0/bias'),
        b_gru=m('dyn/dyngru/bias'), b_logit=m('dyn/obslogit/bias'),
        s0=m('dyn/dynin0norm/scale'), s1=m('dyn/dynin1norm/scale'),
        s_hid=m('dyn/dynhid0norm/scale'), s_obs=m('dyn/obs0norm/scale'))
    args = FwdArgs(B=B, T=T, D=D, H=H, S=S, C=C, G=G, engine=self.engine,
                   unimix=cfg.unimix, eps=1e-4)
    for k, v in {**w, **vec, **sv}.items():
      assert v.is_contiguous(), k
      setattr(args, k, v.data_ptr())
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(self.lib.emb_rssm_observe_fwd(ctypes.byref(args), stream))
    out = dict(
        deter=sv['deter'][:, :B].transpose(0, 1),
        logit=sv['logit'][:, :B].transpose(0, 1).reshape(B, T, S, C),
        index=sv['index'][:, :B].transpose(0, 1))
    return out, sv
