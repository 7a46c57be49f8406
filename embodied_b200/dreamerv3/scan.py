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
import math

import numpy as np
import torch

from .. import _lib

f32 = torch.float32
# bench.py sets this to a list to collect (name, start, end) CUDA event triples
# around the scan launches (live per-launch timing for the roofline line).
PROFILE = None
# name -> _lib.Stopwatch recorded around the scan launches INSIDE a captured
# train step (bench.py sets GRAPH_TIMERS = {} before the agent captures).
GRAPH_TIMERS = None
ROWS = 16
ENG_F32, ENG_BF16, ENG_LEGACY = 0, 1, 2     # fp32 FFMA | bf16 mma + TMA ring | bf16 register-staged

_vp, _i32, _fl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float


class FwdArgs(ctypes.Structure):
  _fields_ = (
      [(n, _i32) for n in ('B', 'T', 'D', 'H', 'S', 'C', 'G', 'engine', 'ncta', 'tma_cfg')] +
      [('unimix', _fl), ('eps', _fl)] +
      [(n, _vp) for n in (
          'w_ph1', 'w_logit', 'w_hid', 'w_gru', 'w_in1',
          'b0', 'b1', 'b_hid', 'b_gru', 'b_logit', 's0', 's1', 's_hid', 's_obs',
          'deter0', 'x2', 'pre_tok', 'keep', 'gumbel',
          'deter', 'logit', 'index',
          'y0', 'y1', 'yhid', 'gates', 'yobs', 'sumsq', 'probs', 'rstd', 'deterA', 'barrier', 'timing',
          'hid_pre', 'sumsq_obs')])


def _graph_timer(name):
  if GRAPH_TIMERS is None or not torch.cuda.is_current_stream_capturing():
    return None
  watch = GRAPH_TIMERS.get(name)
  if watch is None:
    watch = GRAPH_TIMERS[name] = _lib.Stopwatch()
  return watch


def _bind(lib):
  if getattr(lib, '_rssm_bound', False):
    return
  lib.emb_rssm_observe_fwd.argtypes = [ctypes.POINTER(FwdArgs), _vp]
  lib.emb_rssm_observe_fwd.restype = ctypes.c_int
  lib._rssm_bound = True


def tile_assignment(tiles, ncta, unit=1, groups=1):
  """Which n8 column tiles each CTA owns: (per, ids[ncta][per]) with -1 = padding.

  groups == 1: CTA c owns the run [c*per, (c+1)*per), per = ceil(tiles/unit/ncta)*unit.
  groups  > 1 (block-diagonal layers): every CTA stays inside ONE group, so it
  builds a single A operand per phase: ncta // groups CTAs per group, each a run
  of per = ceil(tiles_per_group/unit / ctas_per_group)*unit tiles of that group.
  The kernels (rssm_fwd.cu / rssm_bwd.cu) use the same formulas."""
  ids = []
  if groups == 1:
    per = -(-(tiles // unit) // ncta) * unit
    for c in range(ncta):
      run = list(range(c * per, min(tiles, (c + 1) * per)))
      ids.append(run + [-1] * (per - len(run)))
    return per, ids
  tpg = tiles // groups
  cpg = max(1, ncta // groups)
  per = -(-(tpg // unit) // cpg) * unit
  for c in range(ncta):
    g, j = divmod(c, cpg)
    run = list(range(g * tpg + j * per, g * tpg + min(tpg, (j + 1) * per))) if g < groups else []
    ids.append(run + [-1] * (per - len(run)))
  return per, ids


_IDS = {}


def tile_groups(per):
  """rssm_tma.cuh tile_groups(): tile groups the 8 consumer warps form."""
  tg = 1
  while tg < 8 and tg * 4 <= per:
    tg *= 2
  return tg


def pad_tiles(per, unit=1):
  """rssm_tma.cuh pad_tiles(): tiles per CTA block of the TMA engine."""
  while per % tile_groups(per):
    per += unit
  return per


def pack_matrix(w, engine, ncta, unit=1, groups=1):
  """w: (K, N) fp32 -> the engine's streaming layout (rssm_common.cuh).

  bf16: every CTA's n8 column tiles (see tile_assignment) are stored as one
  contiguous block [K/16][per][32 lanes][4 bf16] in mma.m16n8k16 B-fragment
  order: element (k, n) of a 16 x 8 tile sits in lane (n%8)*4 + (k%8)/2,
  register k/8, half k%2.
  fp32: [N/8][K][8]."""
  K, N = w.shape
  assert K % 16 == 0 and N % 8 == 0, (K, N)
  if engine == ENG_F32:
    return w.reshape(K, N // 8, 8).transpose(0, 1).contiguous()
  key = (N // 8, ncta, unit, groups, engine, str(w.device))
  hit = _IDS.get(key)
  if hit is None:                      # host -> device once (never inside a stream capture)
    per, ids = tile_assignment(N // 8, ncta, unit, groups)
    if engine == ENG_BF16:          # TMA engine: zero tiles up to a multiple of the tile groups
      padded = pad_tiles(per, unit)
      ids = [row + [-1] * (padded - per) for row in ids]
      per = padded
    hit = _IDS[key] = (per, torch.tensor(ids, dtype=torch.long, device=w.device))
  per, ids = hit                                                      # (ncta, per)
  wt = torch.cat([w.reshape(K, N // 8, 8).to(torch.bfloat16),
                  torch.zeros((K, 1, 8), dtype=torch.bfloat16, device=w.device)], 1)
  x = wt[:, ids.reshape(-1)]                                          # (K, ncta*per, 8); -1 -> zero tile
  x = x.reshape(K // 16, 2, 4, 2, ncta, per, 8)           # kstep, reg, kq, half, cta, tile, nn
  return x.permute(4, 0, 5, 6, 2, 1, 3).contiguous()      # cta, kstep, tile, nn, kq, reg, half


# ---------------------------------------------------------------- fused packing
FUSED_PACK = True      # False: the op-by-op torch formulation below (tests compare the two)
# bf16 engines, inside a backward pass: the three large weight-gradient GEMMs of the scan
# (dynhid0, dyngru, obs0[:D]) accumulate in fp32 STRAIGHT INTO the flat gradient buffer
# (baddbmm with beta = 1 on the tensor's view) instead of bf16 result -> cast -> AccumulateGrad add.
DIRECT_WGRAD = True
# The bf16 engines pack straight from the flat fp32 parameter buffer with one
# kernel per matrix (emb_pack_tiles, csrc/pack.cu).  A logical (K, N) matrix is
# described per n8 column tile by (element offset of (k=0, n=8t), k stride,
# n stride) -- block-diagonal layers, gate-interleaved column orders, column
# concatenations and transposes are all just different tables.

def _tiles_rowmajor(off, N, cols=None, first=0):
  """Columns [first, first+cols) of a row-major (K, N) tensor at `off`."""
  cols = N if cols is None else cols
  t = np.arange(cols // 8, dtype=np.int64)
  return off + first + 8 * t, np.full(len(t), N, np.int32), np.ones(len(t), np.int32)


def _tiles_transposed(off, rows, K):
  """The transpose of a row-major (rows, K) tensor at `off`: logical (K, rows)."""
  t = np.arange(rows // 8, dtype=np.int64)
  return off + 8 * t * K, np.ones(len(t), np.int32), np.full(len(t), K, np.int32)


def _cat_tiles(*parts):
  return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))


def slot_tables(segments, engine, ncta, unit=1, groups=1):
  """Per CTA slot (cta*per + tile) the (offset, k stride, n stride) of its tile,
  offset -1 for padding: (per, [(first k-step, k-steps, off, kstride, nstride)])."""
  tiles = len(segments[0][2][0])
  per, ids = tile_assignment(tiles, ncta, unit, groups)
  if engine == ENG_BF16:          # TMA engine: zero tiles up to a multiple of the tile groups
    padded = pad_tiles(per, unit)
    ids = [row + [-1] * (padded - per) for row in ids]
    per = padded
  ids = np.asarray(ids, np.int64).reshape(-1)
  real = np.maximum(ids, 0)
  tabs = []
  for first, count, (off, ks, ns) in segments:
    assert len(off) == tiles, (len(off), tiles)
    tabs.append((first, count, np.where(ids >= 0, off[real], -1).astype(np.int64),
                 ks[real].astype(np.int32), ns[real].astype(np.int32)))
  return per, tabs


def pack_tiles(store, name, ksteps, segments, engine, ncta, unit=1, groups=1):
  """segments: [(first k-step, k-steps, (off, kstride, nstride) per logical tile)];
  every segment covers all N/8 tiles.  Returns the same bytes as pack_matrix()."""
  lib = _lib.load()
  dev = store.master.device
  tables = store.__dict__.setdefault('_pack_tables', {})    # offsets belong to this store
  key = (name, ncta, unit, groups, engine)
  hit = tables.get(key)
  if hit is None:                      # host -> device once (never inside a stream capture)
    per, tabs = slot_tables(segments, engine, ncta, unit, groups)
    tabs = [(first, count) + tuple(torch.from_numpy(x).to(dev) for x in rest)
            for first, count, *rest in tabs]
    lib.emb_pack_tiles.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64] + [ctypes.c_int32] * 4 + [_vp]
    lib.emb_pack_tiles.restype = ctypes.c_int
    hit = tables[key] = (per, tabs)
  per, tabs = hit
  dst = torch.empty((ncta, ksteps, per, 8, 4, 2, 2), dtype=torch.bfloat16, device=dev)
  stream = torch.cuda.current_stream(dev).cuda_stream
  for first, count, off, ks, ns in tabs:
    _lib.check(lib.emb_pack_tiles(
        store.master.data_ptr(), off.data_ptr(), ks.data_ptr(), ns.data_ptr(), dst.data_ptr(),
        ncta * per, per, first, count, ksteps, stream))
  return dst


def _pack_fused(store, cfg, engine, ncta, pack_tiles=pack_tiles):
  D, G, H, SC = cfg.deter, cfg.blocks, cfg.hidden, cfg.stoch * cfg.classes
  Dg = D // G
  o = store.offsets
  Kh = store.specs['dyn/dynhid0/kernel'][0][1]
  keep = Dg + 2 * H if engine == ENG_BF16 else Kh     # the TMA engine hoists the action rows
  g = np.arange(G, dtype=np.int64)
  # dynhid0 (G, Kh, Dg) -> (keep, D): column g*Dg + j
  hid = (np.repeat(o['dyn/dynhid0/kernel'] + g * Kh * Dg, Dg // 8) + np.tile(8 * np.arange(Dg // 8), G),
         np.full(D // 8, Dg, np.int32), np.ones(D // 8, np.int32))
  # dyngru (G, Dg, 3Dg) -> (Dg, 3D): tiles ordered (group, j/8, gate)
  gg, j8, gate = np.meshgrid(g, np.arange(Dg // 8), np.arange(3), indexing='ij')
  gru = ((o['dyn/dyngru/kernel'] + gg * Dg * 3 * Dg + gate * Dg + 8 * j8).reshape(-1).astype(np.int64),
         np.full(3 * D // 8, 3 * Dg, np.int32), np.ones(3 * D // 8, np.int32))
  ph1 = _cat_tiles(_tiles_rowmajor(o['dyn/obs0/kernel'], H), _tiles_rowmajor(o['dyn/dynin0/kernel'], H))
  logit = _tiles_rowmajor(o['dyn/obslogit/kernel'], SC)
  pk = lambda name, K, tiles, **kw: pack_tiles(store, name, K // 16, [(0, K // 16, tiles)], engine, ncta, **kw)
  return dict(
      w_ph1=pk('w_ph1', D, ph1), w_logit=pk('w_logit', H, logit),
      w_hid=pk('w_hid', keep, hid, groups=G), w_gru=pk('w_gru', Dg, gru, unit=3, groups=G))


def _pack_bwd_fused(store, cfg, engine, ncta, pack_tiles=pack_tiles):
  D, G, H, SC = cfg.deter, cfg.blocks, cfg.hidden, cfg.stoch * cfg.classes
  Dg = D // G
  o = store.offsets
  Kh = store.specs['dyn/dynhid0/kernel'][0][1]
  keep = Dg + 2 * H
  g = np.arange(G, dtype=np.int64)
  # gru_t (3Dg, D): element (k, g*Dg + i) = dyngru[g, i, k]
  gi, i8 = np.meshgrid(g, np.arange(Dg // 8), indexing='ij')
  gru_t = ((o['dyn/dyngru/kernel'] + gi * Dg * 3 * Dg + 8 * i8 * 3 * Dg).reshape(-1).astype(np.int64),
           np.ones(D // 8, np.int32), np.full(D // 8, 3 * Dg, np.int32))
  # hid_t (Dg, G*keep): element (k, g*keep + r) = dynhid0[g, r, k]
  gr, r8 = np.meshgrid(g, np.arange(keep // 8), indexing='ij')
  hid_t = ((o['dyn/dynhid0/kernel'] + gr * Kh * Dg + 8 * r8 * Dg).reshape(-1).astype(np.int64),
           np.ones(G * keep // 8, np.int32), np.full(G * keep // 8, Dg, np.int32))
  pk = lambda name, K, tiles, **kw: pack_tiles(store, name, K // 16, [(0, K // 16, tiles)], engine, ncta, **kw)
  # ph1_t (2H, D): rows [0, H) from obs0[:D], rows [H, 2H) from dynin0
  wt_ph1 = pack_tiles(store, 'wt_ph1', 2 * H // 16, [
      (0, H // 16, _tiles_transposed(o['dyn/obs0/kernel'], D, H)),
      (H // 16, H // 16, _tiles_transposed(o['dyn/dynin0/kernel'], D, H))], engine, ncta)
  return dict(
      wt_in1=pk('wt_in1', H, _tiles_transposed(o['dyn/dynin1/kernel'], SC, H),
                unit=(math.lcm(cfg.classes, 8) // 8 if engine == ENG_BF16 else 1)),
      wt_logit=pk('wt_logit', SC, _tiles_transposed(o['dyn/obslogit/kernel'], H, SC)),
      wt_ph1=wt_ph1,
      wt_gru=pk('wt_gru', 3 * Dg, gru_t, groups=G),
      wt_hid=pk('wt_hid', Dg, hid_t, groups=G))


@torch.no_grad()
def pack(store, cfg, engine, ncta):
  """The packed copies of the in-scan weights (dreamerv3/rssm.py:135-159, 81-86).
  The block-diagonal layers become one (K, N) matrix whose column decides the
  group: dynhid0 -> (D/G+3H, D); dyngru -> (D/G, 3D) with columns ordered
  (group, j/8, gate, j%8) so the three gates of a deter column are adjacent tiles."""
  D, G = cfg.deter, cfg.blocks
  Dg = D // G
  m = lambda n: store.view('master', n)
  wobs = m('dyn/obs0/kernel')
  hid = m('dyn/dynhid0/kernel')                             # (G, Kh, Dg)
  extra = {}
  if engine == ENG_BF16:
    # the action branch leaves the scan: rows [deter_g | x0 | x1] stay in the
    # streamed block, rows of x2 become one (B*T)-row GEMM (hid_pre)
    keep_rows = Dg + 2 * cfg.hidden
    extra['w_hid_x2'] = hid[:, keep_rows:].permute(1, 0, 2).reshape(-1, D).to(torch.bfloat16)
    hid = hid[:, :keep_rows]
  cd = f32 if engine == ENG_F32 else torch.bfloat16
  if engine != ENG_F32 and FUSED_PACK:
    return dict(**extra, **_pack_fused(store, cfg, engine, ncta),
                w_in1=m('dyn/dynin1/kernel').to(cd).contiguous())
  # the torch formulation of the same layouts (fp32 parity engine, FUSED_PACK off)
  hid = hid.permute(1, 0, 2).reshape(hid.shape[1], D)
  gru = m('dyn/dyngru/kernel')                              # (G, Dg, 3*Dg), columns (gate, j)
  gru = gru.reshape(G, Dg, 3, Dg // 8, 8).permute(1, 0, 3, 2, 4).reshape(Dg, 3 * D)
  return dict(
      **extra,
      w_ph1=pack_matrix(torch.cat([wobs[:D], m('dyn/dynin0/kernel')], 1), engine, ncta),
      w_logit=pack_matrix(m('dyn/obslogit/kernel'), engine, ncta),
      w_hid=pack_matrix(hid, engine, ncta, groups=G),
      w_gru=pack_matrix(gru, engine, ncta, unit=3, groups=G),
      w_in1=m('dyn/dynin1/kernel').to(cd).contiguous())


def time_major(x, B):
  """(B, T, ...) -> contiguous fp32 (T, 16, ...), rows >= B zero."""
  T = x.shape[1]
  out = torch.zeros((T, ROWS, *x.shape[2:]), dtype=f32, device=x.device)
  out[:, :B] = x.transpose(0, 1)
  return out


def a_fragments(x):
  """fp32 (T, 16, K) -> bf16 mma.m16n8k16 A fragments (T, K/16, 32 lanes, 4 regs, 2):
  element (r, k) in lane (r%8)*4 + (k%8)/2, register r/8 + 2*((k%16)/8), half k%2."""
  T, R, K = x.shape
  v = x.reshape(T, 2, 8, K // 16, 2, 4, 2)            # rhi, rlo, ks, khi, kq, half
  v = v.permute(0, 3, 2, 5, 4, 1, 6)                  # ks, rlo, kq, khi, rhi, half
  return v.to(torch.bfloat16).contiguous()


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
    self.packed_bwd = None
    self.packed_bwd_step = -1
    self.timing = False
    self.supported = True       # False: the model is too wide for either fused engine
    self.ncta = int(self.lib.emb_device_sm_count())
    if self.ncta <= 0:
      _lib.check(self.ncta)
    if self.engine == ENG_BF16:
      # very wide models leave no shared memory for the TMA ring next to the operands:
      # the register-staged bf16 kernels (engine 2) take over
      self.lib.emb_rssm_tma_fits.argtypes = [_i32] * 6
      self.lib.emb_rssm_tma_fits.restype = ctypes.c_int
      dims = (cfg.deter, cfg.hidden, cfg.stoch, cfg.classes, cfg.blocks, self.ncta)
      self.lib.emb_rssm_legacy_fits.argtypes = [_i32] * 6
      self.lib.emb_rssm_legacy_fits.restype = ctypes.c_int
      if not self.lib.emb_rssm_tma_fits(*dims):
        self.engine = ENG_LEGACY
        self.supported = bool(self.lib.emb_rssm_legacy_fits(*dims))

  def invalidate(self):
    """Forget the packed weight copies (the next use re-packs; a CUDA-graph
    capture calls this so that the packing is recorded inside the graph)."""
    self.packed = self.packed_bwd = None

  def weights(self):
    if self.packed is None or self.packed_step != self.store.version:
      self.packed = pack(self.store, self.cfg, self.engine, self.ncta)
      self.packed_step = self.store.version
    return self.packed

  def relaunch(self, args):
    """Launch again on the same buffers (micro-benchmarks)."""
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(self.lib.emb_rssm_observe_fwd(ctypes.byref(args), stream))

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
        index=torch.zeros((T, ROWS, S), dtype=torch.int32, device=dev),
        y0=torch.zeros((T + 1, ROWS, H), dtype=f32, device=dev),
        y1=torch.zeros((T + 1, ROWS, H), dtype=f32, device=dev), yhid=z(T, ROWS, D),
        probs=torch.zeros((T, ROWS, S * C), dtype=f32, device=dev),
        rstd=torch.zeros((T + 1, 3, ROWS), dtype=f32, device=dev),
        gates=z(T, 4, ROWS, D), yobs=z(T, ROWS, H),
        sumsq=torch.zeros((T, ROWS), dtype=f32, device=dev),
        sumsq_obs=torch.zeros((T, ROWS), dtype=f32, device=dev),
        deterA=torch.zeros(2 * ROWS * D + 2 * ROWS * H, dtype=torch.bfloat16, device=dev),
        barrier=torch.zeros(64, dtype=torch.int32, device=dev))      # [0] grid barrier, [32] x1 counter
    sv['x2_f32'] = sv['x2']
    if self.engine == ENG_LEGACY:
      sv['x2'] = a_fragments(sv['x2'])
    if self.engine == ENG_BF16:
      # hoisted action branch of dynhid0 (+ its bias): one (T*16)-row GEMM
      # (bf16 operands, fp32 accumulation and output -- what the in-scan mma did)
      sv['hid_pre'] = (torch.mm(sv['x2'].reshape(T * ROWS, H).to(torch.bfloat16), w['w_hid_x2'],
                                out_dtype=f32) + m('dyn/dynhid0/bias')).reshape(T, ROWS, D)
    sv['y0'][0] = rows16(y0.to(f32), B)
    sv['y1'][0] = rows16(y1.to(f32), B)
    vec = dict(
        b0=m('dyn/dynin0/bias'), b1=m('dyn/dynin1/bias'), b_hid=m('dyn/dynhid0/bias'),
        b_gru=m('dyn/dyngru/bias'), b_logit=m('dyn/obslogit/bias'),
        s0=m('dyn/dynin0norm/scale'), s1=m('dyn/dynin1norm/scale'),
        s_hid=m('dyn/dynhid0norm/scale'), s_obs=m('dyn/obs0norm/scale'))
    args = FwdArgs(B=B, T=T, D=D, H=H, S=S, C=C, G=G, engine=self.engine, ncta=self.ncta,
                   unimix=cfg.unimix, eps=1e-4)
    if self.timing:
      sv['timing'] = torch.zeros((2 * T * 16 + 4 * 160,), dtype=torch.int64, device=dev)
    for k, v in {**w, **vec, **sv}.items():
      if k in ('x2_f32', 'w_hid_x2'):
        continue
      assert v.is_contiguous(), k
      setattr(args, k, v.data_ptr())
    stream = torch.cuda.current_stream(dev).cuda_stream
    prof = None if torch.cuda.is_current_stream_capturing() else PROFILE
    if prof is not None:
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
    watch = _graph_timer('rssm_fwd')
    if watch:
      watch.start(stream)
    _lib.check(self.lib.emb_rssm_observe_fwd(ctypes.byref(args), stream))
    if watch:
      watch.stop(stream)
    if prof is not None:
      e1.record()
      prof.append(('rssm_fwd', e0, e1, T))
    self.last_args = args
    out = dict(
        deter=sv['deter'][:, :B].transpose(0, 1),
        logit=sv['logit'][:, :B].transpose(0, 1).reshape(B, T, S, C),
        index=sv['index'][:, :B].transpose(0, 1))
    return out, sv


# ---------------------------------------------------------------------- backward
class BwdArgs(ctypes.Structure):
  _fields_ = (
      [(n, _i32) for n in ('B', 'T', 'D', 'H', 'S', 'C', 'G', 'engine', 'ncta', 'hoist_x2')] +
      [('unimix', _fl), ('eps', _fl)] +
      [(n, _vp) for n in (
          'wt_in1', 'wt_logit', 'wt_ph1', 'wt_gru', 'wt_hid', 's0', 's1', 's_hid', 's_obs',
          'keep', 'deter0', 'deter', 'y0', 'y1', 'yobs', 'yhid', 'gates', 'sumsq', 'probs', 'rstd',
          'G_deter', 'G_logit', 'G_stoch',
          'g_xo', 'g_logit', 'g_gates', 'g_h', 'g_x0', 'g_x1', 'g_x2',
          'g_stoch', 'gd_carry', 'gd_tmp', 'dots', 'barrier', 'frag_scratch', 'gx_part', 'timing')])


@torch.no_grad()
def pack_bwd(store, cfg, engine, ncta):
  """Transposed copies for the backward scan, in the same packed layouts as the
  forward pass.  With the bf16 engines the action rows of dynhid0 are left out -- their input gradient is one
  (B*T)-row GEMM after the scan (hoisted like in the forward pass)."""
  D, G, H = cfg.deter, cfg.blocks, cfg.hidden
  Dg = D // G
  lay = engine       # fp32 [tile][K][8] | bf16 padded blocks (TMA ring) | bf16 plain blocks
  m = lambda n: store.view('master', n)
  hid = m('dyn/dynhid0/kernel')                                       # (G, Kh, Dg) -> (Dg, G*Kh)
  extra = {}
  if engine != ENG_F32:
    keep_rows = Dg + 2 * H
    extra['w_hid_x2'] = hid[:, keep_rows:].permute(1, 0, 2).reshape(-1, D).to(torch.bfloat16)  # (H, D)
    hid = hid[:, :keep_rows]
  if engine != ENG_F32 and FUSED_PACK:
    return dict(**extra, **_pack_bwd_fused(store, cfg, engine, ncta))
  # the torch formulation of the same layouts (fp32 parity engine, FUSED_PACK off)
  wobs = m('dyn/obs0/kernel')
  ph1 = torch.cat([wobs[:D], m('dyn/dynin0/kernel')], 1)              # (D, 2H)
  gru = m('dyn/dyngru/kernel')                                        # (G, Dg, 3Dg) -> (3Dg, G*Dg)
  gru_t = gru.permute(2, 0, 1).reshape(3 * Dg, D)
  hid_t = hid.permute(2, 0, 1).reshape(Dg, G * hid.shape[1])
  return dict(
      **extra,
      # whole latents per CTA (unit = lcm(C, 8) / 8 tiles) with the TMA engine: the softmax
      # jacobian then runs in the phase's epilogue
      wt_in1=pack_matrix(m('dyn/dynin1/kernel').t(), lay, ncta,
                         unit=(math.lcm(cfg.classes, 8) // 8 if engine == ENG_BF16 else 1)),
      wt_logit=pack_matrix(m('dyn/obslogit/kernel').t(), lay, ncta),
      wt_ph1=pack_matrix(ph1.t(), lay, ncta),
      wt_gru=pack_matrix(gru_t, lay, ncta, groups=G),
      wt_hid=pack_matrix(hid_t, lay, ncta, groups=G))


def _bind_bwd(lib):
  if getattr(lib, '_rssm_bwd_bound', False):
    return
  lib.emb_rssm_observe_bwd.argtypes = [ctypes.POINTER(BwdArgs), _vp]
  lib.emb_rssm_observe_bwd.restype = ctypes.c_int
  lib._rssm_bwd_bound = True


def _silu(x):
  return x * torch.sigmoid(x)


def _dsilu(n):
  sg = torch.sigmoid(n)
  return sg * (1 + n * (1 - sg))


def _norm_bwd(gx, y, s, eps=1e-4):
  """Through x = silu(rms(y) * s): returns (g_y, g_s, x)."""
  rstd = torch.rsqrt(y.square().mean(-1, keepdim=True) + eps)
  yhat = y * rstd
  n = yhat * s
  gn = gx * _dsilu(n)
  gs = (gn * yhat).reshape(-1, y.shape[-1]).sum(0)
  gy = rstd * s * gn - yhat * rstd * (gn * s * yhat).mean(-1, keepdim=True)
  return gy, gs, _silu(n)


@torch.no_grad()
def scan_backward(scan, sv, B, G_deter, G_logit, G_stoch, direct=None):
  """Runs emb_rssm_observe_bwd and forms every parameter / input gradient.
  G_*: batch-major upstream gradients (B, T, ..).  Returns (input grads, param grads).
  `direct(name)`: the fp32 view of `name` in the flat gradient buffer, for the weight gradients
  that are added there in place (their entry in the returned dict is None)."""
  cfg, dev, lib = scan.cfg, G_deter.device, scan.lib
  _bind_bwd(lib)
  T = sv['keep'].shape[0] - 1
  D, H, S, C, G = cfg.deter, cfg.hidden, cfg.stoch, cfg.classes, cfg.blocks
  Dg, SC, Kh = D // G, S * C, D // G + 3 * H
  if scan.packed_bwd is None or scan.packed_bwd_step != scan.store.version:
    scan.packed_bwd = pack_bwd(scan.store, cfg, scan.engine, scan.ncta)
    scan.packed_bwd_step = scan.store.version
  m = lambda n: scan.store.view('master', n)
  zeros = lambda *s: torch.zeros(s, dtype=f32, device=dev)
  empty = lambda *s: torch.empty(s, dtype=f32, device=dev)
  x2f = sv['x2_f32']
  buf = dict(
      G_deter=time_major(G_deter, B), G_logit=time_major(G_logit.reshape(B, T, SC), B),
      G_stoch=time_major(G_stoch.reshape(B, T, SC), B),
      g_xo=empty(T, ROWS, H), g_logit=zeros(T, ROWS, SC), g_gates=empty(T, ROWS, 3 * D),
      g_h=empty(T, ROWS, D), g_x0=zeros(T + 1, ROWS, H), g_x1=zeros(T + 1, ROWS, H),
      g_x2=zeros(T, ROWS, H), g_stoch=empty(ROWS, SC), gd_carry=zeros(ROWS, D),
      gd_tmp=empty(ROWS, D), dots=zeros(T + 1, 4, ROWS),
      barrier=torch.zeros(4, dtype=torch.int32, device=dev),
      frag_scratch=torch.zeros(ROWS * (5 * D + 4 * H), dtype=torch.bfloat16, device=dev),
      gx_part=torch.zeros((G, ROWS, 2 * H), dtype=f32, device=dev))
  vec = dict(s0=m('dyn/dynin0norm/scale'), s1=m('dyn/dynin1norm/scale'),
             s_hid=m('dyn/dynhid0norm/scale'), s_obs=m('dyn/obs0norm/scale'))
  saved = {k: sv[k] for k in ('keep', 'deter0', 'deter', 'y0', 'y1', 'yobs', 'yhid', 'gates',
                              'sumsq', 'probs', 'rstd')}
  hoist = scan.engine != ENG_F32
  args = BwdArgs(B=B, T=T, D=D, H=H, S=S, C=C, G=G, engine=scan.engine, ncta=scan.ncta,
                 hoist_x2=int(hoist), unimix=cfg.unimix, eps=1e-4)
  if scan.timing:
    buf['timing'] = torch.zeros((T, 16), dtype=torch.int64, device=dev)
  scan.last_bwd_buf = buf
  for k, v in {**scan.packed_bwd, **vec, **saved, **buf}.items():
    if k == 'w_hid_x2':
      continue
    assert v.is_contiguous(), k
    setattr(args, k, v.data_ptr())
  stream = torch.cuda.current_stream(dev).cuda_stream
  capturing = torch.cuda.is_current_stream_capturing()
  ev = None if capturing else getattr(scan, 'bwd_events', None)
  prof = None if capturing else PROFILE
  if ev is not None or prof is not None:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
  watch = _graph_timer('rssm_bwd')
  if watch:
    watch.start(stream)
  _lib.check(lib.emb_rssm_observe_bwd(ctypes.byref(args), stream))
  if watch:
    watch.stop(stream)
  if ev is not None or prof is not None:
    e1.record()
    if ev is not None:
      ev.append((e0, e1))
    if prof is not None:
      prof.append(('rssm_bwd', e0, e1, T))

  # ---- parameter gradients: (T*16)-row GEMMs over the per-step layer gradients
  cd = f32 if scan.engine == ENG_F32 else torch.bfloat16
  mm = lambda a, b: (a.to(cd).t() @ b.to(cd)).to(f32)
  keep = sv['keep'][:T]                                               # (T, 16)
  deter = sv['deter']
  dprev = torch.cat([sv['deter0'][None], deter[:-1]], 0) * keep[..., None]
  R = T * ROWS
  s0, s1, s_hid, s_obs = vec['s0'], vec['s1'], vec['s_hid'], vec['s_obs']
  g_y0, g_s0, x0 = _norm_bwd(buf['g_x0'][:T], sv['y0'][:T], s0)
  g_y1, g_s1, x1 = _norm_bwd(buf['g_x1'][:T], sv['y1'][:T], s1)
  g_yhid, g_shid, h = _norm_bwd(buf['g_h'], sv['yhid'], s_hid)
  g_yobs, g_sobs, xo = _norm_bwd(buf['g_xo'], sv['yobs'], s_obs)
  pg = {}
  # dynin0 / dynin1: steps >= 1 run inside the kernel (step 0 is an input)
  pg['dyn/dynin0/kernel'] = mm(dprev[1:].reshape(-1, D), g_y0[1:].reshape(-1, H))
  pg['dyn/dynin0/bias'] = g_y0[1:].reshape(-1, H).sum(0)
  pg['dyn/dynin0norm/scale'] = g_s0
  w1 = torch.zeros((SC, H), dtype=f32, device=dev)
  if T > 1:
    # x1_pre = sum over latents of dynin1[latent*C + class]: its weight gradient is
    # multihot^T @ g_y1 -- one small GEMM instead of a 32x expanded scatter-add
    # (row block t pairs stoch[t] with g_y1[t+1]; the last block is zero so that the
    # GEMM's inner dimension stays T*16 like every other weight-gradient GEMM here)
    idx = sv['index'][:T].long()                                      # (T, 16, S)
    rows = (idx + torch.arange(S, device=dev) * C).reshape(-1, S)     # row of dynin1 per latent
    hot = torch.zeros((R, SC), dtype=cd, device=dev).scatter_(1, rows, 1.0)
    valid = torch.zeros(ROWS, dtype=f32, device=dev)
    valid[:B] = 1.0
    contrib = torch.zeros((T, ROWS, H), dtype=f32, device=dev)
    contrib[:T - 1] = g_y1[1:] * (keep[1:] * valid[None, :])[:, :, None]
    w1 = mm(hot, contrib.reshape(R, H))
  pg['dyn/dynin1/kernel'] = w1
  pg['dyn/dynin1/bias'] = g_y1[1:].reshape(-1, H).sum(0)
  pg['dyn/dynin1norm/scale'] = g_s1
  x012 = torch.cat([x0, x1, x2f], -1)                                 # (T, 16, 3H)
  gyh = g_yhid.reshape(R, G, Dg)

  def block_wgrad(name, a, b):                     # (R, G, K), (R, G, N) -> (G, K, N)
    if direct is None or cd == f32:
      return torch.einsum('rgk,rgn->gkn', a.to(cd), b.to(cd)).to(f32)
    gv = direct(name)
    torch.baddbmm(gv, a.to(cd).permute(1, 2, 0), b.to(cd).permute(1, 0, 2), out_dtype=f32, out=gv)
    return None
  if direct is None or cd == f32:
    inp = torch.cat([dprev.reshape(T, ROWS, G, Dg),
                     x012[:, :, None, :].expand(-1, -1, G, -1)], -1)  # (T, 16, G, Kh)
    pg['dyn/dynhid0/kernel'] = block_wgrad('dyn/dynhid0/kernel', inp.reshape(R, G, Kh), gyh)
  else:
    # rows [deter_g | x0 x1 x2] of every group are contiguous 2-D blocks of the gradient view: the
    # x012 operand is shared by the groups, so it is never replicated G times (33 M elements)
    gv = direct('dyn/dynhid0/kernel')                                 # (G, Kh, Dg)
    dp = dprev.reshape(R, G, Dg).to(cd)
    xt = x012.reshape(R, 3 * H).to(cd).t()
    gy = gyh.to(cd)
    for g in range(G):
      torch.addmm(gv[g, :Dg], dp[:, g].t(), gy[:, g], out_dtype=f32, out=gv[g, :Dg])
      torch.addmm(gv[g, Dg:], xt, gy[:, g], out_dtype=f32, out=gv[g, Dg:])
    pg['dyn/dynhid0/kernel'] = None
  pg['dyn/dynhid0/bias'] = g_yhid.reshape(R, D).sum(0)
  pg['dyn/dynhid0norm/scale'] = g_shid
  gg = buf['g_gates'].reshape(R, G, 3 * Dg)
  pg['dyn/dyngru/kernel'] = block_wgrad('dyn/dyngru/kernel', h.reshape(R, G, Dg), gg)
  pg['dyn/dyngru/bias'] = gg.reshape(R, 3 * D).sum(0)
  if direct is None or cd == f32:
    wobs = torch.zeros_like(m('dyn/obs0/kernel'))
    wobs[:D] = mm(deter.reshape(R, D), g_yobs.reshape(R, H))
    pg['dyn/obs0/kernel'] = wobs
  else:                                            # rows [D:] (the token half) get theirs through autograd
    gv = direct('dyn/obs0/kernel')[:D]
    torch.addmm(gv, deter.reshape(R, D).to(cd).t(), g_yobs.reshape(R, H).to(cd), out_dtype=f32, out=gv)
    pg['dyn/obs0/kernel'] = None
  pg['dyn/obs0norm/scale'] = g_sobs
  pg['dyn/obslogit/kernel'] = mm(xo.reshape(R, H), buf['g_logit'].reshape(R, SC))
  pg['dyn/obslogit/bias'] = buf['g_logit'].reshape(R, SC).sum(0)
  bm = lambda x: x[:, :B].transpose(0, 1)
  g_x2 = buf['g_x2']
  if hoist:      # the action rows of dynhid0, hoisted: g_x2 = g_yhid @ W[g][Dg+2H:]^T summed over groups
    g_x2 = torch.mm(g_yhid.reshape(R, D).to(torch.bfloat16), scan.packed_bwd['w_hid_x2'].t(),
                    out_dtype=f32).reshape(T, ROWS, H)
  ig = dict(y0=g_y0[0, :B], y1=g_y1[0, :B], x2=bm(g_x2), pre_tok=bm(g_yobs))
  return ig, pg, buf


PARAMS = (
    'dyn/dynin0/kernel', 'dyn/dynin0/bias', 'dyn/dynin0norm/scale',
    'dyn/dynin1/kernel', 'dyn/dynin1/bias', 'dyn/dynin1norm/scale',
    'dyn/dynhid0/kernel', 'dyn/dynhid0/bias', 'dyn/dynhid0norm/scale',
    'dyn/dyngru/kernel', 'dyn/dyngru/bias',
    'dyn/obs0/kernel', 'dyn/obs0norm/scale',
    'dyn/obslogit/kernel', 'dyn/obslogit/bias')


class ObserveFn(torch.autograd.Function):
  """deter, logit, stoch = RSSM.observe over T steps, one kernel each way."""

  @staticmethod
  def forward(ctx, scan, deter0, y0, y1, x2, pre_tok, keep, gumbel, *weights):
    out, sv = scan.forward(deter0, y0, y1, x2, pre_tok, keep, gumbel)
    cfg = scan.cfg
    B = keep.shape[0]
    ctx.scan, ctx.sv, ctx.B = scan, sv, B
    stoch = torch.nn.functional.one_hot(out['index'].long(), cfg.classes).to(deter0.dtype)
    ctx.mark_non_differentiable(out['index'])
    return out['deter'].to(deter0.dtype), out['logit'], stoch, out['index']

  @staticmethod
  def backward(ctx, G_deter, G_logit, G_stoch, _):
    scan, sv, B = ctx.scan, ctx.sv, ctx.B
    T = sv['keep'].shape[0] - 1
    cfg = scan.cfg
    z = lambda g, *s: torch.zeros((B, T, *s), dtype=f32, device=sv['deter'].device) \
        if g is None else g.to(f32)
    store = scan.store
    direct = None
    if DIRECT_WGRAD and scan.engine != ENG_F32 and all(ctx.needs_input_grad[8:]):
      direct = lambda name: store.view('grad', name)
    ig, pg, _ = scan_backward(
        scan, sv, B, z(G_deter, cfg.deter), z(G_logit, cfg.stoch, cfg.classes),
        z(G_stoch, cfg.stoch, cfg.classes), direct)
    if store.on_grad is not None:                  # the bucketed exchange counts arrivals per tensor
      for n in PARAMS:
        if pg[n] is None:
          store.on_grad(n)
    return (None, None, ig['y0'], ig['y1'], ig['x2'], ig['pre_tok'], None, None,
            *[pg[n] for n in PARAMS])
