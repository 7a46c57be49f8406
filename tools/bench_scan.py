"""Times the fused RSSM scan kernels at BASELINE config-2 size (size200m, B=16,
T=64) with CUDA events; prints achieved weight-streaming GB/s vs the measured
HBM peak.  Algorithmic bytes per launch = T x (in-scan weight bytes)."""
import json
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200.dreamerv3 import config as C, params as P, scan as S  # noqa: E402

ROOT = pathlib.Path(__file__).resolve().parent.parent
try:
  PEAK = json.load(open(ROOT / 'MEASURED_PEAKS.json'))['hbm_gbs']
except Exception:
  PEAK = 6650.0


def main():
  size = sys.argv[1] if len(sys.argv) > 1 else 'size200m'
  ename = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
  engine = {'bf16': S.ENG_BF16, 'legacy': S.ENG_LEGACY, 'f32': S.ENG_F32}[ename]
  B, T = 16, 64
  cfg = C.make(size)
  store = P.ParamStore(cfg, 'cuda', torch.float32, 0)
  sc = S.Scan(cfg, store, engine)
  D, H, Sx, Cx, G = cfg.deter, cfg.hidden, cfg.stoch, cfg.classes, cfg.blocks
  g = torch.Generator(device='cuda').manual_seed(0)
  r = lambda *s: torch.randn(s, generator=g, device='cuda')
  args = (r(B, D) * 0.3, r(B, H), r(B, H), r(B, T, H), r(B, T, H),
          torch.ones(B, T, device='cuda'), r(B, T, Sx, Cx))
  # in-scan weights streamed per step (the TMA engine hoists the action rows of dynhid0)
  nw = D * 2 * H + H * Sx * Cx + D * (D // G + (2 if engine == S.ENG_BF16 else 3) * H) + D * 3 * (D // G)
  wbytes = nw * (4 if engine == S.ENG_F32 else 2)
  sc.timing = True
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  for _ in range(3):
    sc.forward(*args)
  torch.cuda.synchronize()
  times = []
  for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # time the kernel launch alone: prepare buffers outside the events
    out, sv = sc.forward(*args)
    torch.cuda.synchronize()
    sv['sumsq'].zero_(); sv['barrier'].zero_()
    fa = sc.last_args
    a.record()
    sc.relaunch(fa)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b) * 1e-3)
  t = float(np.median(times))
  raw = sv['timing'].cpu().numpy().astype(np.int64)
  tm = raw[:2 * T * 16].reshape(-1, T, 16)
  per = raw[2 * T * 16:].reshape(-1, 4)[:148]
  hid = [c for c in range(len(per)) if per[c, 2] and per[c, 3] > per[c, 2] > per[c, 1] > per[c, 0]]
  if hid and T > 20:
    # per-CTA marks of dynhid0 at t = 20 (CTAs with block-diagonal tiles): k ranges 0/1 done, x1 landed,
    # k range 2 done, in us after the CTA's own start of the phase
    rel = (per[hid] - per[hid][:, :1]) / 1e3
    for name, col in zip(('ranges 0/1 done', 'x1 landed', 'range 2 done'), rel.T[[1, 2, 3]]):
      print(f'  dynhid0 per CTA, {name}: min {col.min():.2f} median {np.median(col):.2f} max {col.max():.2f} us')
  names = ['P4 prologue', 'P4 gemm', 'bar', 'P5', 'bar', 'P1', 'bar', 'P2', 'bar', 'P3', 'bar']
  for which, tmx in enumerate(tm):
    d = np.diff(tmx[:, :12], axis=1)[8:].mean(0) / 1e3
    print(f'phase us (CTA set {which}, mean over steps 8..):',
          {n + str(i): round(float(x), 2) for i, (n, x) in enumerate(zip(names, d))})
    if tmx[:, 12:].any():
      x = tmx[8:]
      sub = {'P5 A-build': x[:, 12] - x[:, 3], 'P5 gemm': x[:, 13] - x[:, 12], 'P5 epilogue': x[:, 4] - x[:, 13],
             'P2 A-build': x[:, 14] - x[:, 7], 'P2 gemm': x[:, 15] - x[:, 14], 'P2 rest': x[:, 8] - x[:, 15]}
      print('   ', {k: round(float(v.mean()) / 1e3, 2) for k, v in sub.items()})
  print(json.dumps({
      'kernel': 'rssm_fwd_kernel', 'size': size, 'engine': ename,
      'B': B, 'T': T, 'ms': t * 1e3, 'us_per_step': t / T * 1e6,
      'weights_per_step_MB': wbytes / 1e6, 'algorithmic_bytes': wbytes * T,
      'GBs': wbytes * T / t / 1e9, 'frac_of_measured_peak': wbytes * T / t / 1e9 / PEAK}))
  bench_bwd(sc, args, B, T, cfg, wbytes, size, ename)


def bench_bwd(sc, args, B, T, cfg, wbytes, size, engine):
  sc.timing = engine == 'bf16'
  sc.bwd_events = []
  g = torch.Generator(device='cuda').manual_seed(1)
  r = lambda *s: torch.randn(s, generator=g, device='cuda') * 0.01
  Gd, Gl, Gs = r(B, T, cfg.deter), r(B, T, cfg.stoch, cfg.classes), r(B, T, cfg.stoch, cfg.classes)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  for i in range(8):
    out, sv = sc.forward(*args)
    flush.zero_()
    S.scan_backward(sc, sv, B, Gd, Gl, Gs)
  torch.cuda.synchronize()
  times = [a.elapsed_time(b) * 1e-3 for a, b in sc.bwd_events[3:]]
  t = float(np.median(times))
  if 'timing' in getattr(sc, 'last_bwd_buf', {}):
    tm = sc.last_bwd_buf['timing'].cpu().numpy().astype(np.int64)
    d = np.diff(tm[:, :11], axis=1)[4:-4].mean(0) / 1e3
    names = ['B1', 'bar', 'B2', 'bar', 'B3', 'bar', 'B4', 'bar', 'B5', 'bar']
    print('bwd phase us (CTA 0):', {n + str(i): round(float(x), 2) for i, (n, x) in enumerate(zip(names, d))})
  print(json.dumps({
      'kernel': 'rssm_bwd_kernel', 'size': size, 'engine': engine,
      'B': B, 'T': T, 'ms': t * 1e3, 'us_per_step': t / T * 1e6,
      'algorithmic_bytes': wbytes * T, 'GBs': wbytes * T / t / 1e9,
      'frac_of_measured_peak': wbytes * T / t / 1e9 / PEAK}))


if __name__ == '__main__':
  main()
