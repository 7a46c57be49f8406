"""Read-only streaming bandwidth of the B200 with the scan kernels' weight path
(TMA bulk copies into a shared-memory ring) and with plain loads, next to the
measured copy peak.  Usage: python tools/probe_read.py"""
import ctypes
import json
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from embodied_b200 import _lib  # noqa: E402

lib = _lib.load()
vp, i32 = ctypes.c_void_p, ctypes.c_int32
lib.emb_probe_read.argtypes = [vp, ctypes.c_uint64, i32, i32, i32, i32, vp, vp]
lib.emb_probe_read.restype = ctypes.c_int
nbytes = 2 << 30
buf = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
buf.random_(0, 255)
sink = torch.zeros(4, dtype=torch.int32, device='cuda')
sms = lib.emb_device_sm_count()


def run(ncta, mode, nstages=7, stage=16384):
  stream = torch.cuda.current_stream().cuda_stream
  best = 1e9
  for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.emb_probe_read(buf.data_ptr(), nbytes, ncta, mode, nstages, stage, sink.data_ptr(), stream))
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) * 1e-3)
  moved = nbytes // ncta // stage * stage * ncta if mode == 0 else nbytes // 16 // ncta * 16 * ncta
  return moved / best / 1e9


out = {}
for ncta in (sms, 128, 2 * sms, 4 * sms):
  out[f'ldg_{ncta}cta'] = round(run(ncta, 1), 1)
for ncta in (sms, 128):
  for nstages, stage in ((4, 16384), (7, 16384), (12, 16384), (6, 32768), (3, 65536), (13, 16384), (14, 8192)):
    out[f'tma_{ncta}cta_{nstages}x{stage // 1024}K'] = round(run(ncta, 0, nstages, stage), 1)
print(json.dumps(out, indent=1))
