"""Device side of Driver._mask for host-resident actions
(embodied/core/driver.py:72-74,84-87): actions of a host agent are masked by
the same ``emb_driver_scatter_mask_actions`` launch the fused step uses."""
import numpy as np
import torch

from .. import _lib
from . import store as storelib


class DeviceOps:

  def __init__(self, device=None):
    if not torch.cuda.is_available():
      raise RuntimeError(
          'embodied_b200.Driver masks actions on the GPU; no CUDA device is '
          'visible and there is no CPU fallback.')
    self.lib = _lib.load()
    self.device = torch.device(device if device is not None else 'cuda')

  def mask_actions(self, acts, is_last):
    stream = torch.cuda.current_stream(self.device)
    n = len(is_last)
    flags = torch.from_numpy(np.ascontiguousarray(is_last)).to(self.device)
    keys, outs, keep = [], {}, []
    for name, val in acts.items():
      val = np.ascontiguousarray(val)
      row_bytes = val.dtype.itemsize * int(np.prod(val.shape[1:], dtype=np.int64))
      if row_bytes == 0:
        outs[name] = (None, val)
        continue
      src = torch.from_numpy(val.reshape(n, -1).view(np.uint8)).to(self.device)
      dst = torch.empty_like(src)
      keep += [src, dst]
      outs[name] = (dst, val)
      keys.append(_lib.Key(
          src=src.data_ptr(), dst2=dst.data_ptr(), aux=flags.data_ptr(),
          aux_stride=1, src_stride=row_bytes, dst2_stride=row_bytes,
          row_bytes=row_bytes, op=_lib.OP_MASK,
          dtype=_lib.DTYPES[val.dtype]))
    if keys:
      _lib.check(self.lib.emb_driver_scatter_mask_actions(
          _lib.keys_array(keys), len(keys), None, n, stream.cuda_stream))
    result = {}
    for name, (dst, val) in outs.items():
      if dst is None:
        result[name] = val
      else:
        result[name] = dst.cpu().numpy().reshape(-1).view(val.dtype).reshape(
            val.shape)
    return result
