"""ctypes binding of include/embodied_b200.h.

There is no fallback: if libembodied_b200.so is missing or CUDA is unavailable
every product entry point raises.
"""
import ctypes
import threading

import numpy as np

from . import build as _build

MAX_KEYS = 32

OP_COPY, OP_FIRST, OP_LAST, OP_FILL32, OP_MASK, OP_NORM_U8_F32, OP_NOT = range(7)

DTYPES = {
    np.dtype(np.uint8): 0, np.dtype(bool): 1, np.dtype(np.int32): 2,
    np.dtype(np.int64): 3, np.dtype(np.float32): 4, np.dtype(np.float64): 5,
    np.dtype(np.float16): 6, np.dtype(np.int16): 8, np.dtype(np.int8): 9,
    np.dtype(np.uint16): 10, np.dtype(np.uint32): 11, np.dtype(np.uint64): 12,
}
DTYPE_BF16 = 7


class Key(ctypes.Structure):
  _fields_ = [
      ('src', ctypes.c_void_p), ('dst', ctypes.c_void_p),
      ('dst2', ctypes.c_void_p), ('aux', ctypes.c_void_p),
      ('src_stride', ctypes.c_uint64), ('dst_stride', ctypes.c_uint64),
      ('dst2_stride', ctypes.c_uint64), ('aux_stride', ctypes.c_uint64),
      ('row_bytes', ctypes.c_uint32), ('op', ctypes.c_uint32),
      ('dtype', ctypes.c_uint32), ('fill', ctypes.c_int32),
  ]


_LIB = None
_LOCK = threading.Lock()

EXPORTS = (
    'emb_last_error', 'emb_abi_version', 'emb_launch_count', 'emb_launch_count_add',
    'emb_device_sm_count', 'emb_rows_copy', 'emb_replay_gather',
    'emb_replay_append_rows', 'emb_replay_scatter_update',
    'emb_driver_stage_obs', 'emb_driver_scatter_mask_actions',
    'emb_replay_export_chunk', 'emb_replay_import_chunk',
    # learner (bound where they are used: dreamerv3/scan.py, ops.py, optim.py)
    'emb_rssm_observe_fwd', 'emb_rssm_observe_bwd', 'emb_rssm_tma_fits', 'emb_rssm_legacy_fits', 'emb_rmsnorm_act_fwd',
    'emb_rmsnorm_act_bwd', 'emb_rssm_kl_fwd', 'emb_rssm_kl_bwd', 'emb_lambda_return', 'emb_onehot_sample', 'emb_opt_agc_rms_momentum', 'emb_opt_agc_rms_momentum_cast', 'emb_allreduce_bucket_update', 'emb_maxpool2_nhwc_fwd',
    'emb_maxpool2_nhwc_bwd', 'emb_upsample2_nhwc_fwd', 'emb_upsample2_nhwc_bwd',
    'emb_rmsnorm_grouped_fwd', 'emb_gru_gates_fwd', 'emb_pack_tiles', 'emb_conv_patches_nhwc', 'emb_conv_tapsum_nhwc', 'emb_probe_read', 'emb_twohot_loss_fwd', 'emb_twohot_loss_bwd', 'emb_twohot_pred', 'emb_loss_reduce', 'emb_conv5x5_nhwc_tc', 'emb_conv5x5_wgrad_tc', 'emb_conv_nhwc_tc', 'emb_conv_wgrad_tc', 'emb_event_create', 'emb_event_record', 'emb_event_elapsed_ms', 'emb_event_destroy',
    'emb_gae_advantage', 'emb_opt_clip_adam', 'emb_gumbel_fill',
)


def load():
  """dlopen the library (building is __graft_entry__.build()'s job)."""
  global _LIB
  with _LOCK:
    if _LIB is not None:
      return _LIB
    if not _build.LIB.exists():
      raise RuntimeError(
          f'{_build.LIB} is missing: the CUDA extension was not built. Run '
          '`python -c "import __graft_entry__ as g; g.build()"`. '
          'embodied_b200 has no CPU fallback.')
    lib = ctypes.CDLL(str(_build.LIB))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    kp = ctypes.POINTER(Key)
    lib.emb_last_error.restype = ctypes.c_char_p
    lib.emb_abi_version.restype = ctypes.c_int
    lib.emb_launch_count.restype = ctypes.c_uint64
    lib.emb_launch_count_add.argtypes = [ctypes.c_uint64]
    lib.emb_launch_count_add.restype = None
    lib.emb_device_sm_count.restype = ctypes.c_int
    lib.emb_rows_copy.argtypes = [kp, ctypes.c_int, vp, vp, i64, i32, vp]
    lib.emb_replay_gather.argtypes = [kp, ctypes.c_int, vp, i64, i32, vp]
    for name in ('emb_replay_append_rows', 'emb_replay_scatter_update',
                 'emb_driver_stage_obs', 'emb_driver_scatter_mask_actions'):
      getattr(lib, name).argtypes = [kp, ctypes.c_int, vp, i64, vp]
    for name in ('emb_replay_export_chunk', 'emb_replay_import_chunk'):
      getattr(lib, name).argtypes = [kp, ctypes.c_int, i64, i64, vp]
    for name in EXPORTS[5:13]:
      getattr(lib, name).restype = ctypes.c_int
    lib.emb_event_create.argtypes = [ctypes.POINTER(vp)]
    lib.emb_event_record.argtypes = [vp, vp]
    lib.emb_event_elapsed_ms.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_float)]
    lib.emb_event_destroy.argtypes = [vp]
    for name in ('emb_event_create', 'emb_event_record', 'emb_event_elapsed_ms', 'emb_event_destroy'):
      getattr(lib, name).restype = ctypes.c_int
    fl, u8p = ctypes.c_float, vp
    lib.emb_gae_advantage.argtypes = [vp, vp, u8p, u8p, vp, vp, i64, i32, fl, fl, vp]
    lib.emb_gae_advantage.restype = ctypes.c_int
    lib.emb_opt_clip_adam.argtypes = [vp, vp, vp, vp, vp, i64, vp, i32, vp, fl, i32, fl, fl, fl, fl, fl, vp]
    lib.emb_opt_clip_adam.restype = ctypes.c_int
    lib.emb_gumbel_fill.argtypes = [vp, i64, ctypes.c_uint64, vp]
    lib.emb_gumbel_fill.restype = ctypes.c_int
    _LIB = lib
    return lib


def check(code):
  if code != 0:
    raise RuntimeError(
        f'libembodied_b200: {load().emb_last_error().decode()} (code {code})')


def launch_count():
  return int(load().emb_launch_count())


def launch_count_add(n):
  load().emb_launch_count_add(int(n))


def keys_array(keys):
  if len(keys) > MAX_KEYS:
    raise ValueError(f'{len(keys)} keys > EMB_MAX_KEYS={MAX_KEYS}')
  return (Key * len(keys))(*keys)


class Stopwatch:
  """A pair of CUDA events around a launch, usable inside a stream capture
  (include/embodied_b200.h emb_event_*): `start(stream)`, `stop(stream)`, and
  after a synchronise `ms()`.  Inside a CUDA graph every replay re-records."""

  def __init__(self):
    lib = load()
    self.a, self.b = ctypes.c_void_p(), ctypes.c_void_p()
    check(lib.emb_event_create(ctypes.byref(self.a)))
    check(lib.emb_event_create(ctypes.byref(self.b)))

  def start(self, stream):
    check(load().emb_event_record(self.a, stream))

  def stop(self, stream):
    check(load().emb_event_record(self.b, stream))

  def ms(self):
    out = ctypes.c_float()
    check(load().emb_event_elapsed_ms(self.a, self.b, ctypes.byref(out)))
    return out.value
