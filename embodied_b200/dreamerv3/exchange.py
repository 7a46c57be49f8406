"""Bucketed gradient exchange + optimiser, overlapped with the backward pass.

Reference: embodied/jax/opt.py:52-54 (`pmean` of the gradients over the data-parallel axes,
which XLA pipelines with the backward pass: embodied/jax/internal.py:63-79) followed by the
optimiser chain (opt.py:109-164).  Here the flat gradient buffer (params.py) is cut into
contiguous BUCKETS of whole tensors -- heads | decoder | encoder | dynamics, i.e. roughly the
order in which the backward pass completes them -- and the moment the last gradient of a
bucket has been accumulated, `emb_allreduce_bucket_update` is launched for it on a side
stream: ncclAllReduce(avg) of the bucket in place, then AGC + RMS + momentum on exactly those
tensors (also refreshing their bf16 copies).  The exchange and the update of the early buckets
run underneath the rest of the backward pass; the main stream joins the side stream once, at
the end.  With one process there is no NCCL call and only the optimiser overlap remains.

Which gradient arrives when is learnt, not assumed: during the first (eager) update every
arrival is counted per bucket and everything is flushed at the end; from then on a bucket is
flushed when its count is reached.  The pattern is static (same autograd graph every step), and
the CUDA-graph capture records the same sequence.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .optim import CHUNK

_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
GROUPS = ('dyn', 'enc', 'dec', 'heads')


class _UniqueId(ctypes.Structure):
  _fields_ = [('internal', ctypes.c_byte * 128)]


class NcclComm:
  """An ncclComm_t of this process' own, over the ranks of torch.distributed's default group
  (the unique id travels through a torch.distributed broadcast).  The handle is what the C ABI
  takes (`include/embodied_b200.h emb_allreduce_bucket_update`)."""

  def __init__(self, device):
    import torch.distributed as dist
    self.lib = ctypes.CDLL('libnccl.so.2')              # the copy torch loaded
    self.lib.ncclGetErrorString.restype = ctypes.c_char_p
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = _UniqueId()
    if rank == 0:
      self._check(self.lib.ncclGetUniqueId(ctypes.byref(uid)))
    wire = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(device)
    dist.broadcast(wire, 0)
    ctypes.memmove(ctypes.byref(uid), bytes(wire.cpu().numpy().tobytes()), 128)
    self.handle = _vp()
    self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(_vp), ctypes.c_int, _UniqueId, ctypes.c_int]
    with torch.cuda.device(device):
      self._check(self.lib.ncclCommInitRank(ctypes.byref(self.handle), world, uid, rank))
    self.world = world

  def _check(self, code):
    if code != 0:
      raise RuntimeError(f'NCCL: {self.lib.ncclGetErrorString(code).decode()}')


def group_of(name):
  top = name.split('/', 1)[0]
  return top if top in ('dyn', 'enc', 'dec') else 'heads'


class GradExchange:

  def __init__(self, store, opt, comm=None):
    assert opt.fused, 'the bucketed exchange drives the fused optimiser kernels'
    self.store, self.opt, self.comm = store, opt, comm
    self.lib = _lib.load()
    self.lib.emb_allreduce_bucket_update.argtypes = [
        _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp]
    self.lib.emb_allreduce_bucket_update.restype = ctypes.c_int
    names = list(store.specs)
    # chunk table rows of tensor i (optim.Optimizer builds them in storage order)
    per = [-(-int(np.prod(store.specs[n][0])) // CHUNK) for n in names]
    first = np.concatenate([[0], np.cumsum(per)])
    self.buckets, self.bucket_of = [], {}
    i = 0
    while i < len(names):                      # maximal runs of one group, in storage order
      j = i
      while j < len(names) and group_of(names[j]) == group_of(names[i]):
        j += 1
      begin = store.offsets[names[i]]
      end = store.offsets[names[j]] if j < len(names) else store.total
      b = dict(group=group_of(names[i]), elem_begin=begin, elem_count=end - begin,
               chunk_begin=int(first[i]), nchunks=int(first[j] - first[i]),
               tensor_begin=i, tensor_count=j - i)
      for n in names[i:j]:
        self.bucket_of[n] = len(self.buckets)
      self.buckets.append(b)
      i = j
    self.side = torch.cuda.Stream(device=store.device)
    self.expected = None                       # arrivals per bucket in one backward pass
    self.seen = [0] * len(self.buckets)
    self.flushed = [True] * len(self.buckets)  # nothing pending outside begin() .. finish()
    self.active = False
    store.on_grad = self._arrived

  # ------------------------------------------------------------------ one update
  def begin(self):
    """Before the backward pass: step scalars on the device, counters reset."""
    self.opt.device_hyper()
    self.main = torch.cuda.current_stream(self.store.device)
    self.seen = [0] * len(self.buckets)
    self.flushed = [False] * len(self.buckets)
    self.active = True

  def _arrived(self, name):
    if not self.active:
      return
    b = self.bucket_of[name]
    if self.flushed[b]:
      raise RuntimeError(
          f'gradient of {name} arrived after its bucket was exchanged: the backward pass is not '
          'the one the arrival counts were learnt from (GradExchange assumes a static graph)')
    self.seen[b] += 1
    if self.expected is not None and self.seen[b] == self.expected[b] and not self.flushed[b]:
      self._flush(b)

  def _flush(self, b):
    st, bk = self.store, self.buckets[b]
    # the bucket's gradients are final on the stream the backward pass runs on (and, for a
    # leaf whose accumulation autograd placed on another stream of the same capture state, there)
    self.side.wait_stream(self.main)
    here = torch.cuda.current_stream(st.device)
    if here != self.main and here != self.side:
      capturing = torch.cuda.is_current_stream_capturing()
      with torch.cuda.stream(self.main):
        main_capturing = torch.cuda.is_current_stream_capturing()
      if capturing == main_capturing:
        self.side.wait_stream(here)
    low = st.low_buffer() if st.compute_dtype == torch.bfloat16 else None
    _lib.check(self.lib.emb_allreduce_bucket_update(
        None if self.comm is None else self.comm.handle,
        st.grad.data_ptr(), st.master.data_ptr(), st.nu.data_ptr(), st.mu.data_ptr(),
        None if low is None else low.data_ptr(), bk['elem_begin'], bk['elem_count'],
        self.opt.chunks.data_ptr() + 16 * bk['chunk_begin'], bk['nchunks'],
        self.opt.norms.data_ptr(), bk['tensor_begin'], bk['tensor_count'],
        self.opt.hyper.data_ptr(), self.opt.partials.data_ptr(), self.opt.tensor_first.data_ptr(),
        bk['chunk_begin'], self.side.cuda_stream))
    self.flushed[b] = True

  def finish(self):
    """After the backward pass: flush what is left, join the side stream; returns the
    gradient norm (of the averaged gradient) like Optimizer.launch."""
    for b in range(len(self.buckets)):
      if not self.flushed[b]:
        self._flush(b)
    if self.expected is None:
      self.expected = list(self.seen)
    elif self.expected != self.seen:
      # the autograd graph changed (e.g. another batch signature): learn again next step
      self.expected = None
    self.active = False
    main = torch.cuda.current_stream(self.store.device)
    # bench.py: how long the main stream will sit in the join below = the part of the exchange +
    # update that the backward pass did NOT hide (side-stream end minus backward end)
    from . import scan as scanlib
    watch = scanlib._graph_timer('exchange_tail')
    if watch:
      watch.start(main.cuda_stream)
      watch.stop(self.side.cuda_stream)
    main.wait_stream(self.side)
    if self.store.compute_dtype == torch.bfloat16:
      self.store.low_is_fresh()
    return self.opt.norms[0::2].sum().sqrt()
