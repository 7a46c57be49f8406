"""Device-resident row storage of the replay buffer.

Layout in HBM (one table per key, struct-of-arrays): ``table[k]`` is
``uint8[nslabs * chunksize, row_bytes(k)]``.  A *slab* is ``chunksize``
consecutive rows and backs exactly one reference ``Chunk``
(embodied/core/chunk.py:13-23); global row id = slab * chunksize + index.
All data movement is the row engine of libembodied_b200.so
(include/embodied_b200.h); the host only produces int64 row-id lists.

Hot-path calls never touch the host allocator: pinned staging, row-id buffers
and (optionally) output batches are recycled from small rings guarded by CUDA
events.
"""
import numpy as np
import torch

from .. import _lib

STEPID_BYTES = 20

# bench.py sets this to a list to collect ('gather', start, end) CUDA event
# pairs around the launches (the live per-launch timing of the roofline line).
PROFILE = None


def _np_to_torch_dtype(dtype):
  return {
      np.dtype(bool): torch.bool, np.dtype(np.uint8): torch.uint8,
      np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16,
      np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
      np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
      np.dtype(np.float64): torch.float64, np.dtype(np.uint16): torch.uint16,
      np.dtype(np.uint32): torch.uint32, np.dtype(np.uint64): torch.uint64,
  }[np.dtype(dtype)]


class KeySpec:
  __slots__ = ('name', 'dtype', 'shape', 'row_bytes', 'tdtype')

  def __init__(self, name, dtype, shape):
    self.name = name
    self.dtype = np.dtype(dtype)
    self.shape = tuple(int(x) for x in shape)
    self.row_bytes = int(self.dtype.itemsize * int(np.prod(self.shape, dtype=np.int64)))
    self.tdtype = _np_to_torch_dtype(self.dtype)


class _Staging:
  """One pinned host block + its device mirror, for `rows` rows of every key."""

  def __init__(self, specs, rows, device):
    self.rows = rows
    offsets, total = {}, 0
    for spec in specs.values():
      total = (total + 255) // 256 * 256
      offsets[spec.name] = total
      total += rows * spec.row_bytes
    total = max((total + 255) // 256 * 256, 256)
    self.host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    self.dev = torch.empty(total, dtype=torch.uint8, device=device)
    hostnp = self.host.numpy()
    self.views, self.devptr, self.spans = {}, {}, {}
    for spec in specs.values():
      off = offsets[spec.name]
      nbytes = rows * spec.row_bytes
      self.views[spec.name] = hostnp[off: off + nbytes].view(spec.dtype).reshape(
          (rows, *spec.shape))
      self.devptr[spec.name] = self.dev.data_ptr() + off
      self.spans[spec.name] = (off, nbytes)
    self.rowids = torch.empty(rows, dtype=torch.int64, pin_memory=True)
    self.rowids_np = self.rowids.numpy()
    self.rowids_dev = torch.empty(rows, dtype=torch.int64, device=device)
    self.event = None
    self.acts_host = {}

  def wait(self):
    if self.event is not None:
      self.event.synchronize()
      self.event = None


class DeviceStore:
  """HBM row tables + the launches that fill and read them."""

  def __init__(self, chunksize, device=None, staging_rows=256, nslabs=0):
    if not torch.cuda.is_available():
      raise RuntimeError(
          'embodied_b200 replay storage lives in GPU memory; no CUDA device is '
          'visible and there is no CPU fallback.')
    self.lib = _lib.load()
    self.device = torch.device(device if device is not None else 'cuda')
    self.chunksize = int(chunksize)
    self.specs = None
    self.tables = {}
    self.nslabs = 0
    self._want_slabs = int(nslabs)
    self._staging_rows = int(staging_rows)
    self._stagings = []
    self._turn = 0
    self._idx_ring = []
    self._gather_plans = {}      # (keys, annotate, consec?, nslabs) -> (emb_key_t array, key order)

  # ------------------------------------------------------------ configuration
  @property
  def configured(self):
    return self.specs is not None

  def configure(self, specs):
    """specs: ordered dict name -> (dtype, shape); fixed by the first row
    (reference Chunk.append lazy allocation, embodied/core/chunk.py:43-47)."""
    self.specs = {k: KeySpec(k, *v) for k, v in specs.items()}
    if len(self.specs) + 1 > _lib.MAX_KEYS:
      raise ValueError(f'at most {_lib.MAX_KEYS - 1} keys per transition')
    self._stagings = [
        _Staging(self.specs, self._staging_rows, self.device) for _ in range(2)]
    self.reserve(max(self._want_slabs, 1))

  @property
  def bytes_per_row(self):
    return sum(s.row_bytes for s in self.specs.values())

  def reserve(self, nslabs):
    """Grow every table to hold at least `nslabs` slabs (realloc + D2D copy)."""
    nslabs = int(nslabs)
    if nslabs <= self.nslabs:
      return
    rows = nslabs * self.chunksize
    for spec in self.specs.values():
      new = torch.empty((rows, max(spec.row_bytes, 1)), dtype=torch.uint8,
                        device=self.device)
      old = self.tables.get(spec.name)
      if old is not None:
        new[: old.shape[0]].copy_(old)
      self.tables[spec.name] = new
    self.nslabs = nslabs

  # ------------------------------------------------------------------ staging
  def staging(self):
    """Pinned numpy views {key: (rows, *shape)} + rowid buffer to fill."""
    st = self._stagings[self._turn]
    st.wait()
    return st

  def commit_staging(self, n, device_values=None):
    """Rows [0, n) of the current staging block -> their table rows
    (st.rowids_np[:n]).  Keys in `device_values` come from device tensors
    (n, *shape) instead of the staging block.  One H2D + one launch."""
    st = self._stagings[self._turn]
    self._turn ^= 1
    if n == 0:
      return
    stream = torch.cuda.current_stream(self.device)
    device_values = device_values or {}
    with torch.cuda.stream(stream):
      lo = min(off for k, (off, _) in st.spans.items() if k not in device_values)
      hi = 0
      for k, (off, _) in st.spans.items():
        if k not in device_values:
          hi = max(hi, off + n * self.specs[k].row_bytes)
      if hi > lo:
        st.dev[lo:hi].copy_(st.host[lo:hi], non_blocking=True)
      st.rowids_dev[:n].copy_(st.rowids[:n], non_blocking=True)
    keys = []
    for spec in self.specs.values():
      if spec.row_bytes == 0:
        continue
      if spec.name in device_values:
        val = device_values[spec.name]
        src = self._as_rows(val, spec, n)
        srcptr = src.data_ptr()
      else:
        srcptr = st.devptr[spec.name]
      keys.append(_lib.Key(
          src=srcptr, dst=self.tables[spec.name].data_ptr(),
          src_stride=spec.row_bytes, dst_stride=spec.row_bytes,
          row_bytes=spec.row_bytes, op=_lib.OP_COPY))
    _lib.check(self.lib.emb_replay_append_rows(
        _lib.keys_array(keys), len(keys), st.rowids_dev.data_ptr(), n,
        stream.cuda_stream))
    st.event = torch.cuda.Event()
    st.event.record(stream)

  def _as_rows(self, val, spec, n):
    if not isinstance(val, torch.Tensor) or not val.is_cuda:
      raise TypeError(f"device value for '{spec.name}' must be a CUDA tensor")
    if val.dtype != spec.tdtype:
      raise TypeError(
          f"'{spec.name}': dtype {val.dtype} != stored {spec.tdtype}")
    if tuple(val.shape) != (n, *spec.shape):
      raise ValueError(
          f"'{spec.name}': shape {tuple(val.shape)} != {(n, *spec.shape)}")
    return val if val.is_contiguous() else val.contiguous()

  def _rowids_to_device(self, rows_np, stream=None):
    """int64 row ids -> device, through a small ring of pinned buffers."""
    n = len(rows_np)
    slot = None
    for cand in self._idx_ring:
      if cand[0].numel() >= n and (cand[2] is None or cand[2].query()):
        slot = cand
        break
    if slot is None:
      cap = max(4096, 1 << int(np.ceil(np.log2(max(n, 1)))))
      host = torch.empty(cap, dtype=torch.int64, pin_memory=True)
      slot = [host, torch.empty(cap, dtype=torch.int64, device=self.device), None, host.numpy()]
      self._idx_ring.append(slot)
      if len(self._idx_ring) > 8:
        self._idx_ring.pop(0)
    host, dev, _, host_np = slot
    host_np[:n] = rows_np
    if stream is None:
      stream = torch.cuda.current_stream(self.device)
    dev[:n].copy_(host[:n], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    slot[2] = ev
    return dev

  # ------------------------------------------------- fused driver launches
  def device_view(self, st, name, n):
    spec = self.specs[name]
    off, _ = st.spans[name]
    flat = st.dev[off: off + n * spec.row_bytes]
    return flat.view(spec.tdtype).reshape((n, *spec.shape))

  def stage_obs(self, st, n, obs_keys, norm_keys):
    """emb_driver_stage_obs: staged observation rows (+ stepid) -> their table
    rows, and float(u8)/255-0.5 of every key in `norm_keys` -> dense float32
    (n, *shape), in one launch after one H2D.  Returns ({key: device tensor},
    {key: normalised float32 tensor})."""
    stream = torch.cuda.current_stream(self.device)
    names = list(obs_keys) + ['stepid']
    lo = min(st.spans[k][0] for k in names)
    hi = max(st.spans[k][0] + n * self.specs[k].row_bytes for k in names)
    st.dev[lo:hi].copy_(st.host[lo:hi], non_blocking=True)
    st.rowids_dev[:n].copy_(st.rowids[:n], non_blocking=True)
    keys, obs, normed = [], {}, {}
    for name in names:
      spec = self.specs[name]
      if name != 'stepid':
        obs[name] = self.device_view(st, name, n)
      if spec.row_bytes == 0:
        continue
      key = _lib.Key(
          src=st.devptr[name], dst=self.tables[name].data_ptr(),
          src_stride=spec.row_bytes, dst_stride=spec.row_bytes,
          row_bytes=spec.row_bytes, op=_lib.OP_COPY)
      if name in norm_keys:
        assert spec.dtype == np.uint8, (name, spec.dtype)
        normed[name] = torch.empty(
            (n, *spec.shape), dtype=torch.float32, device=self.device)
        key.op = _lib.OP_NORM_U8_F32
        key.dst2 = normed[name].data_ptr()
        key.dst2_stride = spec.row_bytes * 4
      keys.append(key)
    _lib.check(self.lib.emb_driver_stage_obs(
        _lib.keys_array(keys), len(keys), st.rowids_dev.data_ptr(), n,
        stream.cuda_stream))
    return obs, normed

  def commit_acts(self, st, n, acts, outs, is_last):
    """emb_driver_scatter_mask_actions: acts * ~is_last -> table rows AND dense
    masked copies (returned on the host, the envs need them); policy outputs
    -> table rows.  `acts`/`outs`: {key: CUDA tensor (n, *shape)}."""
    stream = torch.cuda.current_stream(self.device)
    keys, keep, dense = [], [], {}
    for name, val in acts.items():
      spec = self.specs[name]
      src = self._as_rows(val, spec, n)
      keep.append(src)
      dense[name] = torch.empty_like(src)
      keys.append(_lib.Key(
          src=src.data_ptr(), dst=self.tables[name].data_ptr(),
          dst2=dense[name].data_ptr(), aux=is_last.data_ptr(), aux_stride=1,
          src_stride=spec.row_bytes, dst_stride=spec.row_bytes,
          dst2_stride=spec.row_bytes, row_bytes=spec.row_bytes,
          op=_lib.OP_MASK, dtype=_lib.DTYPES[spec.dtype]))
    for name, val in outs.items():
      spec = self.specs[name]
      src = self._as_rows(val, spec, n)
      keep.append(src)
      keys.append(_lib.Key(
          src=src.data_ptr(), dst=self.tables[name].data_ptr(),
          src_stride=spec.row_bytes, dst_stride=spec.row_bytes,
          row_bytes=spec.row_bytes, op=_lib.OP_COPY))
    _lib.check(self.lib.emb_driver_scatter_mask_actions(
        _lib.keys_array(keys), len(keys), st.rowids_dev.data_ptr(), n,
        stream.cuda_stream))
    host = {}
    for name, val in dense.items():
      buf = st.acts_host.get(name)
      if buf is None or buf.shape != val.shape:
        buf = st.acts_host[name] = torch.empty(
            val.shape, dtype=val.dtype, pin_memory=True)
      buf.copy_(val, non_blocking=True)
      host[name] = buf
    st.event = torch.cuda.Event()
    st.event.record(stream)
    st.event.synchronize()          # the envs need the actions now
    return {k: v.numpy().copy() for k, v in host.items()}

  # ------------------------------------------------------------------- gather
  def gather(self, src_rows, batch, window, consec=None, annotate=True,
             keys=None, out=None):
    """Replay._assemble_batch + _annotate_batch (+ the 'consec' key) in one
    launch.  `src_rows`: int64[batch*window] table rows; returns dense
    (batch, window, *shape) device tensors."""
    rows = np.asarray(src_rows, np.int64).reshape(-1)
    assert len(rows) == batch * window, (len(rows), batch, window)
    stream = torch.cuda.current_stream(self.device)
    rows_dev = self._rowids_to_device(rows, stream)
    names = list(self.specs) if keys is None else list(keys)
    out = {} if out is None else out
    # The key table of a (key set, annotate, consec) signature is built once; per call only the
    # destination (and, after a table re-allocation, source) pointers are patched in place.
    sig = (tuple(names), bool(annotate), consec is not None, self.nslabs)
    plan = self._gather_plans.get(sig)
    if plan is None:
      klist, slots = [], []
      first = self.specs.get('is_first')
      for name in names:
        spec = self.specs[name]
        if spec.row_bytes == 0:
          continue
        op, aux, aux_stride = _lib.OP_COPY, None, 0
        if annotate and name == 'is_first':
          op = _lib.OP_FIRST
        if annotate and name == 'is_last' and first is not None:
          op, aux = _lib.OP_LAST, self.tables['is_first'].data_ptr()
          aux_stride = first.row_bytes
        slots.append(name)
        klist.append(_lib.Key(
            src=self.tables[name].data_ptr(), aux=aux, aux_stride=aux_stride,
            src_stride=spec.row_bytes, dst_stride=spec.row_bytes,
            row_bytes=spec.row_bytes, op=op))
      if consec is not None:
        slots.append('consec')
        klist.append(_lib.Key(dst_stride=4, row_bytes=4, op=_lib.OP_FILL32))
      if len(self._gather_plans) > 16:
        self._gather_plans.clear()
      plan = self._gather_plans[sig] = (_lib.keys_array(klist), slots)
    karr, slots = plan
    for name in names:
      if name not in out:
        spec = self.specs[name]
        out[name] = torch.empty(
            (batch, window, *spec.shape), dtype=spec.tdtype, device=self.device)
    if consec is not None:
      out['consec'] = torch.empty(
          (batch, window), dtype=torch.int32, device=self.device)
      karr[len(slots) - 1].fill = int(consec)
    for i, name in enumerate(slots):
      karr[i].dst = out[name].data_ptr()
    prof = PROFILE
    if prof is not None:
      t0 = torch.cuda.Event(enable_timing=True)
      t1 = torch.cuda.Event(enable_timing=True)
      t0.record(stream)
    _lib.check(self.lib.emb_replay_gather(
        karr, len(slots), rows_dev.data_ptr(),
        batch * window, window, stream.cuda_stream))
    if prof is not None:
      t1.record(stream)
      prof.append(('gather', t0, t1))
    return out

  # ------------------------------------------------------------ write (update)
  def scatter(self, dst_rows, values):
    """Replay.update: rows of `values` (device (n, *shape) tensors) -> table
    rows `dst_rows` (int64[n], -1 = skip)."""
    n = len(dst_rows)
    if n == 0 or not values:
      return
    stream = torch.cuda.current_stream(self.device)
    rows_dev = self._rowids_to_device(np.asarray(dst_rows, np.int64))
    keep, klist = [], []
    for name, val in values.items():
      spec = self.specs[name]
      if not isinstance(val, torch.Tensor):
        val = torch.as_tensor(np.ascontiguousarray(val)).to(
            self.device, non_blocking=False)
      val = val.reshape(n, *spec.shape)
      src = self._as_rows(val, spec, n)
      keep.append(src)
      klist.append(_lib.Key(
          src=src.data_ptr(), dst=self.tables[name].data_ptr(),
          src_stride=spec.row_bytes, dst_stride=spec.row_bytes,
          row_bytes=spec.row_bytes, op=_lib.OP_COPY))
    _lib.check(self.lib.emb_replay_scatter_update(
        _lib.keys_array(klist), len(klist), rows_dev.data_ptr(), n,
        stream.cuda_stream))
    del keep   # same-stream temporaries: the caching allocator orders reuse

  # --------------------------------------------------------------- chunk I/O
  def export_slab(self, slab, length):
    """Rows [0, length) of a slab as host numpy arrays (chunk save): emb_replay_export_chunk,
    one strided device-to-pinned-host copy per key, one stream synchronise."""
    lo = slab * self.chunksize
    stream = torch.cuda.current_stream(self.device)
    hosts, keys = {}, []
    for spec in self.specs.values():
      if spec.row_bytes == 0 or length == 0:
        continue
      hosts[spec.name] = torch.empty((length, spec.row_bytes), dtype=torch.uint8, pin_memory=True)
      keys.append(_lib.Key(src=self.tables[spec.name].data_ptr(), dst=hosts[spec.name].data_ptr(),
                           src_stride=self.tables[spec.name].stride(0), dst_stride=spec.row_bytes,
                           row_bytes=spec.row_bytes))
    if keys:
      _lib.check(self.lib.emb_replay_export_chunk(
          _lib.keys_array(keys), len(keys), lo, length, stream.cuda_stream))
      stream.synchronize()
    out = {}
    for spec in self.specs.values():
      if spec.name in hosts:
        out[spec.name] = hosts[spec.name].numpy().reshape(-1).view(spec.dtype).reshape(
            (length, *spec.shape)).copy()
      else:
        out[spec.name] = np.empty((length, *spec.shape), spec.dtype)
    return out

  def import_slab(self, slab, data):
    """Host arrays {key: (length, *shape)} -> rows [0, length) of a slab (chunk load):
    emb_replay_import_chunk, one strided copy per key."""
    lo = slab * self.chunksize
    stream = torch.cuda.current_stream(self.device)
    keys, keep, n = [], [], 0
    for spec in self.specs.values():
      arr = np.ascontiguousarray(data[spec.name], spec.dtype)
      n = len(arr)
      if n > self.chunksize:
        raise ValueError(f'{n} rows do not fit a slab of {self.chunksize}')
      if spec.row_bytes == 0 or n == 0:
        continue
      raw = torch.from_numpy(arr.reshape(n, -1).view(np.uint8))
      keep.append(raw)
      keys.append(_lib.Key(src=raw.data_ptr(), dst=self.tables[spec.name].data_ptr(),
                           src_stride=spec.row_bytes, dst_stride=self.tables[spec.name].stride(0),
                           row_bytes=spec.row_bytes))
    if keys:
      _lib.check(self.lib.emb_replay_import_chunk(
          _lib.keys_array(keys), len(keys), lo, n, stream.cuda_stream))
      stream.synchronize()          # pageable sources: the arrays may go away after this call
