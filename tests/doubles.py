"""Test doubles (NOT product code): numpy stand-ins for the two device-backed
collaborators so that the host-side index logic of Replay / Driver / run.train
can be exercised on a machine without a GPU.  The product never imports this
file; without CUDA, `Replay()` / `Driver` masking raise."""
import numpy as np

from oracle import host_oracle


class _Spec:
  def __init__(self, name, dtype, shape):
    self.name, self.dtype, self.shape = name, np.dtype(dtype), tuple(shape)
    self.row_bytes = int(self.dtype.itemsize * np.prod(self.shape, dtype=np.int64))


class _Staging:
  def __init__(self, specs, rows):
    self.rows = rows
    self.views = {k: np.zeros((rows, *s.shape), s.dtype) for k, s in specs.items()}
    self.rowids_np = np.zeros(rows, np.int64)
    self.acts_host = {}

  def wait(self):
    pass


class HostStore:
  """Same interface as embodied_b200.core.store.DeviceStore, numpy inside."""

  def __init__(self, chunksize, staging_rows=256):
    self.chunksize = chunksize
    self.specs = None
    self.tables = {}
    self.nslabs = 0
    self._staging_rows = staging_rows
    self._turn = 0
    self.launches = []

  @property
  def configured(self):
    return self.specs is not None

  @property
  def bytes_per_row(self):
    return sum(s.row_bytes for s in self.specs.values())

  def configure(self, specs):
    self.specs = {k: _Spec(k, *v) for k, v in specs.items()}
    self._stagings = [_Staging(self.specs, self._staging_rows) for _ in range(2)]
    self.reserve(1)

  def reserve(self, nslabs):
    if nslabs <= self.nslabs:
      return
    rows = nslabs * self.chunksize
    for k, s in self.specs.items():
      new = np.zeros((rows, *s.shape), s.dtype)
      if k in self.tables:
        new[:len(self.tables[k])] = self.tables[k]
      self.tables[k] = new
    self.nslabs = nslabs

  def staging(self):
    return self._stagings[self._turn]

  def commit_staging(self, n, device_values=None):
    st = self._stagings[self._turn]
    self._turn ^= 1
    if n == 0:
      return
    rows = st.rowids_np[:n]
    # duplicates inside one launch would be a race on the device
    assert len(set(rows.tolist())) == n, 'two staged rows share a table row'
    self.launches.append(('append', n))
    for k in self.specs:
      src = (device_values or {}).get(k, st.views[k])
      self.tables[k][rows] = np.asarray(src)[:n]

  def gather(self, src_rows, batch, window, consec=None, annotate=True,
             keys=None, out=None):
    rows = np.asarray(src_rows, np.int64)
    self.launches.append(('gather', len(rows)))
    names = list(self.specs) if keys is None else keys
    data = {k: self.tables[k][rows].reshape(
        (batch, window, *self.specs[k].shape)) for k in names}
    if annotate:
      data = host_oracle.annotate(data)
    if consec is not None:
      data['consec'] = np.full((batch, window), consec, np.int32)
    return data

  def scatter(self, dst_rows, values):
    rows = np.asarray(dst_rows, np.int64)
    ok = rows >= 0
    self.launches.append(('scatter', len(rows)))
    for k, v in values.items():
      v = np.asarray(v).reshape((len(rows), *self.specs[k].shape))
      self.tables[k][rows[ok]] = v[ok]

  def export_slab(self, slab, length):
    lo = slab * self.chunksize
    return {k: self.tables[k][lo: lo + length].copy() for k in self.specs}

  def import_slab(self, slab, data):
    lo = slab * self.chunksize
    for k in self.specs:
      n = len(data[k])
      self.tables[k][lo: lo + n] = data[k]


class HostOps:
  def mask_actions(self, acts, is_last):
    return host_oracle.mask_actions(acts, is_last)
