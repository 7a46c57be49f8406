"""Replay buffer with HBM-resident rows and host-side index bookkeeping.

Same surface and semantics as the reference ``embodied.core.Replay``
(embodied/core/replay.py:14-394): ``Replay(length, capacity, directory,
chunksize, online, selector, save_wait, name, seed)``, ``len()``, ``add``,
``sample``, ``update``, ``stats``, ``save``, ``load``.  What differs is where
bytes live: transition rows are stored in device tables (core/store.py) and
``sample`` returns dense *device* tensors ``(B, L, ...)``; every byte is moved
by the row engine of libembodied_b200.so.  The host keeps only the index
structures, in exactly the reference's order of operations so that sampling is
bit-exact for a given seed and call sequence:

  chunks / refs / succ links     replay.py:82-105,181-191,362-370
  per-worker stream -> items     replay.py:107-118,171-179
  FIFO eviction                  replay.py:172-173,181-191
  Uniform selector               selectors.py:29-57 (numpy default_rng(seed))
  online queue                   replay.py:114-118,158-160

Extra (not in the reference): ``add_batch`` (N workers' rows in one launch,
values may already be device tensors) and ``sample(..., consec=...)`` used by
``streams.Consec`` to gather sub-windows straight from the tables.
"""
import collections
import concurrent.futures
import io
import threading

import numpy as np

from .. import elements
from . import limiters
from . import selectors


class Chunk:
  """Index record of one slab (reference embodied/core/chunk.py:9-39)."""

  __slots__ = ('time', 'uuid', 'succ', 'length', 'size', 'slab', 'saved',
               'idbytes')

  def __init__(self, size, slab, uuid=None):
    self.time = elements.timestamp(millis=True)
    self.uuid = elements.UUID() if uuid is None else uuid
    self.succ = elements.UUID(0)
    self.length = 0
    self.size = size
    self.slab = slab
    self.saved = False
    self.idbytes = np.frombuffer(bytes(self.uuid), np.uint8)

  @property
  def filename(self):
    # chunk.py:29-33  {time}-{uuid}-{succ}-{length}.npz
    return f'{self.time}-{self.uuid}-{self.succ}-{self.length}.npz'

  def __repr__(self):
    return f'Chunk({self.filename}, slab={self.slab})'


class Replay:

  def __init__(
      self, length, capacity=None, directory=None, chunksize=1024,
      online=False, selector=None, save_wait=False, name='unnamed', seed=0,
      store=None, device=None, staging_rows=256, workers=1):
    self.length = int(length)
    self.capacity = capacity and int(capacity)
    self.chunksize = int(chunksize)
    self.name = name
    self.sampler = selector if selector is not None else selectors.Uniform(seed)

    self.chunks = {}      # UUID -> Chunk
    self.refs = {}        # UUID -> int
    self.items = {}       # itemid -> (chunk uuid, index)
    self.fifo = collections.deque()
    self.itemid = 0
    self.current = {}     # worker -> (chunk uuid, index)
    self.streams = collections.defaultdict(collections.deque)

    self.online = online
    if online:
      self.lengths = collections.defaultdict(int)
      self.queue = collections.deque()

    if directory:
      self.directory = elements.Path(directory)
      self.directory.mkdir()
      self.workers = concurrent.futures.ThreadPoolExecutor(16, 'replay_saver')
      self.saved = set()
    else:
      self.directory = None
    self.save_wait = save_wait
    self.metrics = {'samples': 0, 'inserts': 0, 'updates': 0}

    if store is None:
      from . import store as storelib
      store = storelib.DeviceStore(
          self.chunksize, device=device, staging_rows=staging_rows)
    if int(store.chunksize) != self.chunksize:
      raise ValueError(f'store.chunksize={store.chunksize} != chunksize={self.chunksize}: one slab '
                       'of the store backs exactly one chunk')
    self.store = store
    self._workers_hint = int(workers)
    self._lock = threading.RLock()
    self._free = []           # slabs ready for a new chunk
    self._limbo = []          # slabs freed since the last flush (see _flush)
    self._pending = 0         # rows staged on the host, not yet launched
    self._batch_open = False  # add_batch is filling a staging block
    self._staging = None
    self._recent = collections.deque(maxlen=4)   # sampled stepid tensors

  # ------------------------------------------------------------------ basics
  def __len__(self):
    return len(self.items)

  def stats(self):
    ratio = lambda x, y: x / y if y else np.nan
    m = self.metrics
    nbytes = len(self.chunks) * self.chunksize * (
        self.store.bytes_per_row if self.store.configured else 0)
    stats = {
        'items': len(self.items),
        'chunks': len(self.chunks),
        'streams': len(self.streams),
        'ram_gb': nbytes / (1024 ** 3),
        'inserts': m['inserts'],
        'samples': m['samples'],
        'updates': m['updates'],
        'replay_ratio': ratio(self.length * m['samples'], m['inserts']),
    }
    for key in m:
      m[key] = 0
    return stats

  # ------------------------------------------------------------ slab handling
  def _plan_slabs(self):
    if self.capacity:
      per = -(-(self.capacity + self.length) // self.chunksize) + 1
      return per + 2 * self._workers_hint
    return 4 * self._workers_hint

  def _take_slab(self):
    if not self._free and self._limbo and not self._batch_open:
      self._flush()
    if not self._free:
      old = self.store.nslabs
      new = max(old * 2, old + self._workers_hint, self._plan_slabs())
      self.store.reserve(new)
      self._free.extend(range(new - 1, old - 1, -1))
    return self._free.pop()

  def _new_chunk(self, refs, uuid=None, size=None):
    chunk = Chunk(size or self.chunksize, self._take_slab(), uuid)
    self.chunks[chunk.uuid] = chunk
    self.refs[chunk.uuid] = refs
    return chunk

  def _drop_chunk(self, uuid):
    chunk = self.chunks.pop(uuid)
    # The slab may still be the target of staged rows or the source of a
    # gather that is queued but not launched; it becomes reusable at the next
    # flush, after which stream order protects it.
    self._limbo.append(chunk.slab)
    return chunk

  def _configure(self, step):
    specs = {k: (v.dtype, v.shape) for k, v in step.items()}
    specs['stepid'] = (np.uint8, (20,))
    self.store.configure(specs)
    self.store.reserve(self._plan_slabs())
    self._free = list(range(self.store.nslabs - 1, -1, -1))

  # --------------------------------------------------------------------- add
  @elements.timer.section('replay_add')
  def add(self, step, worker=0):
    """One transition of one worker (reference replay.py:77-118)."""
    step = {k: np.asarray(v) for k, v in step.items()
            if not k.startswith('log/')}
    with self._lock:
      if not self.store.configured:
        self._configure(step)
      chunk, index = self._slot(worker)
      if self._staging is None:
        self._staging = self.store.staging()
      st, row = self._staging, self._pending
      views = st.views
      if len(step) + 1 != len(views):
        raise KeyError(
            f'transition keys {sorted(step)} != stored {sorted(views)}')
      for key, value in step.items():
        views[key][row] = value
      sid = views['stepid'][row]
      sid[:16] = chunk.idbytes                       # replay.py:90-91
      sid[16:] = np.frombuffer(index.to_bytes(4, 'big'), np.uint8)
      st.rowids_np[row] = chunk.slab * self.chunksize + index
      self._pending += 1
      if self._pending == st.rows:
        self._flush()
      self._advance(worker, chunk, index)

  def add_batch(self, trans, workers=None):
    """Transitions of N workers at once.  `trans`: {key: (N, ...)} of numpy
    arrays and/or CUDA tensors (device values never visit the host).
    Index bookkeeping runs worker by worker in order, i.e. exactly as N calls of
    add() would (reference Driver callback order, driver.py:77-79)."""
    import torch
    trans = {k: v for k, v in trans.items() if not k.startswith('log/')}
    n = len(next(iter(trans.values())))
    workers = range(n) if workers is None else workers
    if not self.store.configured:
      self._workers_hint = max(self._workers_hint, n)     # planned before the first reserve
    with self._lock:
      self._flush()
      if not self.store.configured:
        self._configure({
            k: (v[0].cpu().numpy() if isinstance(v, torch.Tensor)
                else np.asarray(v[0])) for k, v in trans.items()})
      done = 0
      while done < n:
        st = self.store.staging()
        m = min(st.rows, n - done)
        dev = {}
        for key, value in trans.items():
          if isinstance(value, torch.Tensor) and value.is_cuda:
            dev[key] = value[done: done + m]
          else:
            st.views[key][:m] = value[done: done + m]
        sid = st.views['stepid']
        self._batch_open = True   # no slab may be recycled into this launch
        for j in range(m):
          worker = workers[done + j]
          chunk, index = self._slot(worker)
          sid[j, :16] = chunk.idbytes
          sid[j, 16:] = np.frombuffer(index.to_bytes(4, 'big'), np.uint8)
          st.rowids_np[j] = chunk.slab * self.chunksize + index
          self._advance(worker, chunk, index)
        self._batch_open = False
        self.store.commit_staging(m, dev)
        self._recycle()
        done += m

  # -- three-phase add used by the fused Driver step -------------------------
  def configure_spaces(self, obs_space, act_space, ext_space=None, workers=None):
    """Fix the row layout up front from spaces (obs minus log/, actions minus
    reset, the agent's replay-context entries) instead of from a first row.
    `workers`: how many streams will append (the Driver's env count): every one
    holds a current chunk, so the slab pool is planned for them up front."""
    if workers:
      self._workers_hint = max(self._workers_hint, int(workers))
    specs = {}
    for k, s in obs_space.items():
      if not k.startswith('log/'):
        specs[k] = (s.dtype, s.shape)
    for k, s in act_space.items():
      if k != 'reset':
        specs[k] = (s.dtype, s.shape)
    for k, s in (ext_space or {}).items():
      if k not in ('consec', 'stepid'):        # formed by the replay itself (replay.py:90-91, streams.py:118)
        specs[k] = (s.dtype, s.shape)
    with self._lock:
      self._configure({k: np.zeros(sh, dt) for k, (dt, sh) in specs.items()})

  def open_batch(self, n, workers=None):
    """Phase 1: a pinned staging block to stack N observations into, with the
    table row of every worker's next step already resolved."""
    with self._lock:
      self._flush()
      st = self.store.staging()
      if n > st.rows:
        raise ValueError(f'{n} envs > staging_rows={st.rows}')
      st.n = n
      st.workers = list(range(n) if workers is None else workers)
      st.slots = []
      sid = st.views['stepid']
      self._batch_open = True
      for j, worker in enumerate(st.workers):
        chunk, index = self._slot(worker)
        sid[j, :16] = chunk.idbytes
        sid[j, 16:] = np.frombuffer(index.to_bytes(4, 'big'), np.uint8)
        st.rowids_np[j] = chunk.slab * self.chunksize + index
        st.slots.append((chunk, index))
    return st

  def stage_obs(self, st, obs_keys, norm_keys=None):
    """Phase 2: one H2D + emb_driver_stage_obs.  Returns the observation dict
    of device tensors for the policy; uint8 images also come normalised
    (x/255-0.5, float32) under ``obs.normalized[key]``."""
    if norm_keys is None:
      norm_keys = [k for k in obs_keys
                   if self.store.specs[k].dtype == np.uint8 and
                   len(self.store.specs[k].shape) == 3]
    obs, normed = self.store.stage_obs(st, st.n, obs_keys, norm_keys)
    out = DeviceObs(obs)
    out.normalized = normed
    return out

  def commit_batch(self, st, acts, outs):
    """Phase 3: emb_driver_scatter_mask_actions, then the index bookkeeping of
    N add() calls in worker order.  Returns the masked actions on the host."""
    with self._lock:
      is_last = self.store.device_view(st, 'is_last', st.n)
      host_acts = self.store.commit_acts(st, st.n, acts, outs, is_last)
      self.store._turn ^= 1
      for worker, (chunk, index) in zip(st.workers, st.slots):
        self._advance(worker, chunk, index)
      self._batch_open = False
      self._recycle()
    return host_acts

  def _slot(self, worker):
    if worker not in self.current:                   # replay.py:82-87
      chunk = self._new_chunk(1)
      self.current[worker] = (chunk.uuid, 0)
    uuid, index = self.current[worker]
    return self.chunks[uuid], index

  def _advance(self, worker, chunk, index):
    """Index bookkeeping after row `index` of `chunk` was written
    (replay.py:93-118)."""
    uuid = chunk.uuid
    assert chunk.length == index, (chunk.length, index)
    chunk.length = index + 1
    stream = self.streams[worker]
    stream.append((uuid, index))
    self.refs[uuid] += 1
    if index + 1 < chunk.size:
      self.current[worker] = (uuid, index + 1)
    else:
      self._complete(chunk, worker)
    if len(stream) >= self.length:
      self.metrics['inserts'] += 1
      start = stream.popleft()
      self._insert(*start)
      if self.online and self.lengths[worker] % self.length == 0:
        self.queue.append(start)
    if self.online:
      self.lengths[worker] += 1

  def _complete(self, chunk, worker):                # replay.py:362-370
    succ = self._new_chunk(2)
    self.refs[chunk.uuid] -= 1
    self.current[worker] = (succ.uuid, 0)
    chunk.succ = succ.uuid
    return succ

  def _insert(self, uuid, index):                    # replay.py:171-179
    while self.capacity and len(self.items) >= self.capacity:
      self._remove()
    itemid = self.itemid
    self.itemid += 1
    self.items[itemid] = (uuid, index)
    # Uniform / Recency ignore the step ids (selectors.py:45-48,93-96); priority-based selectors
    # key their priorities by them (selectors.py:172-177): formed on the host, no device read
    wants = getattr(self.sampler, 'wants_stepids', False)
    self.sampler[itemid] = self._window_stepids(uuid, index) if wants else None
    self.fifo.append(itemid)

  def _window_stepids(self, uuid, index):
    """The `length` step ids of the window starting at (chunk, index), as the 20-byte strings
    the table holds: 16 B chunk uuid | be32 index (replay.py:90-91), following succ links."""
    out, left = [], self.length
    while left > 0:
      chunk = self.chunks[uuid]
      prefix = bytes(uuid)
      take = min(left, chunk.length - index)
      out += [prefix + (index + i).to_bytes(4, 'big') for i in range(take)]
      left -= take
      uuid, index = chunk.succ, 0
    return out

  def _remove(self):                                 # replay.py:181-191
    itemid = self.fifo.popleft()
    del self.sampler[itemid]
    uuid, _ = self.items.pop(itemid)
    self.refs[uuid] -= 1
    if self.refs[uuid] < 1:
      del self.refs[uuid]
      chunk = self._drop_chunk(uuid)
      if chunk.succ in self.refs:
        self.refs[chunk.succ] -= 1

  def _flush(self):
    """Launch the staged rows, then let freed slabs be reused."""
    if self._pending:
      self.store.commit_staging(self._pending)
      self._pending = 0
      self._staging = None
    self._recycle()

  def _recycle(self):
    if self._limbo:
      self._free.extend(self._limbo)
      self._limbo.clear()

  # ------------------------------------------------------------------ sample
  def _rows_of(self, uuid, index, count):
    """Table rows of `count` steps starting at (chunk, index), following succ
    links (replay.py:193-214).  KeyError if a chunk is gone."""
    chunk = self.chunks[uuid]
    base = chunk.slab * self.chunksize
    avail = chunk.length - index
    if avail >= count:
      return np.arange(base + index, base + index + count, dtype=np.int64)
    parts = [np.arange(base + index, base + index + avail, dtype=np.int64)]
    left = count - avail
    while left > 0:
      chunk = self.chunks[chunk.succ]
      used = min(left, chunk.length)
      base = chunk.slab * self.chunksize
      parts.append(np.arange(base, base + used, dtype=np.int64))
      left -= used
    return np.concatenate(parts)

  def _draw(self, mode):                             # replay.py:151-169
    assert mode in ('train', 'report', 'eval'), mode
    if mode == 'train':
      self.metrics['samples'] += 1
    while True:
      try:
        if self.online and self.queue and mode == 'train':
          uuid, index = self.queue.popleft()
        else:
          uuid, index = self.items[self.sampler()]
        return (uuid, index), self._rows_of(uuid, index, self.length)
      except KeyError:
        continue

  def _draw_many(self, batch, mode):
    """`batch` windows and their (batch * length) table rows.  Same attempts in the same order
    as `batch` calls of _draw (online queue first, then one selector draw per attempt, an
    attempt whose chunk is gone is skipped: replay.py:151-169) -- but a selector that can draw
    in bulk (`Uniform.draw`) is asked once per round, and windows that lie inside one chunk
    (all but the few straddling a chunk end) become rows by one broadcast add."""
    bulk = getattr(self.sampler, 'draw', None)
    if bulk is None:
      picks = [self._draw(mode) for _ in range(batch)]
      return [p[0] for p in picks], np.concatenate([p[1] for p in picks])
    assert mode in ('train', 'report', 'eval'), mode
    L, cs, chunks, items = self.length, self.chunksize, self.chunks, self.items
    wins, starts, avails, nexts, odd = [], [], [], [], []
    if mode == 'train':
      self.metrics['samples'] += batch
    while len(wins) < batch:
      need = batch - len(wins)
      if self.online and self.queue and mode == 'train':
        attempts = [self.queue.popleft() for _ in range(min(need, len(self.queue)))]
        keyed = False
      else:
        attempts, keyed = bulk(need), True
      for attempt in attempts:
        try:
          uuid, index = items[attempt] if keyed else attempt
          chunk = chunks[uuid]
          avail = chunk.length - index
          if avail >= L:                                   # the window lies inside one chunk
            nxt = 0
          else:
            succ = chunks[chunk.succ]
            if succ.length >= L - avail:                   # ... or ends in its successor
              nxt = succ.slab * cs
            else:                                          # several hops (short chunks): the general walk
              odd.append((len(wins), self._rows_of(uuid, index, L)))
              avail, nxt = L, 0
          starts.append(chunk.slab * cs + index)
          avails.append(avail)
          nexts.append(nxt)
          wins.append((uuid, index))
        except KeyError:
          continue
    t = np.arange(L, dtype=np.int64)
    rows = np.asarray(starts, np.int64)[:, None] + t
    if min(avails) < L:
      avail = np.asarray(avails, np.int64)[:, None]
      rows = np.where(t < avail, rows, np.asarray(nexts, np.int64)[:, None] + (t - avail))
    for at, r in odd:
      rows[at] = r
    return wins, rows.reshape(-1)

  @elements.timer.section('replay_sample')
  def sample(self, batch, mode='train', consec=None):
    """Dense device batch {key: (B, L, ...)} incl. stepid u8[B, L, 20]
    (replay.py:121-127 + _assemble_batch :256-275 + _annotate_batch :278-292).
    `consec=i` also emits the int32 'consec' key of streams.Consec
    (streams.py:134) from the same launch."""
    limiters.wait(
        lambda: len(self.sampler), f'Replay buffer {self.name} is empty')
    with self._lock:
      wins, rows = self._draw_many(batch, mode)
      self._flush()
      data = self.store.gather(rows, batch, self.length, consec=consec)
      self._recent.append((data['stepid'], wins))
    return data

  # ------------------------------------------------------------------ update
  @elements.timer.section('replay_update')
  def update(self, data):
    """Write (B, T, ...) values back to the rows named by stepid[:, 0]
    (replay.py:130-149; evicted chunks are skipped)."""
    data = dict(data)
    stepid = data.pop('stepid')
    priority = data.pop('priority', None)
    assert len(stepid.shape) == 3, stepid.shape
    self.metrics['updates'] += int(np.prod(stepid.shape[:-1]))
    if priority is not None:
      self.sampler.prioritize(
          np.asarray(_to_host(stepid)).reshape((-1, stepid.shape[-1])),
          np.asarray(_to_host(priority)).flatten())
    if not data:
      return
    B, T = int(stepid.shape[0]), int(stepid.shape[1])
    with self._lock:
      firsts = self._first_steps(stepid)
      rows = np.full((B, T), -1, np.int64)
      for b, (uuid, index) in enumerate(firsts):
        try:
          rows[b] = self._rows_of(uuid, index, T)
        except KeyError:
          # reference: KeyError from the first chunk -> row skipped entirely;
          # a missing successor leaves the leading piece written (:224-235).
          rows[b] = self._rows_reachable(uuid, index, T)
      flat = rows.reshape(-1)
      # Overlapping windows name the same table row more than once.  The
      # reference applies rows b = 0..B-1 in order, so the LAST writer wins;
      # inside one launch that must be made explicit.
      _, at = np.unique(flat[::-1], return_index=True)
      keep = np.zeros(len(flat), bool)
      keep[len(flat) - 1 - at] = True
      flat = np.where(keep, flat, -1)
      self._flush()
      self.store.scatter(flat, data)

  def _rows_reachable(self, uuid, index, count):
    out = np.full(count, -1, np.int64)
    chunk = self.chunks.get(uuid)
    t = 0
    while chunk is not None and t < count:
      start = index if t == 0 else 0
      used = min(count - t, chunk.length - start)
      base = chunk.slab * self.chunksize + start
      out[t: t + used] = np.arange(base, base + used)
      t += used
      chunk = self.chunks.get(chunk.succ)
    return out

  def _first_steps(self, stepid):
    """(chunk uuid, index) of stepid[b, 0] for every b, without a device sync
    when `stepid` is a view of a batch this buffer sampled."""
    import torch
    if isinstance(stepid, torch.Tensor) and stepid.is_cuda:
      B, T = stepid.shape[0], stepid.shape[1]
      for base, wins in self._recent:
        if (stepid.untyped_storage().data_ptr() ==
            base.untyped_storage().data_ptr() and
            stepid.stride() == base.stride() and B == base.shape[0]):
          off = stepid.storage_offset() - base.storage_offset()
          t0, rem = divmod(off, 20)
          if rem == 0 and 0 <= t0 and t0 + T <= base.shape[1]:
            return [self._shift(w, t0) for w in wins]
      stepid = stepid[:, 0].cpu().numpy()
    else:
      stepid = np.asarray(stepid)[:, 0]
    out = []
    for row in stepid:
      raw = row.tobytes()
      out.append((elements.UUID(raw[:16]), int.from_bytes(raw[16:], 'big')))
    return out

  def _shift(self, window, steps):
    uuid, index = window
    while steps:
      chunk = self.chunks.get(uuid)
      if chunk is None:
        return (uuid, index + steps)    # gone: _rows_of raises KeyError later
      # a chunk that save() completed early is shorter than its slab: its
      # successor starts right after `length`, not after `size`
      full = chunk.length if chunk.succ != elements.UUID(0) else chunk.size
      room = full - index
      if steps < room:
        return (uuid, index + steps)
      steps -= room
      uuid, index = chunk.succ, 0
    return (uuid, index)

  # ------------------------------------------------------------- save / load
  @elements.timer.section('replay_save')
  def save(self):
    """Completed + current chunks -> `{time}-{uuid}-{succ}-{length}.npz`
    (replay.py:295-309, chunk.py:64-74)."""
    if self.directory:
      with self._lock:
        self._flush()
        for worker, (uuid, _) in list(self.current.items()):
          chunk = self.chunks[uuid]
          if chunk.length > 0:
            self._complete(chunk, worker)
        promises = []
        for chunk in list(self.chunks.values()):
          if chunk.length > 0 and chunk.uuid not in self.saved:
            self.saved.add(chunk.uuid)
            rows = self.store.export_slab(chunk.slab, chunk.length)
            name = chunk.filename
            chunk.saved = True
            promises.append(self.workers.submit(
                _write_npz, self.directory / name, rows))
        if self.save_wait:
          [p.result() for p in promises]
    return None

  @elements.timer.section('replay_load')
  def load(self, data=None, directory=None, amount=None):
    """Newest chunks first, up to `amount` items (replay.py:312-359)."""
    directory = directory or self.directory
    amount = amount or self.capacity or np.inf
    if not directory:
      return
    directory = elements.Path(directory)
    newest_first = lambda names: sorted(names, reverse=True)
    loaded = newest_first(c.filename for c in list(self.chunks.values()))
    ondisk = newest_first(p.name for p in directory.glob('*.npz'))
    ondisk = [n for n in ondisk if n not in loaded]
    if not ondisk:
      return
    numitems = self._numitems(loaded + ondisk)
    total, numchunks = 0, 0
    for name in ondisk:
      numchunks += 1
      total += numitems[elements.UUID(name.split('-')[1])]
      if total >= amount:
        break
    with concurrent.futures.ThreadPoolExecutor(16, 'replay_loader') as pool:
      read = list(pool.map(_read_npz, [directory / n for n in ondisk[:numchunks]]))
    read = [r for r in read if r is not None]
    numitems = self._numitems([name for name, _ in read])
    with self._lock:
      self._flush()
      chunks = []
      for name, arrays in read:
        time, uuid, succ, length = elements.Path(name).stem.split('-')
        length = int(length)
        if length > self.chunksize:
          raise ValueError(
              f'chunk {name} holds {length} steps but this buffer was created with '
              f'chunksize={self.chunksize}; load it with chunksize >= {length}')
        if not self.store.configured:
          first = {k: v[0] for k, v in arrays.items() if k != 'stepid'}
          self._configure(first)
        chunk = self._new_chunk(0, elements.UUID(uuid), size=length)
        chunk.time, chunk.succ = time, elements.UUID(succ)
        chunk.length, chunk.saved = length, True
        self.store.import_slab(chunk.slab, arrays)
        chunks.append(chunk)
      self.saved.update(c.uuid for c in chunks)
      for chunk in reversed(chunks):
        count = int(numitems[chunk.uuid])
        self.refs[chunk.uuid] += count
        if chunk.succ in self.refs:
          self.refs[chunk.succ] += 1
        for index in range(count):
          self._insert(chunk.uuid, index)

  def _numitems(self, names):                        # replay.py:372-388
    if not names:
      return {}
    stems = sorted((elements.Path(n).stem for n in names), reverse=True)
    fields = [s.split('-') for s in stems]
    uuids = [elements.UUID(f[1]) for f in fields]
    succs = [elements.UUID(f[2]) for f in fields]
    lengths = {u: int(f[3]) for u, f in zip(uuids, fields)}
    future = {}
    for u, s in zip(uuids, succs):
      future[u] = lengths[u] + future.get(s, 0)
    out = {}
    for u, s in zip(uuids, succs):
      n = lengths[u] + 1 - self.length + future.get(s, 0)
      out[u] = int(np.clip(n, 0, lengths[u]))
    return out


class DeviceObs(dict):
  """Observation dict of device tensors; `.normalized` holds float32
  x/255-0.5 versions of the uint8 image keys (dreamerv3/rssm.py:230)."""
  normalized = None


def _to_host(x):
  return x.cpu().numpy() if hasattr(x, 'cpu') else x


def _write_npz(filename, arrays):
  with io.BytesIO() as stream:
    np.savez_compressed(stream, **arrays)
    stream.seek(0)
    elements.Path(filename).write(stream.read(), mode='wb')


def _read_npz(filename):
  try:
    with open(filename, 'rb') as f:
      data = np.load(f)
      return filename.name, {k: data[k] for k in data.keys()}
  except Exception as e:   # corrupt chunk files are skipped (chunk.py:85-91)
    print(f'Error loading chunk {filename}: {e}')
    return None
