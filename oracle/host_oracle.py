"""numpy restatement of the reference's host-side hot path (TEST INFRASTRUCTURE).

PARITY PINNED: ``tests/test_oracle_pinned.py`` drives this file and the real
reference (``oracle/refload.py``) with identical seeded transition streams and
requires byte-identical results; ``tests/golden/*.npz`` hold outputs of the real
reference for the machines where /root/reference is absent.

Restated here (reference file:line):
  OracleUniform   embodied/core/selectors.py:29-57
  OracleReplay    embodied/core/replay.py:77-118 (add), :121-127 (sample),
                  :130-149 (update), :151-169 (_sample), :171-191
                  (_insert/_remove), :193-235 (_getseq/_setseq), :256-292
                  (_assemble_batch/_annotate_batch), :362-370 (_complete);
                  embodied/core/chunk.py:13-62 (append/update/slice)
  consec_view     embodied/core/streams.py:120-140
  OracleDriver    embodied/core/driver.py:34-39 (reset), :55-87 (_step/_mask)
  normalize_image dreamerv3/rssm.py:228-231
Chunk ids are plain ints handed out by ``ids`` (an iterator), so a test can give
the oracle, the reference (UUID.reset(debug=True)) and the product the same ids.
"""
import collections
import itertools

import numpy as np


def stepid_bytes(uuid, index):
  # replay.py:90-91 -- 16 byte big-endian chunk id, then the row index as be32.
  return np.frombuffer(
      int(uuid).to_bytes(16, 'big') + int(index).to_bytes(4, 'big'), np.uint8)


class OracleUniform:
  """selectors.py:29-57."""

  def __init__(self, seed=0):
    self.slot = {}
    self.keys = []
    self.rng = np.random.default_rng(seed)          # :34

  def __len__(self):
    return len(self.keys)

  def draw(self):
    return self.keys[self.rng.integers(0, len(self.keys)).item()]   # :42-43

  def insert(self, key):
    self.slot[key] = len(self.keys)                 # :47-48
    self.keys.append(key)

  def remove(self, key):
    assert len(self.keys) >= 2                      # :52
    hole = self.slot.pop(key)
    tail = self.keys.pop()
    if hole != len(self.keys):                      # :55-57 swap-with-last
      self.keys[hole] = tail
      self.slot[tail] = hole


class _Chunk:
  def __init__(self, uuid, size):
    self.uuid, self.succ, self.size, self.length = uuid, 0, size, 0
    self.data = None


class OracleReplay:

  def __init__(self, length, capacity=None, chunksize=1024, online=False,
               seed=0, ids=None):
    self.length, self.capacity, self.chunksize = length, capacity, chunksize
    self.online = online
    self.ids = ids or itertools.count(1)
    self.sampler = OracleUniform(seed)
    self.chunks, self.refs = {}, {}
    self.items, self.fifo, self.next_item = {}, collections.deque(), 0
    self.current = {}
    self.streams = collections.defaultdict(collections.deque)
    self.lengths = collections.defaultdict(int)
    self.queue = collections.deque()

  def __len__(self):
    return len(self.items)

  # -- append ---------------------------------------------------------------
  def _new_chunk(self, refs):
    c = _Chunk(next(self.ids), self.chunksize)
    self.chunks[c.uuid] = c
    self.refs[c.uuid] = refs
    return c

  def add(self, step, worker=0):
    step = {k: np.asarray(v) for k, v in step.items()
            if not k.startswith('log/')}             # :78-80
    if worker not in self.current:                   # :82-87
      self.current[worker] = (self._new_chunk(1).uuid, 0)
    cid, idx = self.current[worker]
    step['stepid'] = stepid_bytes(cid, idx)          # :90-91
    chunk = self.chunks[cid]
    if chunk.data is None:                           # chunk.py:43-47
      chunk.data = {k: np.empty((chunk.size, *v.shape), v.dtype)
                    for k, v in step.items()}
    for k, v in step.items():                        # chunk.py:48-50
      chunk.data[k][chunk.length] = v
    chunk.length += 1
    stream = self.streams[worker]
    stream.append((cid, idx))                        # :97-99
    self.refs[cid] += 1
    if idx + 1 < chunk.size:                         # :101-105
      self.current[worker] = (cid, idx + 1)
    else:                                            # _complete :362-370
      succ = self._new_chunk(2)
      self.refs[cid] -= 1
      self.current[worker] = (succ.uuid, 0)
      chunk.succ = succ.uuid
    if len(stream) >= self.length:                   # :107-115
      start = stream.popleft()
      self._insert(*start)
      if self.online and self.lengths[worker] % self.length == 0:
        self.queue.append(start)
    if self.online:                                  # :117-118
      self.lengths[worker] += 1

  def _insert(self, cid, idx):                       # :171-179
    while self.capacity and len(self.items) >= self.capacity:
      self._evict()
    key = self.next_item
    self.next_item += 1
    self.items[key] = (cid, idx)
    self.sampler.insert(key)
    self.fifo.append(key)

  def _evict(self):                                  # :181-191
    key = self.fifo.popleft()
    self.sampler.remove(key)
    cid, _ = self.items.pop(key)
    self.refs[cid] -= 1
    if self.refs[cid] < 1:
      del self.refs[cid]
      gone = self.chunks.pop(cid)
      if gone.succ in self.refs:
        self.refs[gone.succ] -= 1

  # -- sample ---------------------------------------------------------------
  def _segments(self, cid, idx, count):
    """:193-214 / :216-235 -- (chunk, start, num) pieces of a window that may
    run over into successor chunks."""
    chunk = self.chunks[cid]
    take = min(count, chunk.length - idx)
    pieces = [(chunk, idx, take)]
    left = count - take
    while left > 0:
      chunk = self.chunks[chunk.succ]
      take = min(left, chunk.length)
      pieces.append((chunk, 0, take))
      left -= take
    return pieces

  def draw_windows(self, batch, mode='train'):
    """:151-169; returns [(chunkid, index)] * batch in draw order."""
    out = []
    while len(out) < batch:
      if self.online and self.queue and mode == 'train':
        cid, idx = self.queue.popleft()
      else:
        cid, idx = self.items[self.sampler.draw()]
      try:
        self._segments(cid, idx, self.length)
      except KeyError:                               # :168-169 retry
        continue
      out.append((cid, idx))
    return out

  def sample(self, batch, mode='train'):
    windows = self.draw_windows(batch, mode)
    first = self._segments(*windows[0], self.length)[0][0]
    data = {k: np.empty((batch, self.length, *v.shape[1:]), v.dtype)
            for k, v in first.data.items()}          # :257-260
    for n, (cid, idx) in enumerate(windows):         # :261-274
      t = 0
      for chunk, start, num in self._segments(cid, idx, self.length):
        for k in data:
          data[k][n, t: t + num] = chunk.data[k][start: start + num]
        t += num
    return annotate(data)

  # -- latent write-back ----------------------------------------------------
  def update(self, data):                            # :130-149
    data = dict(data)
    stepid = data.pop('stepid')
    assert stepid.ndim == 3
    for n in range(len(stepid)):
      raw = stepid[n, 0].tobytes()
      cid = int.from_bytes(raw[:16], 'big')
      idx = int.from_bytes(raw[16:], 'big')
      rows = {k: v[n] for k, v in data.items()}
      count = len(next(iter(rows.values())))
      try:
        pieces = self._segments(cid, idx, count)
      except KeyError:                               # evicted: skipped :148-149
        # The reference writes the leading pieces it could still reach before
        # the KeyError fires (chunk.update happens piece by piece, :224-235).
        pieces = self._reachable(cid, idx, count)
      t = 0
      for chunk, start, num in pieces:               # chunk.py:54-58
        for k, v in rows.items():
          chunk.data[k][start: start + num] = v[t: t + num]
        t += num

  def _reachable(self, cid, idx, count):
    if cid not in self.chunks:
      return []
    chunk = self.chunks[cid]
    take = min(count, chunk.length - idx)
    pieces, left = [(chunk, idx, take)], count - take
    while left > 0 and chunk.succ in self.chunks:
      chunk = self.chunks[chunk.succ]
      take = min(left, chunk.length)
      pieces.append((chunk, 0, take))
      left -= take
    return pieces


def annotate(data):
  """replay.py:278-292 with is_first=True."""
  data = dict(data)
  if 'is_first' in data:
    first = data['is_first'].copy()
    first[:, 0] = True
    data['is_first'] = first
    if 'is_last' in data:
      nxt = np.zeros_like(first)
      nxt[:, :-1] = first[:, 1:]
      data['is_last'] = data['is_last'] | nxt
  return data


def consec_view(batch, length, index, prefix):
  """streams.py:120-140 -- the ``index``-th consecutive sub-batch, contiguous."""
  lo = index * length
  hi = lo + length + prefix
  out = {k: np.ascontiguousarray(v[:, lo: hi]) for k, v in batch.items()}
  out['consec'] = np.full(out['is_first'].shape, index, np.int32)
  return out


def normalize_image(img_u8):
  """dreamerv3/rssm.py:230 -- uint8 -> float32, x / 255 - 0.5."""
  return img_u8.astype(np.float32) / np.float32(255) - np.float32(0.5)


def mask_actions(acts, is_last):
  """driver.py:72-75,84-87."""
  if not is_last.any():
    return dict(acts)
  keep = ~is_last
  out = {}
  for k, v in acts.items():
    m = keep.reshape(keep.shape + (1,) * (v.ndim - keep.ndim))
    out[k] = v * m.astype(v.dtype)
  return out


class OracleDriver:
  """driver.py serial mode: reset :34-39, one step :55-82."""

  def __init__(self, envs, act_space):
    self.envs, self.act_space = envs, act_space
    self.callbacks = []
    self.reset()

  def reset(self, init_policy=None):
    n = len(self.envs)
    self.acts = {k: np.zeros((n,) + tuple(s.shape), s.dtype)
                 for k, s in self.act_space.items()}
    self.acts['reset'] = np.ones(n, bool)
    self.carry = init_policy and init_policy(n)

  def step(self, policy):
    n = len(self.envs)
    obs = [env.step({k: v[i] for k, v in self.acts.items()})
           for i, env in enumerate(self.envs)]                       # :59-64
    obs = {k: np.stack([o[k] for o in obs]) for k in obs[0]}         # :65
    logs = {k: v for k, v in obs.items() if k.startswith('log/')}
    obs = {k: v for k, v in obs.items() if not k.startswith('log/')}
    self.carry, acts, outs = policy(self.carry, obs)                 # :69
    acts = mask_actions(acts, obs['is_last'])                        # :72-74
    self.acts = {**acts, 'reset': obs['is_last'].copy()}             # :75
    trans = {**obs, **acts, **outs, **logs}        # :76 (no 'reset' key: the
    # reference merges the masked policy acts, not self.acts; its own
    # test_driver.py:57 still expects 'reset' and is stale -- code wins)
    for i in range(n):                                               # :77-79
      row = {k: v[i] for k, v in trans.items()}
      for fn in self.callbacks:
        fn(row, i)
    return trans
