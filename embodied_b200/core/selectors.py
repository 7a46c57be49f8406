"""Item selectors (embodied/core/selectors.py).

``Uniform`` is the one on the hot path (default ``replay.fracs.uniform: 1.0``,
dreamerv3/configs.yaml:42); its draw sequence is a bit-exact contract: numpy
``default_rng(seed).integers(0, n)`` over a key list with swap-with-last deletion
(selectors.py:29-57).

The non-uniform selectors (``Prioritized`` + ``SumTree``, ``Recency``, ``Mixture``;
selectors.py:60-229) are host-side index structures like ``Uniform``: they decide WHICH
window is gathered, the gather itself is the same ``emb_replay_gather`` launch.  They are
restated here on flat arrays instead of linked node objects, but the draw sequence under a
seed is the reference's (tests/test_selectors_host.py drives both side by side): the same
tree shape, the same child order after a removal, the same sums in the same order, one
``Generator.choice`` per level.
"""
import collections
import threading
import time

import numpy as np


class Uniform:

  def __init__(self, seed=0):
    self.indices = {}
    self.keys = []
    self.rng = np.random.default_rng(seed)
    self.lock = threading.Lock()

  def __len__(self):
    return len(self.keys)

  def __call__(self):
    with self.lock:
      return self.keys[self.rng.integers(0, len(self.keys)).item()]

  def draw(self, count):
    """`count` consecutive draws in one call.  numpy fills an array of bounded integers from the
    same 32-bit stream, value by value, as `count` scalar calls would (pinned by
    tests/test_selectors_host.py), so the sequence under a seed is unchanged."""
    with self.lock:
      keys = self.keys
      return [keys[i] for i in self.rng.integers(0, len(keys), size=count).tolist()]

  def __setitem__(self, key, stepids):
    with self.lock:
      self.indices[key] = len(self.keys)
      self.keys.append(key)

  def __delitem__(self, key):
    with self.lock:
      assert 2 <= len(self.keys), len(self.keys)
      hole = self.indices.pop(key)
      moved = self.keys.pop()
      if hole != len(self.keys):
        self.keys[hole] = moved
        self.indices[moved] = hole


class Fifo:
  """Oldest item first (embodied/core/selectors.py:7-26)."""

  def __init__(self):
    import collections
    self.queue = collections.deque()

  def __call__(self):
    return self.queue[0]

  def __len__(self):
    return len(self.queue)

  def __setitem__(self, key, stepids):
    self.queue.append(key)

  def __delitem__(self, key):
    if self.queue[0] == key:
      self.queue.popleft()
    else:
      self.queue.remove(key)


class SumTree:
  """Sampling proportional to an unnormalised weight per key (selectors.py:232-378).

  The reference grows a tree of node objects whose entries all sit at the same depth and
  fill the deepest level left to right; a removal takes the entry out of its leaf node's
  child list (later siblings shift left) and re-appends the tree's LAST entry to that
  node.  That shape is an implicit complete ``branching``-ary trie over a sequence of leaf
  buckets, so it is kept as such: ``buckets[k]`` = ordered keys of leaf node k (all full but
  the last), ``sums[l][k]`` = weight below node k of level l (0 = leaf nodes).  The depth only
  grows (a new root on top, selectors.py:258-261) and restarts when the tree runs empty.
  Weights are summed with the builtin ``sum`` over the children in order, fresh on every
  change, exactly like ``SampleTreeNode.recompute`` (:343-345): the probabilities handed to
  ``Generator.choice`` -- hence the draws -- are bit-identical.
  """

  def __init__(self, branching=16, seed=0):
    assert 2 <= branching
    self.branching = branching
    self.rng = np.random.default_rng(seed)
    self.buckets = []
    self.bucket_of = {}
    self.weight = {}
    self.depth = 1            # levels of nodes above the entries
    self.sums = [[0]]         # sums[l][k]; the top level may hold one node only
    self.root_sum = 0

  def __len__(self):
    return len(self.weight)

  @property
  def total(self):
    return self.sums[self.depth - 1][0] if self.buckets else 0

  def insert(self, key, uprob):
    B = self.branching
    if not self.buckets or len(self.buckets[-1]) >= B:
      if not self.buckets:
        self.depth = 1
      while len(self.buckets) + 1 > B ** (self.depth - 1):
        self.depth += 1       # a new root above the old one
      self.buckets.append([])
    self.buckets[-1].append(key)
    self.bucket_of[key] = len(self.buckets) - 1
    self.weight[key] = uprob
    self._resum(len(self.buckets) - 1)

  def remove(self, key):
    k = self.bucket_of.pop(key)
    del self.weight[key]
    last_bucket = self.buckets[-1]
    last_key = last_bucket[-1]
    self.buckets[k].remove(key)
    if last_key != key:
      # the tree's last entry moves to the END of the bucket that lost one
      last_bucket.remove(last_key) if k != len(self.buckets) - 1 else self.buckets[k].remove(last_key)
      self.buckets[k].append(last_key)
      self.bucket_of[last_key] = k
    if not self.buckets[-1]:
      self.buckets.pop()
    if not self.buckets:
      self.depth = 1
      self.sums = [[0]]
      return
    self._resum(k)
    self._resum(len(self.buckets) - 1)

  def update(self, key, uprob):
    self.weight[key] = uprob
    self._resum(self.bucket_of[key])

  def sample(self):
    node = 0
    for level in range(self.depth - 1, -1, -1):
      values = self._children(level, node)
      uprobs = np.array(values)
      total = uprobs.sum()
      if not np.isfinite(total):
        finite = np.isinf(uprobs)
        probs = finite / finite.sum()
      elif total == 0:
        probs = np.ones(len(uprobs)) / len(uprobs)
      else:
        probs = uprobs / total
      choice = self.rng.choice(np.arange(len(uprobs)), p=probs).item()
      if level == 0:
        return self.buckets[node][choice]
      node = node * self.branching + choice

  def _children(self, level, node):
    """Weights of the children of node `node` of level `level`, in order."""
    if level == 0:
      return [self.weight[x] for x in self.buckets[node]]
    below = self.sums[level - 1]
    return below[node * self.branching:(node + 1) * self.branching]

  def _resum(self, bucket):
    """Recompute the sums on the path from leaf node `bucket` to the root."""
    B = self.branching
    nbuckets = len(self.buckets)
    while len(self.sums) < self.depth:
      self.sums.append([])
    count = nbuckets
    for level in range(self.depth):
      row = self.sums[level]
      del row[count:]
      row.extend([0] * (count - len(row)))
      count = -(-count // B)
    node = bucket
    for level in range(self.depth):
      if node < len(self.sums[level]):
        self.sums[level][node] = sum(self._children(level, node))
      node //= B
    # nodes to the right of a shrunken tail were dropped above; the parents of the new tail
    # are on the path of `bucket == nbuckets - 1`, which every caller also resums


class Prioritized:
  """Priority per time step, aggregated per item (selectors.py:128-196): item weight =
  ``maxfrac * max + (1 - maxfrac) * mean`` of ``prio ** exponent`` over its steps; steps never
  prioritised count as ``initial``.  The arithmetic keeps the operand types it is given
  (float32 priorities from the learner stay float32), as the reference's does."""

  wants_stepids = True

  def __init__(self, exponent=1.0, initial=1.0, zero_on_sample=False, maxfrac=0.0, branching=16, seed=0):
    assert 0 <= maxfrac <= 1, maxfrac
    self.exponent = float(exponent)
    self.initial = float(initial)
    self.zero_on_sample = zero_on_sample
    self.maxfrac = maxfrac
    self.tree = SumTree(branching, seed)
    self.prios = collections.defaultdict(lambda: self.initial)
    self.users = collections.defaultdict(list)     # step id -> items containing it
    self.items = {}

  def __len__(self):
    return len(self.items)

  def __call__(self):
    key = self.tree.sample()
    if self.zero_on_sample:
      self.prioritize(self.items[key], [0.0] * len(self.items[key]))
    return key

  def __setitem__(self, key, stepids):
    stepids = self._as_bytes(stepids)
    self.items[key] = stepids
    for stepid in stepids:
      self.users[stepid].append(key)
    self.tree.insert(key, self._weight(key))

  def __delitem__(self, key):
    self.tree.remove(key)
    for stepid in self.items.pop(key):
      users = self.users[stepid]
      users.remove(key)
      if not users:
        del self.users[stepid]
        del self.prios[stepid]

  def prioritize(self, stepids, priorities):
    stepids = self._as_bytes(stepids)
    for stepid, priority in zip(stepids, priorities):
      self.prios[stepid] = priority
    touched = []
    for stepid in stepids:
      touched += self.users[stepid]
    for key in list(set(touched)):
      if key in self.items:
        self.tree.update(key, self._weight(key))

  @staticmethod
  def _as_bytes(stepids):
    return stepids if isinstance(stepids[0], bytes) else [x.tobytes() for x in stepids]

  def _weight(self, key):
    prios = [self.prios[stepid] for stepid in self.items[key]]
    if self.exponent != 1.0:
      prios = [x ** self.exponent for x in prios]
    mean = sum(prios) / len(prios)
    if self.maxfrac:
      return self.maxfrac * max(prios) + (1 - self.maxfrac) * mean
    return mean


class Recency:
  """Age-based selection (selectors.py:60-125): ``uprobs[a]`` = weight of the item inserted a
  insertions ago.  The age is drawn by descending a 16-ary table of conditional
  probabilities, one ``Generator.choice`` per level; while fewer items than ages exist the
  age is rescaled.  (The reference's ``_sample`` reads ``len(segment)`` before ``segment`` is
  bound at the first level -- selectors.py:109 -- and so cannot draw at all; the intended
  ``len(p)`` is used here, and tests/test_selectors_host.py pins everything else -- the
  table from ``_build`` and the draws of the corrected loop -- against the reference.)"""

  def __init__(self, uprobs, seed=0, bfactor=16):
    uprobs = np.asarray(uprobs)
    assert uprobs[0] >= uprobs[-1], uprobs
    self.uprobs = uprobs
    self.bfactor = bfactor
    self.table = self._conditional_table(uprobs, bfactor)
    self.rng = np.random.default_rng(seed)
    self.step = 0
    self.when = {}
    self.items = {}

  def __len__(self):
    return len(self.items)

  def __call__(self):
    for retry in range(10):
      try:
        age = self._draw_age()
        if len(self.items) < len(self.uprobs):
          age = int(age / len(self.uprobs) * len(self.items))
        return self.items[self.step - 1 - age]
      except KeyError:          # removed a moment ago
        if retry == 9:
          raise
        time.sleep(0.01)

  def __setitem__(self, key, stepids):
    self.when[key] = self.step
    self.items[self.step] = key
    self.step += 1

  def __delitem__(self, key):
    del self.items[self.when.pop(key)]

  def _draw_age(self):
    path = []
    for probs in self.table:
      p = probs[tuple(path)] if path else probs
      path.append(self.rng.choice(len(p), p=p))
    depth = len(self.table)
    return sum(index * self.bfactor ** (depth - level - 1) for level, index in enumerate(path))

  @staticmethod
  def _conditional_table(uprobs, bfactor):
    assert np.isfinite(uprobs).all(), uprobs
    assert (uprobs >= 0).all(), uprobs
    depth = int(np.ceil(np.log(len(uprobs)) / np.log(bfactor)))
    padded = np.concatenate([uprobs, np.zeros(bfactor ** depth - len(uprobs))])
    levels = [padded]
    for _ in range(depth - 1):
      levels.insert(0, levels[0].reshape((-1, bfactor)).sum(-1))
    table = []
    for level, mass in enumerate(levels):
      mass = mass.reshape([bfactor] * (1 + level))
      total = mass.sum(-1, keepdims=True)
      with np.errstate(divide='ignore', invalid='ignore'):
        table.append(np.where(total, mass / total, mass))
    return table


class Mixture:
  """One of several selectors per draw, chosen with fixed probabilities (selectors.py:199-229);
  selectors with fraction 0 are dropped, the rest are ordered by name."""

  wants_stepids = True

  def __init__(self, selectors, fractions, seed=0):
    assert set(selectors.keys()) == set(fractions.keys())
    assert sum(fractions.values()) == 1, fractions
    names = sorted(k for k in selectors if fractions[k])
    self.selectors = [selectors[k] for k in names]
    self.fractions = np.array([fractions[k] for k in names], np.float32)
    self.rng = np.random.default_rng(seed)

  def __call__(self):
    return self.selectors[self.rng.choice(len(self.selectors), p=self.fractions)]()

  def __len__(self):
    return len(self.selectors[0])

  def __setitem__(self, key, stepids):
    for selector in self.selectors:
      selector[key] = stepids

  def __delitem__(self, key):
    for selector in self.selectors:
      del selector[key]

  def prioritize(self, stepids, priorities):
    for selector in self.selectors:
      if hasattr(selector, 'prioritize'):
        selector.prioritize(stepids, priorities)
