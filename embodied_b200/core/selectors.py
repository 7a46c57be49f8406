"""Item selectors.  Only ``Uniform`` is on the hot path (default
``replay.fracs.uniform: 1.0``, dreamerv3/configs.yaml:42); its draw sequence is
a bit-exact contract: numpy ``default_rng(seed).integers(0, n)`` over a key list
with swap-with-last deletion (embodied/core/selectors.py:29-57).
"""
import threading

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
