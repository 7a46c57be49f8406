"""Mini stand-in for the third-party ``elements`` package (TEST INFRASTRUCTURE).

Written from scratch; covers only the surface the reference's
``embodied/core/*.py`` and ``embodied/envs/dummy.py`` use:
UUID, timestamp, Path, timer.section, RWLock, tree.map, Space.
The one value-bearing behaviour is ``bytes(UUID)`` = 16 bytes big-endian, which
fixes the stepid layout ``uuid16 || be32(index)`` (reference
``embodied/core/replay.py:90-91``).
"""
import contextlib
import datetime
import itertools
import pathlib
import string
import threading
import types
import uuid as uuidlib

import numpy as np


# --------------------------------------------------------------------- UUID

class UUID:

  __slots__ = ('value', '_hash')
  DEBUG_ID = None
  _LOCK = threading.Lock()
  _ALPHABET = string.digits + string.ascii_letters
  _REV = {c: i for i, c in enumerate(_ALPHABET)}

  @classmethod
  def reset(cls, *, debug):
    cls.DEBUG_ID = itertools.count(1) if debug else None

  def __init__(self, value=None):
    if value is None:
      if self.DEBUG_ID is None:
        value = uuidlib.uuid4().int
      else:
        with self._LOCK:
          value = next(type(self).DEBUG_ID)
    elif isinstance(value, UUID):
      value = value.value
    elif isinstance(value, (int, np.integer)):
      value = int(value)
    elif isinstance(value, str):
      if value.isdigit() and self.DEBUG_ID is not None:
        value = int(value)
      else:
        acc = 0
        for c in value:
          acc = acc * 62 + self._REV[c]
        value = acc
    elif isinstance(value, np.ndarray):
      value = int.from_bytes(value.tobytes(), 'big')
    elif isinstance(value, (bytes, bytearray)):
      value = int.from_bytes(bytes(value), 'big')
    else:
      raise ValueError(value)
    assert 0 <= value < 2 ** 128, value
    self.value = value
    self._hash = hash(value)

  def __int__(self):
    return self.value

  def __bytes__(self):
    return self.value.to_bytes(16, 'big')

  def __array__(self, dtype=None, copy=None):
    return np.frombuffer(bytes(self), np.uint8)

  def __str__(self):
    if self.DEBUG_ID is not None:
      return str(self.value)
    v, out = self.value, []
    while v:
      v, r = divmod(v, 62)
      out.append(self._ALPHABET[r])
    return ''.join(reversed(out)).rjust(22, '0')

  def __repr__(self):
    return f'UUID({self})'

  def __eq__(self, other):
    return isinstance(other, UUID) and self.value == other.value

  def __lt__(self, other):
    return self.value < other.value

  def __hash__(self):
    return self._hash


def timestamp(now=None, millis=False):
  now = now or datetime.datetime.now()
  text = now.strftime('%Y%m%dT%H%M%S')
  if millis:
    text += f'F{now.microsecond:06d}'
  return text


# --------------------------------------------------------------------- Path

class Path(type(pathlib.Path())):
  """pathlib path with the few extras the reference calls."""

  def mkdir(self, mode=0o777, parents=True, exist_ok=True):
    super().mkdir(mode=mode, parents=True, exist_ok=True)

  def write(self, content, mode='w'):
    with open(self, mode) as f:
      f.write(content)

  def read(self, mode='r'):
    with open(self, mode) as f:
      return f.read()

  def glob(self, pattern):
    return [Path(x) for x in super().glob(pattern)]


# -------------------------------------------------------------------- timer

class _Section(contextlib.ContextDecorator):
  def __init__(self, name):
    self.name = name

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    return False


timer = types.SimpleNamespace(
    section=_Section,
    stats=lambda: {'summary': ''},
)


# ------------------------------------------------------------------- RWLock

class RWLock:
  """Many readers or one writer (writer preference not needed here)."""

  def __init__(self):
    self._cond = threading.Condition()
    self._readers = 0
    self._writer = False

  @property
  @contextlib.contextmanager
  def reading(self):
    with self._cond:
      while self._writer:
        self._cond.wait()
      self._readers += 1
    try:
      yield
    finally:
      with self._cond:
        self._readers -= 1
        self._cond.notify_all()

  @property
  @contextlib.contextmanager
  def writing(self):
    with self._cond:
      while self._writer or self._readers:
        self._cond.wait()
      self._writer = True
    try:
      yield
    finally:
      with self._cond:
        self._writer = False
        self._cond.notify_all()


# --------------------------------------------------------------------- tree

def _treemap(fn, *trees, isleaf=None):
  first = trees[0]
  if isleaf and isleaf(first):
    return fn(*trees)
  if isinstance(first, dict):
    return type(first)(
        {k: _treemap(fn, *[t[k] for t in trees], isleaf=isleaf) for k in first})
  if isinstance(first, (list, tuple)):
    out = [_treemap(fn, *xs, isleaf=isleaf) for xs in zip(*trees)]
    return type(first)(out) if not hasattr(first, '_fields') else type(first)(*out)
  return fn(*trees)


tree = types.SimpleNamespace(map=_treemap)


# -------------------------------------------------------------------- Space

class Space:

  def __init__(self, dtype, shape=(), low=None, high=None):
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    self._dtype = np.dtype(dtype)
    assert self._dtype is not object, self._dtype
    self._low = self._bound(low, shape, lower=True)
    self._high = self._bound(high, shape, lower=False)
    self._shape = shape if shape else tuple(self._low.shape)
    self._discrete = (
        np.issubdtype(self._dtype, np.integer) or self._dtype == bool)
    self._random = np.random.RandomState()

  def _bound(self, value, shape, lower):
    if value is not None:
      return np.broadcast_to(np.asarray(value, self._dtype), shape).copy()
    if np.issubdtype(self._dtype, np.floating):
      fill = -np.inf if lower else np.inf
    elif np.issubdtype(self._dtype, np.integer):
      info = np.iinfo(self._dtype)
      fill = info.min if lower else info.max
    elif self._dtype == bool:
      fill = not lower
    else:
      raise ValueError(self._dtype)
    return np.full(shape, fill, self._dtype)

  dtype = property(lambda self: self._dtype)
  shape = property(lambda self: self._shape)
  low = property(lambda self: self._low)
  high = property(lambda self: self._high)
  discrete = property(lambda self: self._discrete)

  @property
  def classes(self):
    assert self.discrete
    classes = self._high - self._low
    if not classes.ndim:
      classes = int(classes.item())
    return classes

  def __repr__(self):
    low = None if self.low is None else self.low.min()
    high = None if self.high is None else self.high.max()
    return (f'Space({self.dtype.name}, shape={self.shape}, '
            f'low={low}, high={high})')

  def __contains__(self, value):
    value = np.asarray(value)
    if np.issubdtype(self.dtype, str):
      return np.issubdtype(value.dtype, str)
    if value.shape != self.shape:
      return False
    if (value > self.high).any():
      return False
    if (value < self.low).any():
      return False
    if value.dtype != self.dtype:
      return False
    return True

  def sample(self):
    low, high = self.low, self.high
    if np.issubdtype(self.dtype, np.floating):
      low = np.maximum(np.ones(self.shape) * np.finfo(self.dtype).min, low)
      high = np.minimum(np.ones(self.shape) * np.finfo(self.dtype).max, high)
      return self._random.uniform(low, high, self.shape).astype(self.dtype)
    if self.dtype == bool:
      return self._random.randint(0, 2, self.shape).astype(bool)
    return self._random.randint(low, high, self.shape).astype(self.dtype)
