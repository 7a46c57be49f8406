"""Host utilities the run loop needs (ids, spaces, clocks, aggregation, config).

The reference takes these from the third-party ``elements`` package
(``requirements.txt:5``), which is neither vendored under /root/reference nor
installed here; this module provides the subset ``embodied/core`` and
``embodied/run/train.py`` call, under the same names, so user code written
against ``elements.X`` keeps working as ``embodied_b200.elements.X``.
"""
import collections
import concurrent.futures
import contextlib
import datetime
import itertools
import json
import pathlib
import pickle
import re
import string
import threading
import time
import types
import uuid as uuidlib

import numpy as np


# ---------------------------------------------------------------------------
# UUID: 128-bit chunk ids.  bytes(UUID) is 16 bytes big-endian, which fixes the
# stepid layout uuid16||be32(index) (reference embodied/core/replay.py:90-91).

class UUID:

  __slots__ = ('value', '_hash')
  DEBUG_ID = None
  _LOCK = threading.Lock()
  _DIGITS = string.digits + string.ascii_letters
  _VALUE = {c: i for i, c in enumerate(_DIGITS)}

  @classmethod
  def reset(cls, *, debug):
    """debug=True: ids 1, 2, 3, ... (reference tests/test_replay.py:157)."""
    cls.DEBUG_ID = itertools.count(1) if debug else None

  def __init__(self, value=None):
    cls = type(self)
    if value is None:
      if cls.DEBUG_ID is None:
        value = uuidlib.uuid4().int
      else:
        with cls._LOCK:
          value = next(cls.DEBUG_ID)
    elif isinstance(value, UUID):
      value = value.value
    elif isinstance(value, (int, np.integer)):
      value = int(value)
    elif isinstance(value, str):
      if cls.DEBUG_ID is not None and value.isdigit():
        value = int(value)
      else:
        number = 0
        for char in value:
          number = number * 62 + cls._VALUE[char]
        value = number
    elif isinstance(value, np.ndarray):
      value = int.from_bytes(value.tobytes(), 'big')
    elif isinstance(value, (bytes, bytearray, memoryview)):
      value = int.from_bytes(bytes(value), 'big')
    else:
      raise TypeError(f'cannot make a UUID from {type(value)}')
    if not 0 <= value < 1 << 128:
      raise ValueError(value)
    self.value = value
    self._hash = hash(value)

  def __int__(self):
    return self.value

  def __index__(self):
    return self.value

  def __bytes__(self):
    return self.value.to_bytes(16, 'big')

  def __array__(self, dtype=None, copy=None):
    return np.frombuffer(bytes(self), np.uint8)

  def __str__(self):
    if type(self).DEBUG_ID is not None:
      return str(self.value)
    number, chars = self.value, []
    while number:
      number, rest = divmod(number, 62)
      chars.append(self._DIGITS[rest])
    return ''.join(reversed(chars)).rjust(22, '0')

  def __repr__(self):
    return f'UUID({self})'

  def __eq__(self, other):
    return isinstance(other, UUID) and other.value == self.value

  def __lt__(self, other):
    return self.value < other.value

  def __hash__(self):
    return self._hash


def timestamp(now=None, millis=False):
  now = now or datetime.datetime.now()
  stamp = now.strftime('%Y%m%dT%H%M%S')
  return stamp + f'F{now.microsecond:06d}' if millis else stamp


# ---------------------------------------------------------------------------

class Path(type(pathlib.Path())):

  def mkdir(self, mode=0o777, parents=True, exist_ok=True):
    super().mkdir(mode=mode, parents=True, exist_ok=True)

  def write(self, content, mode='w'):
    with open(self, mode) as f:
      f.write(content)

  def read(self, mode='r'):
    with open(self, mode) as f:
      return f.read()

  def glob(self, pattern):
    return [Path(p) for p in super().glob(pattern)]


class RWLock:
  """Shared/exclusive lock (reference embodied/core/replay.py:37,79,297,347)."""

  def __init__(self):
    self._cv = threading.Condition()
    self._readers = 0
    self._writing = False

  @property
  @contextlib.contextmanager
  def reading(self):
    with self._cv:
      self._cv.wait_for(lambda: not self._writing)
      self._readers += 1
    try:
      yield
    finally:
      with self._cv:
        self._readers -= 1
        self._cv.notify_all()

  @property
  @contextlib.contextmanager
  def writing(self):
    with self._cv:
      self._cv.wait_for(lambda: not self._writing and not self._readers)
      self._writing = True
    try:
      yield
    finally:
      with self._cv:
        self._writing = False
        self._cv.notify_all()


# ---------------------------------------------------------------------------
# timer: wall-clock sections with the reference's section names
# (replay_add, replay_sample, ...); summary goes to the logger.

class _Timer:

  def __init__(self):
    self._stats = collections.defaultdict(lambda: [0, 0.0])
    self.enabled = True

  def section(self, name):
    return _Section(self, name)

  def stats(self, reset=True):
    rows = sorted(self._stats.items(), key=lambda kv: -kv[1][1])
    summary = '\n'.join(
        f'{n:<24} calls={c:<9d} total={t:8.3f}s' for n, (c, t) in rows)
    out = {'summary': summary}
    for name, (calls, total) in rows:
      out[f'{name}/count'] = calls
      out[f'{name}/sum'] = total
    if reset:
      self._stats.clear()
    return out


class _Section(contextlib.ContextDecorator):

  def __init__(self, timer, name):
    self.timer, self.name = timer, name

  def __enter__(self):
    self.start = time.perf_counter()
    return self

  def __exit__(self, *exc):
    if self.timer.enabled:
      entry = self.timer._stats[self.name]
      entry[0] += 1
      entry[1] += time.perf_counter() - self.start
    return False


timer = _Timer()


def _tree_map(fn, *trees, isleaf=None):
  head = trees[0]
  if isleaf is not None and isleaf(head):
    return fn(*trees)
  if isinstance(head, dict):
    return {k: _tree_map(fn, *(t[k] for t in trees), isleaf=isleaf)
            for k in head}
  if isinstance(head, (list, tuple)):
    mapped = [_tree_map(fn, *xs, isleaf=isleaf) for xs in zip(*trees)]
    if hasattr(head, '_fields'):
      return type(head)(*mapped)
    return type(head)(mapped)
  return fn(*trees)


tree = types.SimpleNamespace(map=_tree_map)


# ---------------------------------------------------------------------------

class Space:
  """dtype/shape/bounds of one observation or action key
  (reference usage: embodied/envs/dummy.py:16-36, core/wrappers.py:103,240)."""

  def __init__(self, dtype, shape=(), low=None, high=None):
    if isinstance(shape, (int, np.integer)):
      shape = (int(shape),)
    self._dtype = np.dtype(dtype)
    shape = tuple(int(x) for x in shape)
    self._low = self._limit(low, shape, np.less)
    self._high = self._limit(high, shape, np.greater)
    self._shape = shape or tuple(self._low.shape)
    self._discrete = bool(
        np.issubdtype(self._dtype, np.integer) or self._dtype == bool)
    self._rng = np.random.RandomState()

  def _limit(self, given, shape, side):
    if given is None:
      if np.issubdtype(self._dtype, np.floating):
        given = -np.inf if side is np.less else np.inf
      elif np.issubdtype(self._dtype, np.integer):
        info = np.iinfo(self._dtype)
        given = info.min if side is np.less else info.max
      elif self._dtype == bool:
        given = side is np.greater
      else:
        raise TypeError(self._dtype)
    return np.array(np.broadcast_to(np.asarray(given, self._dtype), shape))

  dtype = property(lambda self: self._dtype)
  shape = property(lambda self: self._shape)
  low = property(lambda self: self._low)
  high = property(lambda self: self._high)
  discrete = property(lambda self: self._discrete)

  @property
  def classes(self):
    assert self._discrete
    span = self._high - self._low
    return int(span.item()) if not span.ndim else span

  def __repr__(self):
    return (f'Space({self._dtype.name}, shape={self._shape}, '
            f'low={self._low.min()}, high={self._high.max()})')

  def __contains__(self, value):
    value = np.asarray(value)
    if value.shape != self._shape or value.dtype != self._dtype:
      return False
    return not ((value > self._high).any() or (value < self._low).any())

  def sample(self):
    low, high = self._low, self._high
    if np.issubdtype(self._dtype, np.floating):
      info = np.finfo(self._dtype)
      low, high = np.maximum(low, info.min), np.minimum(high, info.max)
      return self._rng.uniform(low, high, self._shape).astype(self._dtype)
    if self._dtype == bool:
      return self._rng.randint(0, 2, self._shape).astype(bool)
    return self._rng.randint(low, high, self._shape).astype(self._dtype)


# ---------------------------------------------------------------------------
# when.*: pacing helpers of the run loop.

class _Ratio:
  """How many train steps are due at env step `step`
  (embodied/run/train.py:26,73; pacing pinned by tests/test_train.py:20-23:
  train calls * B * T ~= env steps * train_ratio)."""

  def __init__(self, ratio):
    assert ratio >= 0, ratio
    self._ratio = ratio
    self._prev = None

  def __call__(self, step):
    step = int(step)
    if self._ratio == 0:
      return 0
    if self._prev is None:
      self._prev = step
      return 1
    repeats = int((step - self._prev) * self._ratio)
    self._prev += repeats / self._ratio
    return repeats


class _Every:

  def __init__(self, every, initial=True):
    self._every, self._initial, self._prev = every, initial, None

  def __call__(self, step):
    step = int(step)
    if self._every == 0:
      return False
    if self._every < 0:
      return True
    if self._prev is None:
      self._prev = (step // self._every) * self._every
      return self._initial
    if step >= self._prev + self._every:
      self._prev += self._every
      return True
    return False


class _Clock:

  def __init__(self, every, first=True):
    self._every, self._first, self._prev = every, first, None

  def __call__(self, step=None):
    if self._every == 0:
      return False
    if self._every < 0:
      return True
    now = time.time()
    if self._prev is None:
      self._prev = now
      return self._first
    if now >= self._prev + self._every:
      self._prev = now
      return True
    return False


when = types.SimpleNamespace(Ratio=_Ratio, Every=_Every, Clock=_Clock)


class Counter:

  def __init__(self, initial=0):
    self.value = initial
    self._lock = threading.Lock()

  def __int__(self):
    return int(self.value)

  __index__ = __int__

  def __float__(self):
    return float(self.value)

  def __repr__(self):
    return f'Counter({self.value})'

  def __eq__(self, other):
    return int(self) == other

  def __lt__(self, other):
    return int(self) < other

  def __le__(self, other):
    return int(self) <= other

  def __gt__(self, other):
    return int(self) > other

  def __ge__(self, other):
    return int(self) >= other

  def __add__(self, other):
    return int(self) + other

  def __hash__(self):
    return id(self)

  def increment(self, amount=1):
    with self._lock:
      self.value += amount

  def save(self):
    return self.value

  def load(self, value):
    self.value = value


class FPS:

  def __init__(self):
    self._start = time.time()
    self._total = 0

  def step(self, amount=1):
    self._total += amount

  def result(self, reset=True):
    now = time.time()
    fps = self._total / max(now - self._start, 1e-9)
    if reset:
      self._start, self._total = now, 0
    return fps


class Agg:
  """Metric aggregation between log writes (embodied/run/train.py:19-21)."""

  def __init__(self, maxlen=1e6):
    self._maxlen = int(maxlen)
    self.reset()

  def reset(self):
    self._sum, self._count, self._max, self._stack, self._how = {}, {}, {}, {}, {}

  def add(self, key_or_dict, value=None, agg='default', prefix=None):
    if value is not None:
      self._add(key_or_dict, value, agg, prefix)
      return
    for key, val in key_or_dict.items():
      how = agg[key] if isinstance(agg, dict) else agg
      self._add(key, val, how, prefix)

  def _add(self, key, value, agg, prefix):
    key = f'{prefix}/{key}' if prefix else key
    value = np.asarray(value)
    if agg == 'default':
      agg = 'avg' if value.ndim == 0 else 'last'
    agg = (agg,) if isinstance(agg, str) else tuple(agg)
    self._how[key] = agg
    for how in agg:
      if how in ('avg', 'sum'):
        self._sum[key] = self._sum.get(key, 0) + value
        self._count[key] = self._count.get(key, 0) + 1
      elif how == 'max':
        self._max[key] = np.maximum(self._max.get(key, value), value)
      elif how == 'stack':
        stack = self._stack.setdefault(key, [])
        if len(stack) < self._maxlen:
          stack.append(value)
      elif how == 'last':
        self._stack[key] = [value]
      else:
        raise ValueError(how)

  def result(self, reset=True, prefix=None):
    out = {}
    for key, hows in self._how.items():
      for how in hows:
        name = key if len(hows) == 1 else f'{key}/{how}'
        if how == 'avg':
          out[name] = self._sum[key] / self._count[key]
        elif how == 'sum':
          out[name] = self._sum[key]
        elif how == 'max':
          out[name] = self._max[key]
        elif how == 'stack':
          out[name] = np.stack(self._stack[key])
        elif how == 'last':
          out[name] = self._stack[key][-1]
    if prefix:
      out = {f'{prefix}/{k}': v for k, v in out.items()}
    reset and self.reset()
    return out


class Usage:

  def __init__(self, psutil=True, **unused):
    self._proc = None
    if psutil:
      try:
        import psutil as ps
        self._ps, self._proc = ps, ps.Process()
      except ImportError:
        pass

  def stats(self):
    if not self._proc:
      return {}
    mem = self._ps.virtual_memory()
    return {
        'psutil/proc_cpu_usage': self._proc.cpu_percent() / 100,
        'psutil/proc_ram_gb': self._proc.memory_info().rss / 1024 ** 3,
        'psutil/total_ram_frac': mem.percent / 100,
    }


# ---------------------------------------------------------------------------

class Config(dict):
  """Nested, attribute-accessible, immutable config with regex-key update()
  (reference: elements.Config as used in dreamerv3/main.py:23-31,55-66)."""

  SEP = '.'

  def __init__(self, *args, **kwargs):
    mapping = dict(*args, **kwargs)
    nested = self._nest(self._flatten(mapping))
    super().__init__({k: self._wrap(v) for k, v in nested.items()})

  @classmethod
  def _wrap(cls, value):
    if isinstance(value, dict) and not isinstance(value, Config):
      return Config(value)
    if isinstance(value, list):
      return tuple(value)
    return value

  @classmethod
  def _flatten(cls, mapping, prefix=''):
    flat = {}
    for key, value in mapping.items():
      name = f'{prefix}{cls.SEP}{key}' if prefix else str(key)
      if isinstance(value, dict) and value:
        flat.update(cls._flatten(value, name))
      else:
        flat[name] = value
    return flat

  @classmethod
  def _nest(cls, flat):
    nested = {}
    for key, value in flat.items():
      node = nested
      *parents, leaf = key.split(cls.SEP)
      for part in parents:
        child = node.get(part)
        if type(child) is not dict:          # absent, or an empty-mapping leaf (`dummy: {}`)
          child = node[part] = dict(child) if isinstance(child, dict) else {}
        node = child
      if isinstance(value, dict) and not value and isinstance(node.get(leaf), dict):
        continue                             # an empty mapping never erases a populated one
      node[leaf] = value
    return nested

  @property
  def flat(self):
    return self._flatten(self)

  def __getattr__(self, name):
    if name.startswith('_'):
      raise AttributeError(name)
    try:
      return self[name]
    except KeyError:
      raise AttributeError(name)

  def __setattr__(self, name, value):
    raise AttributeError('Config is immutable; use update()')

  __setitem__ = __setattr__

  def __getitem__(self, name):
    if isinstance(name, str) and self.SEP in name and name not in self.keys():
      node = self
      for part in name.split(self.SEP):
        node = dict.__getitem__(node, part)
      return node
    return dict.__getitem__(self, name)

  def __reduce__(self):
    return (Config, (self._plain(),))

  def _plain(self):
    return {k: v._plain() if isinstance(v, Config) else v
            for k, v in self.items()}

  def update(self, *args, **kwargs):
    flat = self.flat
    for key, new in self._flatten(dict(*args, **kwargs)).items():
      if key in flat:
        targets = [key]
      else:
        pattern = re.compile(key)
        targets = [k for k in flat if pattern.fullmatch(k)]
        if not targets:
          targets = [key]
      for target in targets:
        old = flat.get(target)
        if old is not None and new is not None and not isinstance(
            new, type(old)) and not isinstance(old, (tuple, dict)):
          try:
            new_cast = type(old)(new)
          except (TypeError, ValueError):
            raise TypeError(
                f"cannot set '{target}' of type {type(old).__name__} "
                f"to {new!r}")
          flat[target] = new_cast
        else:
          flat[target] = new
    return Config(self._nest(flat))


class Checkpoint:
  """Attribute registry with save()/load() duck typing
  (embodied/run/train.py:82-89,115-116)."""

  def __init__(self, filename=None):
    object.__setattr__(self, '_filename', filename and Path(filename))
    object.__setattr__(self, '_values', {})

  def __setattr__(self, name, value):
    if name in ('exists', 'save', 'load', 'load_or_save'):
      raise AttributeError(name)
    assert hasattr(value, 'save') and hasattr(value, 'load'), name
    self._values[name] = value

  def __getattr__(self, name):
    if name.startswith('_'):
      raise AttributeError(name)
    try:
      return self._values[name]
    except KeyError:
      raise AttributeError(name)

  def exists(self, filename=None):
    filename = Path(filename or self._filename)
    return filename.exists()

  def save(self, filename=None, keys=None):
    filename = Path(filename or self._filename)
    keys = tuple(self._values.keys() if keys is None else keys)
    data = {k: self._values[k].save() for k in keys}
    data['_timestamp'] = time.time()
    filename.parent.mkdir()
    tmp = filename.parent / (filename.name + '.tmp')
    tmp.write(pickle.dumps(data), mode='wb')
    tmp.replace(filename)

  def load(self, filename=None, keys=None):
    filename = Path(filename or self._filename)
    data = pickle.loads(filename.read(mode='rb'))
    keys = tuple(data.keys() if keys is None else keys)
    for key in keys:
      if key.startswith('_'):
        continue
      self._values[key].load(data[key])

  def load_or_save(self):
    if self.exists():
      self.load()
    else:
      self.save()


# ---------------------------------------------------------------------------

class TerminalOutput:

  def __init__(self, pattern=r'.*', name=None):
    self._pattern = re.compile(pattern)
    self._name = name

  def __call__(self, summaries):
    step = max((s for s, _, _ in summaries), default=0)
    scalars = {k: float(v) for _, k, v in summaries
               if np.asarray(v).ndim == 0 and self._pattern.search(k)}
    head = f'[{self._name or "step"} {step}]'
    body = ' / '.join(f'{k} {v:.4g}' for k, v in list(scalars.items())[:24])
    print(head, body, flush=True)


class JSONLOutput:

  def __init__(self, logdir, filename='metrics.jsonl', pattern=r'.*'):
    self._path = Path(logdir) / filename
    self._path.parent.mkdir()
    self._pattern = re.compile(pattern)

  def __call__(self, summaries):
    bystep = collections.defaultdict(dict)
    for step, name, value in summaries:
      if np.asarray(value).ndim == 0 and self._pattern.search(name):
        bystep[step][name] = float(value)
    lines = ''.join(
        json.dumps({'step': step, **vals}) + '\n'
        for step, vals in bystep.items())
    with open(self._path, 'a') as f:
      f.write(lines)


class Logger:

  def __init__(self, step, outputs, multiplier=1):
    self.step = step
    self.outputs = outputs
    self.multiplier = multiplier
    self._metrics = []

  def add(self, mapping, prefix=None):
    step = int(self.step) * self.multiplier
    for name, value in dict(mapping).items():
      name = f'{prefix}/{name}' if prefix else name
      if isinstance(value, str):
        continue
      self._metrics.append((step, name, np.asarray(value)))

  def scalar(self, name, value):
    self.add({name: value})

  def write(self):
    if self._metrics:
      for output in self.outputs:
        output(tuple(self._metrics))
      self._metrics.clear()

  def close(self):
    self.write()


logger = types.SimpleNamespace(
    TerminalOutput=TerminalOutput, JSONLOutput=JSONLOutput)
