"""Batch streams between Replay and Agent.train
(embodied/core/streams.py:12-140; wiring dreamerv3/main.py:261-272):
``Stateless(replay.sample, B, mode)`` -> ``Consec(length, consec, prefix)``.

Batches are dicts of device tensors.  With ``consec == 1`` (every shipped
config) ``Consec`` costs nothing: the window IS the sub-batch, and the int32
``consec`` key is written by the gather launch itself, so the reference's
second full copy (``np.ascontiguousarray``, streams.py:138) disappears.
"""
import functools
import queue
import threading

import numpy as np

from . import base


class Stateless(base.Stream):

  def __init__(self, nextfn, *args, **kwargs):
    if not callable(nextfn) and hasattr(nextfn, '__next__'):
      nextfn = nextfn.__next__
    self.fn = nextfn
    self.args, self.kwargs = args, kwargs
    self.nextfn = functools.partial(nextfn, *args, **kwargs)

  def __iter__(self):
    return self

  def __next__(self):
    return self.nextfn()

  def save(self):
    return None

  def load(self, data):
    pass


def _full(shape, value, like):
  if isinstance(like, np.ndarray):
    return np.full(shape, value, np.int32)
  import torch
  return torch.full(tuple(shape), value, dtype=torch.int32, device=like.device)


def _contiguous(x):
  if isinstance(x, np.ndarray):
    return np.ascontiguousarray(x)
  return x.contiguous()


class Consec(base.Stream):
  """Cuts a window of consec*length+prefix steps into `consec` overlapping
  sub-batches of length+prefix steps and tags them with their index.

      length=3 consec=3 prefix=2
      source:   0 1 2 3 4 5 6 7 8 9 10
      chunk 1:  p-p-#-#-#
      chunk 2:        p-p-#-#-#
      chunk 3:              p-p-#-#-#
  """

  def __init__(
      self, source, length, consec, prefix=0, strict=True, contiguous=False):
    self.source = source
    self.length = length
    self.consec = consec
    self.prefix = prefix
    self.strict = strict
    self.contiguous = contiguous
    self.index = 0
    self.current = None
    self.it = None
    # consec == 1 over Replay.sample: have the gather launch emit 'consec'.
    self._fused = False
    owner = getattr(getattr(source, 'fn', None), '__self__', None)
    if consec == 1 and owner is not None and hasattr(owner, 'store') and \
        'consec' not in getattr(source, 'kwargs', {}):
      try:
        source.nextfn = functools.partial(
            source.fn, *source.args, **source.kwargs, consec=0)
        self._fused = True
      except TypeError:
        pass

  def __iter__(self):
    self.it = iter(self.source)
    return self

  def __next__(self):
    if self.index >= self.consec:
      self.index = 0
    if self.index == 0:
      self.current = next(self.it)
      available = self.current['is_first'].shape[1]
      need = self.length * self.consec + self.prefix
      assert need <= available, (
          self.length, self.consec, self.prefix, available)
      if self.strict:
        assert need == available, (
            self.consec, self.length, self.prefix, available)
    start = self.index * self.length
    stop = start + self.length + self.prefix
    if self._fused and 'consec' in self.current:
      width = self.current['is_first'].shape[1]
      if start == 0 and stop == width:
        self.index += 1
        return dict(self.current)
    chunk = {k: v[:, start: stop] for k, v in self.current.items()
             if k != 'consec'}
    chunk['consec'] = _full(
        chunk['is_first'].shape, self.index, chunk['is_first'])
    if self.contiguous:
      chunk = {k: _contiguous(v) for k, v in chunk.items()}
    self.index += 1
    return chunk

  def save(self):
    return {'source': self.source.save(), 'index': self.index}

  def load(self, data):
    self.source.load(data['source'])
    self.index = data['index']


class Prefetch(base.Stream):
  """Runs source + transform on a thread, `amount` batches ahead
  (streams.py:32-86).  With device batches the GPU work is asynchronous
  anyway; this only hides the host-side index draw."""

  def __init__(self, source, transform=None, amount=1):
    self.source = iter(source) if hasattr(source, '__iter__') else source()
    self.transform = transform or (lambda x: x)
    self.state = self._getstate()
    self.requests = threading.Semaphore(amount)
    self.amount = amount
    self.queue = queue.Queue()
    self.worker = threading.Thread(target=self._worker, daemon=True)
    self.started = False

  def __iter__(self):
    assert not self.started
    self.worker.start()
    self.started = True
    return self

  def __next__(self):
    assert self.started
    result = self.queue.get()
    self.requests.release()
    if isinstance(result, str):
      raise RuntimeError(result)
    data, self.state = result
    return data

  def save(self):
    return self.state

  def load(self, state):
    if self.started:
      for _ in range(self.amount):
        self.queue.get()
    self.source.load(state)
    if self.started:
      self.requests.release(self.amount)

  def _worker(self):
    try:
      while True:
        self.requests.acquire()
        data = self.transform(next(self.source))
        self.queue.put((data, self._getstate()))
    except Exception as e:
      self.queue.put(str(e))
      raise

  def _getstate(self):
    return self.source.save() if hasattr(self.source, 'save') else None
