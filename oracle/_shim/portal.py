"""Mini stand-in for the third-party ``portal`` package (TEST INFRASTRUCTURE).

Only what ``embodied/core/driver.py:22,43,102-106`` and
``embodied/core/streams.py:41`` touch: ``Process(fn, *args, start=True)`` whose
target receives a ``context`` with ``.running``, ``.kill()``; ``Thread`` with
``.start()``.  ``Client``/``Server`` exist so ``core/clock.py`` imports.
"""
import multiprocessing as mp
import threading


class _Context:
  def __init__(self, event):
    self._event = event

  @property
  def running(self):
    return not self._event.is_set()


def _entry(fn, event, args):
  fn(_Context(event), *args)


class Process:

  def __init__(self, fn, *args, name=None, start=False):
    ctx = mp.get_context()
    self._stop = ctx.Event()
    self._proc = ctx.Process(
        target=_entry, args=(fn, self._stop, args), daemon=True, name=name)
    self.started = False
    start and self.start()

  def start(self):
    self.started = True
    self._proc.start()
    return self

  @property
  def running(self):
    return self._proc.is_alive()

  def join(self, timeout=None):
    self._proc.join(timeout)

  def kill(self, timeout=1):
    self._stop.set()
    self._proc.join(timeout)
    if self._proc.is_alive():
      self._proc.terminate()
      self._proc.join(timeout)


class Thread:

  def __init__(self, fn, *args, name=None, start=False):
    self._thread = threading.Thread(
        target=fn, args=args, daemon=True, name=name)
    self.started = False
    start and self.start()

  def start(self):
    self.started = True
    self._thread.start()
    return self

  @property
  def running(self):
    return self._thread.is_alive()

  def join(self, timeout=None):
    self._thread.join(timeout)

  def kill(self, timeout=1):
    pass


class Client:
  def __init__(self, *a, **k):
    raise NotImplementedError('portal.Client is outside the oracle surface')


class Server(Client):
  pass
