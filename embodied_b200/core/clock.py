"""Wall-clock triggers of the run loop: ``should_log = LocalClock(log_every)`` and
friends in ``run.train`` (reference: embodied/core/clock.py:95-118).  The
multi-replica GlobalClock server of that file belongs to run/parallel.py and is
out of scope.

Semantics of ``every`` (seconds): 0 = never fires, negative = fires on every call,
positive = fires when at least ``every`` seconds have passed since it last fired;
the first call only arms the timer and fires iff ``first``.  ``skip=True`` makes a
call a no-op that does not touch the timer.
"""
import time


class LocalClock:

  def __init__(self, every, first=False):
    self.every = every
    self.first = first
    self.prev = None          # time of the last firing (None = not armed yet)

  def __call__(self, step=None, skip=None):
    if skip or self.every == 0:
      return False
    if self.every < 0:
      return True
    now = time.time()
    armed, self.prev = self.prev, (now if self.prev is None else self.prev)
    if armed is None:
      return bool(self.first)
    due = now >= armed + self.every
    if due:
      self.prev = now
    return due
