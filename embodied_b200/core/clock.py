"""Wall-clock trigger of the run loop (embodied/core/clock.py:95-118).  The
multi-replica GlobalClock server belongs to run/parallel.py (out of scope)."""
import time


class LocalClock:

  def __init__(self, every, first=False):
    self.every = every
    self.prev = None
    self.first = first

  def __call__(self, step=None, skip=None):
    if skip:
      return False
    if self.every == 0:    # zero means off
      return False
    if self.every < 0:     # negative means always
      return True
    now = time.time()
    if self.prev is None:
      self.prev = now
      return self.first
    if now >= self.prev + self.every:
      self.prev = now
      return True
    return False
