"""``wait``: the polling wait ``Replay.sample`` blocks in until the buffer holds a
full window (reference: embodied/core/limiters.py:5-16).  Returns the seconds
spent waiting (0 if the predicate already held) and prints a progress line every
``notify`` seconds so that a starved sampler is visible in the log.

``SamplesPerInsert`` (limiters.py:19-80) couples the learner's sample rate to the actors'
insert rate: a balance that every insert raises by ``samples_per_insert`` (once ``minsize``
items exist) and every sample lowers by one; inserting is refused while the balance is at
``tolerance * samples_per_insert`` or above, sampling while it is at ``-tolerance`` or below.
A rate of 0 or less disables both limits (the buffer only has to hold ``minsize`` items)."""
import threading
import time


def wait(predicate, message, info=None, sleep=0.01, notify=60):
  if predicate():
    return 0
  began = time.time()
  reported = began
  while not predicate():
    time.sleep(sleep)
    now = time.time()
    if now - reported > notify:
      print(f'{message} {now - began:.1f}s: {info}')
      reported = now
  return time.time() - began


class SamplesPerInsert:

  def __init__(self, samples_per_insert, tolerance, minsize):
    assert 1 <= minsize
    self.samples_per_insert = samples_per_insert
    self.minsize = minsize
    self.ceiling = tolerance * samples_per_insert      # want_insert below this balance
    self.floor = -tolerance                            # want_sample above this balance
    self.avail = -minsize
    self.size = 0
    self.lock = threading.Lock()

  # names of the reference's attributes (limiters.py:26-28)
  max_avail = property(lambda self: self.ceiling)
  min_avail = property(lambda self: self.floor)

  def save(self):
    return {'size': self.size, 'avail': self.avail}

  def load(self, data):
    self.size, self.avail = data['size'], data['avail']

  def want_insert(self):
    unlimited = self.size < self.minsize or self.samples_per_insert <= 0
    return unlimited or self.avail < self.ceiling

  def want_sample(self):
    if self.size < self.minsize:
      return False
    return self.samples_per_insert <= 0 or self.floor < self.avail

  def insert(self):
    with self.lock:
      self.size += 1
      if self.size >= self.minsize:
        self.avail += self.samples_per_insert

  def sample(self):
    with self.lock:
      self.avail -= 1
