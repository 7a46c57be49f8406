"""``wait``: the polling wait ``Replay.sample`` blocks in until the buffer holds a
full window (reference: embodied/core/limiters.py:5-16).  Returns the seconds
spent waiting (0 if the predicate already held) and prints a progress line every
``notify`` seconds so that a starved sampler is visible in the log.  The
``SamplesPerInsert`` rate limiter of that file is used by run/parallel.py only
(out of scope)."""
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
