"""Blocking wait used by Replay.sample (embodied/core/limiters.py:5-16)."""
import time


def wait(predicate, message, info=None, sleep=0.01, notify=60):
  if predicate():
    return 0
  start = last = time.time()
  while not predicate():
    now = time.time()
    if now - last > notify:
      print(f'{message} {now - start:.1f}s: {info}')
      last = now
    time.sleep(sleep)
  return time.time() - start
