"""embodied_b200.core: the names ``embodied.core`` exports (embodied/core/__init__.py:1-14)
-- the three protocols, Driver, Replay, RandomAgent, Wrapper, LocalClock -- and the
submodules the run loop and the factories reach into."""
from . import base, clock, driver, limiters, random, replay, selectors, streams, wrappers

Agent, Env, Stream = base.Agent, base.Env, base.Stream
Driver = driver.Driver
Replay = replay.Replay
RandomAgent = random.RandomAgent
Wrapper = wrappers.Wrapper
LocalClock = clock.LocalClock

__all__ = [
    'Agent', 'Env', 'Stream', 'Driver', 'Replay', 'RandomAgent', 'Wrapper', 'LocalClock',
    'base', 'clock', 'driver', 'limiters', 'random', 'replay', 'selectors', 'streams', 'wrappers']
