"""embodied_b200: B200-native actor-learner hot path behind the embodied API.

Mirrors the names the reference exports (embodied/__init__.py:9-12,
embodied/core/__init__.py:1-14): Agent, Env, Driver, Replay, RandomAgent,
Wrapper, LocalClock and the submodules clock, limiters, selectors, streams,
wrappers, replay, run, envs, elements.
"""
__version__ = '0.1.0'

from . import elements
from .core import *
from .core import replay
from . import envs
from . import run
