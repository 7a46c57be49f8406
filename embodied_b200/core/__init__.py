from .base import Agent, Env, Stream
from .clock import LocalClock
from .driver import Driver
from .random import RandomAgent
from .replay import Replay
from .wrappers import Wrapper

from . import clock
from . import limiters
from . import selectors
from . import streams
from . import wrappers
