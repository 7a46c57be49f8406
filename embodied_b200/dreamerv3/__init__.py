"""B200-native DreamerV3 learner behind the embodied Agent protocol."""
from . import config
from .agent import Agent
