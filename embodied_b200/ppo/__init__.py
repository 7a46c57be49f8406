"""ppo agent (reference: ppo/agent.py, ppo/nets.py) on the embodied_b200 runtime."""
from .agent import Agent
