"""Minimal gymnasium.envs (fallback): registry + classic-control CartPole."""
from . import registration  # noqa: F401
from .classic_control import CartPoleEnv  # noqa: F401
from .registration import make, register, registry, spec  # noqa: F401

register(id="CartPole-v1", entry_point=CartPoleEnv, max_episode_steps=500, reward_threshold=475.0)
register(id="CartPole-v0", entry_point=CartPoleEnv, max_episode_steps=200, reward_threshold=195.0)
