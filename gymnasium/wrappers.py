"""Minimal gymnasium.wrappers (fallback): TimeLimit and RecordEpisodeStatistics."""
from __future__ import annotations

import time

from .core import Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps: int):
        super().__init__(env)
        self._max_episode_steps = int(max_episode_steps)
        self._elapsed_steps = 0

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info


class RecordEpisodeStatistics(Wrapper):
    def __init__(self, env, *args, **kwargs):
        super().__init__(env)
        self._ret, self._len, self._t0 = 0.0, 0, time.perf_counter()

    def reset(self, **kwargs):
        self._ret, self._len, self._t0 = 0.0, 0, time.perf_counter()
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._ret += float(reward)
        self._len += 1
        if terminated or truncated:
            info = dict(info, episode=dict(r=self._ret, l=self._len, t=time.perf_counter() - self._t0))
        return obs, reward, terminated, truncated, info
