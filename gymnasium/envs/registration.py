"""Minimal env registry (fallback): register(id, entry_point, max_episode_steps) / make(id, **kwargs)."""
from __future__ import annotations

import importlib
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional, Union


@dataclass
class EnvSpec:
    id: str
    entry_point: Union[str, Callable, None] = None
    max_episode_steps: Optional[int] = None
    reward_threshold: Optional[float] = None
    kwargs: Dict[str, Any] = field(default_factory=dict)


registry: Dict[str, EnvSpec] = {}


def register(id: str, entry_point=None, max_episode_steps=None, reward_threshold=None, kwargs=None, **_ignored) -> None:
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, reward_threshold, dict(kwargs or {}))


def spec(id: str) -> EnvSpec:
    if id not in registry:
        raise KeyError(f"No registered env with id: {id} (offline gymnasium fallback knows: {sorted(registry)})")
    return registry[id]


def make(id, max_episode_steps=None, render_mode=None, **kwargs):
    from ..wrappers import TimeLimit

    s = spec(id) if isinstance(id, str) else id
    ep = s.entry_point
    if isinstance(ep, str):
        mod, _, attr = ep.partition(":")
        ep = getattr(importlib.import_module(mod), attr)
    kw = dict(s.kwargs, **kwargs)
    try:
        env = ep(render_mode=render_mode, **kw)
    except TypeError:
        env = ep(**kw)
    env.spec = s
    steps = max_episode_steps if max_episode_steps is not None else s.max_episode_steps
    if steps is not None and steps > 0:
        env = TimeLimit(env, steps)
    return env
