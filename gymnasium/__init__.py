"""Offline FALLBACK for the `gymnasium` package (this image has no gymnasium wheel and no network).

If a real gymnasium is installed anywhere else on sys.path it is loaded INSTEAD of this directory and takes this module's
place in sys.modules -- the fallback never shadows the real thing.  Otherwise a minimal, dependency-free subset is
provided: `spaces` (Box, Discrete, MultiDiscrete, Tuple, Dict), `Env`, `Wrapper` (+ Observation / Reward / Action
wrappers), `wrappers.TimeLimit`, `register` / `make`, and the classic-control `CartPole-v1` -- what
`sf_examples/train_gym_env.py` (BASELINE.json config 1) and gym-style user envs need.  Written from the published
gymnasium API; no gymnasium source is included."""
import importlib.machinery as _machinery
import importlib.util as _util
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_parent = _os.path.dirname(_here)
_real = None
try:
    _paths = [p for p in _sys.path if _os.path.abspath(p or _os.getcwd()) != _parent]
    _spec = _machinery.PathFinder.find_spec("gymnasium", _paths)
    if _spec is not None and _spec.origin and _os.path.dirname(_os.path.abspath(_spec.origin)) != _here:
        _real = _util.module_from_spec(_spec)
        _sys.modules["gymnasium"] = _real
        _spec.loader.exec_module(_real)
except Exception:   # a broken installation must not take the fallback down with it
    _sys.modules["gymnasium"] = _sys.modules[__name__] if __name__ in _sys.modules else None
    _real = None

if _real is None:
    _sys.modules["gymnasium"] = _sys.modules[__name__]
    from . import spaces  # noqa: E402,F401
    from .core import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: E402,F401
    from .spaces import Space  # noqa: E402,F401
    from . import core, wrappers, envs  # noqa: E402,F401
    from .envs.registration import make, register, registry, spec  # noqa: E402,F401

    __version__ = "0.0.0+sfb200.fallback"
    IS_SFB200_FALLBACK = True
