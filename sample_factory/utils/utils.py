"""sample_factory.utils.utils (utils/utils.py): the helpers example scripts import."""
from __future__ import annotations

import importlib.util
import logging
import os

from sample_factory_b200.cfg import str2bool  # noqa: F401
from sample_factory_b200.train import experiment_dir  # noqa: F401

log = logging.getLogger("sample_factory")
if not log.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("[%(asctime)s] %(message)s"))
    log.addHandler(_h)
    log.setLevel(logging.INFO)


def is_module_available(module_name: str) -> bool:
    try:
        return importlib.util.find_spec(module_name) is not None
    except (ImportError, ValueError):
        return False


def project_root() -> str:
    return os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ensure_dir_exists(path: str) -> str:
    os.makedirs(path, exist_ok=True)
    return path


def debug_log_every_n(n, msg, *args, **kwargs):
    log.debug(msg, *args, **kwargs)


def safe_ensure_dir_exists(path: str) -> str:
    os.makedirs(path, exist_ok=True)
    return path


def project_tmp_dir(mkdir: bool = True) -> str:
    import tempfile

    d = os.path.join(tempfile.gettempdir(), f"sample_factory_{os.environ.get('USER', 'user')}")
    return ensure_dir_exists(d) if mkdir else d


def static_vars(**kwargs):
    """decorator: attach attributes to a function (utils/utils.py)"""
    def decorate(func):
        for k, v in kwargs.items():
            setattr(func, k, v)
        return func

    return decorate
