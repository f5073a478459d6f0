"""sample_factory.algo.utils.context (algo/utils/context.py): the global env registry and model factory."""
from typing import Dict

from sample_factory_b200 import envs as _envs
from sample_factory_b200 import model_factory as _mf
from sample_factory_b200.model_factory import ModelFactory


class SampleFactoryContext:
    def __init__(self):
        self.env_registry = _envs.global_env_registry()     # ONE registry / factory, shared with sample_factory_b200
        self.model_factory = _mf.global_model_factory()


GLOBAL_CONTEXT = None


def sf_global_context() -> SampleFactoryContext:
    global GLOBAL_CONTEXT
    if GLOBAL_CONTEXT is None:
        GLOBAL_CONTEXT = SampleFactoryContext()
    return GLOBAL_CONTEXT


def set_global_context(ctx: SampleFactoryContext):
    global GLOBAL_CONTEXT
    GLOBAL_CONTEXT = ctx


def reset_global_context():
    global GLOBAL_CONTEXT
    _envs.global_env_registry().clear()
    _mf.reset_global_model_factory()
    GLOBAL_CONTEXT = SampleFactoryContext()


def global_env_registry() -> Dict:
    return sf_global_context().env_registry


def global_model_factory() -> ModelFactory:
    return sf_global_context().model_factory
