"""sample_factory.algo.sampling.evaluation_sampling_api (evaluation_sampling_api.py:31-315)"""
from sample_factory_b200.sampling_api import EvalSamplingAPI  # noqa: F401
