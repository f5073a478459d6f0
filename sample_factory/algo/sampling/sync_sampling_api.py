"""sample_factory.algo.sampling.sync_sampling_api (sync_sampling_api.py:16-65) over the device trajectory store"""
from sample_factory_b200.sampling_api import SyncSamplingAPI  # noqa: F401
