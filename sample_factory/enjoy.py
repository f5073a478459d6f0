"""sample_factory.enjoy (enjoy.py:92-190): evaluation of a saved policy over the device sampler."""
from sample_factory_b200.enjoy import enjoy  # noqa: F401
