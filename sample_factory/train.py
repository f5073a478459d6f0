"""sample_factory.train (train.py:12-41): make_runner / run_rl on the device engine."""
from sample_factory_b200.train import Runner, StatusCode, make_runner, run_rl  # noqa: F401
