"""sample_factory.eval (eval.py): do_eval / generate_trajectories over the device sampler"""
from sample_factory_b200.sampling_api import do_eval, generate_trajectories  # noqa: F401
