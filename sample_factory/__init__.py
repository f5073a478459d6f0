"""`sample_factory` import surface over the B200 engine (sample_factory_b200).

The reference's public API -- the module paths `sf_examples/*` and user projects import -- re-exported from the device
engine, so an existing entry script runs unmodified with this repository on PYTHONPATH:

    from sample_factory.cfg.arguments import parse_full_cfg, parse_sf_args      (cfg/arguments.py:24-94)
    from sample_factory.envs.env_utils import register_env                        (envs/env_utils.py:12-31)
    from sample_factory.train import run_rl                                       (train.py:31-41)
    from sample_factory.enjoy import enjoy                                        (enjoy.py:103)
    from sample_factory.algo.utils.context import global_model_factory            (algo/utils/context.py)

Only the hot path lives behind these names: the process tree, PBT, wandb / HF hub tooling of the
reference are not rebuilt.  Custom torch modules registered through the model factory cannot run on hand-written
kernels: registering them is accepted (the API exists) and `run_rl` then fails with an explicit message."""
__version__ = "2.1.3+sfb200"
