"""sample_factory.envs.env_utils (envs/env_utils.py:12-133): env registry + the two optional env interfaces."""
from sample_factory_b200.envs import (  # noqa: F401
    RewardShapingInterface,
    TrainingInfoInterface,
    register_env,
    set_training_info,
)


class EnvCriticalError(Exception):
    pass
