"""sample_factory.pbt.population_based_training (pbt/population_based_training.py:107-415) on the device engine: the PBT
rules live in sample_factory_b200.pbt, the population runner in sample_factory_b200.multi_policy."""
from sample_factory_b200.multi_policy import MultiPolicyRunner  # noqa: F401
from sample_factory_b200.pbt import (  # noqa: F401
    HYPERPARAMS_TO_TUNE,
    REWARD_CATEGORIES_TO_TUNE,
    SPECIAL_PERTURBATION,
    PopulationBasedTraining,
    perturb_batch_size,
    perturb_exponential_decay,
    perturb_float,
    perturb_vtrace,
    policy_cfg_file,
    policy_reward_shaping_file,
)
