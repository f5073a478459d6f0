"""sample_factory.algo.utils.misc (misc.py:7-33): constants example code imports"""
EPS = 1e-8

EPISODIC = "episodic"
LEARNER_ENV_STEPS = "learner_env_steps"
TRAIN_STATS = "train"
TIMING_STATS = "timing"
STATS_KEY = "stats"
SAMPLES_COLLECTED = "samples_collected"
POLICY_ID_KEY = "policy_id"

MAGIC_FLOAT = -4242.42
MAGIC_INT = 43


class ExperimentStatus:
    SUCCESS, FAILURE, INTERRUPTED = range(3)
