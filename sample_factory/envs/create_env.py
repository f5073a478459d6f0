"""sample_factory.envs.create_env (envs/create_env.py:13-46)."""
from sample_factory_b200.envs import create_env  # noqa: F401
