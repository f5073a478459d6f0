"""sample_factory.algo.utils.rl_utils (rl_utils.py:24-108): the host-side helpers env code imports (the returns / GAE / V-trace
arithmetic of that file lives in csrc/scan.cu)"""
from typing import Sequence

import numpy as np
from torch import Tensor


def total_num_envs(cfg) -> int:
    return cfg.num_workers * cfg.num_envs_per_worker


def total_num_agents(cfg, env_info) -> int:
    return total_num_envs(cfg) * env_info.num_agents


def num_agents_per_worker(cfg, env_info) -> int:
    return cfg.num_envs_per_worker * env_info.num_agents


def samples_per_trajectory(trajectory) -> int:
    shape = trajectory["rewards"].shape
    return shape[0] * shape[1]


def make_dones(terminated, truncated):
    """done = terminated | truncated for bools, arrays, tensors or per-agent sequences (gymnasium's two flags -> one)"""
    if isinstance(terminated, (bool, np.bool_, np.ndarray, Tensor)):
        return terminated | truncated
    if isinstance(terminated, Sequence):
        return [t | truncated[i] for i, t in enumerate(terminated)]
    raise ValueError(f"make_dones: unsupported type {type(terminated)}")
