"""HBM-resident trajectory store with the reference layout (algo/utils/shared_buffers.py:79-117).

One allocation per key, shape [num_traj, T(+1), ...], dtypes exactly as the reference allocates them (actions and
policy_version are float32, dones / time_outs / valids are bool, policy_id int32), poison-filled the same way
(shared_buffers.py:45-49,107-115).  In the reference these live in shared memory and slices travel through queues
(BufferMgr :152-239); here they never leave the GPU -- the sampler writes slot [:, t] in place and the learner reads
the whole set, so the Batcher copy (batcher.py:170-218) disappears.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import Tensor

MAGIC_FLOAT = -4242.42  # algo/utils/misc.py:19
MAGIC_INT = 43  # algo/utils/misc.py:20


def alloc_trajectory_tensors(obs_dim: int, num_action_params: int, num_traj: int, rollout: int, device,
                             rnn_size: int = 1, num_actions: int = 1, obs_uint8: bool = False) -> Dict[str, Tensor]:
    T, B = rollout, num_traj
    f32 = dict(dtype=torch.float32, device=device)
    t: Dict[str, Tensor] = {}
    if obs_uint8:   # image observations keep the observation space's dtype (shared_buffers.py:88-96)
        t["obs"] = torch.full((B, T + 1, obs_dim), MAGIC_INT, dtype=torch.uint8, device=device)
    else:
        t["obs"] = torch.full((B, T + 1, obs_dim), MAGIC_FLOAT, **f32)
    t["rnn_states"] = torch.full((B, T + 1, rnn_size), MAGIC_FLOAT, **f32)
    t["actions"] = torch.full((B, T, num_actions), MAGIC_FLOAT, **f32)
    t["action_logits"] = torch.full((B, T, num_action_params), MAGIC_FLOAT, **f32)
    t["log_prob_actions"] = torch.full((B, T), MAGIC_FLOAT, **f32)
    t["values"] = torch.full((B, T + 1), MAGIC_FLOAT, **f32)
    t["policy_version"] = torch.full((B, T), MAGIC_FLOAT, **f32)
    t["rewards"] = torch.full((B, T), -42.42, **f32)
    t["dones"] = torch.ones((B, T), dtype=torch.bool, device=device)
    t["time_outs"] = torch.zeros((B, T), dtype=torch.bool, device=device)
    t["policy_id"] = torch.full((B, T), -1, dtype=torch.int32, device=device)
    t["valids"] = torch.zeros((B, T + 1), dtype=torch.bool, device=device)
    return t


def alloc_for_spec(spec, num_traj: int, rollout: int, device) -> Dict[str, Tensor]:
    """Trajectory set for a ModelSpec: `actions` is [.., 1] for Discrete and [.., A] for Box(A); `action_logits` holds the
    distribution parameters (n logits, or 2A = [means | log_std]) -- shared_buffers.py:67-76 policy_output_shapes."""
    return alloc_trajectory_tensors(spec.obs_dim, spec.num_action_params, num_traj, rollout, device,
                                    rnn_size=spec.rnn_state_size, num_actions=spec.action_width,
                                    obs_uint8=spec.obs_uint8)


def trajectory_bytes_per_env_step(obs_dim: int, num_action_params: int, rnn_size: int = 1) -> int:
    """Algorithmic sampler traffic per env step: obs read + the trajectory record written."""
    write = obs_dim * 4 + rnn_size * 4 + 4 + num_action_params * 4 + 4 + 4 + 4 + 4 + 1 + 1 + 4
    return obs_dim * 4 + write
