"""SURVEY section 8(f) row 4 on the device engine: a population of policies (cfg.num_policies > 1), population-based
training (the reference's tests/algo/test_pbt.py shape: 3 policies, high mutation rate, gamma tuned) and multi-agent CPU envs
with inactive agents (sf_examples/train_custom_multi_env.py's env contract)."""
import json
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

T_ROLL = 16


def _cfg(env_name, argv_extra, experiment):
    from sample_factory_b200.cfg import parse_full_cfg, parse_sf_args

    argv = [f"--env={env_name}", f"--experiment={experiment}", "--train_dir=/tmp/sfb200_tests", "--restart_behavior=overwrite",
            "--batched_sampling=True", "--num_workers=1", "--num_envs_per_worker=1", "--worker_num_splits=1", "--seed=3",
            "--save_every_sec=100000", "--experiment_summaries_interval=100000", "--use_rnn=False", f"--rollout={T_ROLL}",
            "--recurrence=1", "--async_rl=False", "--encoder_mlp_layers", "128", "128"] + argv_extra
    parser, _ = parse_sf_args(argv)
    return parse_full_cfg(parser, argv)


def _register_tape_population(name, n_per_policy, obs_dim=64, num_actions=8):
    """a factory that returns ONE policy's share of a device env (the runner passes policy_index / num_policies)"""
    from sample_factory_b200.envs import TapeVecEnv, register_env

    dev = torch.device("cuda", 0)
    tapes = {}

    def make(_name, cfg, env_config, render_mode=None):
        p = int((env_config or {}).get("policy_index", 0))
        if p not in tapes:
            tapes[p] = torch.randn(2 * T_ROLL + 1, n_per_policy, obs_dim, generator=torch.Generator().manual_seed(100 + p)).to(dev)
        return TapeVecEnv(tapes[p], num_actions, env_index_offset=p * n_per_policy)

    register_env(name, make)


def test_population_members_run_concurrently_and_match_standalone_runners():
    """Two policies on two CUDA streams inside MultiPolicyRunner == the same two single-policy runners run one after the other
    (members are independent: bit-identical weights, statistics and trajectories)."""
    from sample_factory_b200.multi_policy import MultiPolicyRunner
    from sample_factory_b200.train import Runner

    N = 512
    _register_tape_population("tape_population", N)
    extra = [f"--batch_size={N * T_ROLL // 2}", "--num_batches_per_epoch=2", "--num_policies=2"]
    cfg = _cfg("tape_population", extra, "multi_policy_a")
    mp = MultiPolicyRunner(cfg)
    assert mp.init() == 0 and len(mp.subs) == 2 and mp.streams is not None
    for _ in range(3):
        mp.iteration()
    torch.cuda.synchronize()
    assert mp.env_steps_per_policy == [3 * N * T_ROLL] * 2 and mp.env_steps == 6 * N * T_ROLL
    assert not torch.equal(mp.subs[0].model.flat, mp.subs[1].model.flat)          # different members, different weights
    for p in range(2):
        assert int(mp.subs[p].traj["policy_id"].min()) == int(mp.subs[p].traj["policy_id"].max()) == p
        solo = Runner(MultiPolicyRunner(_cfg("tape_population", extra, "multi_policy_b")).member_cfg(p), population=(p, 2))
        solo.init()
        for _ in range(3):
            solo.iteration()
        torch.cuda.synchronize()
        assert torch.equal(solo.model.flat, mp.subs[p].model.flat), p
        assert torch.equal(solo.traj["actions"], mp.subs[p].traj["actions"])
        assert torch.equal(solo.learner.minibatch_log(), mp.subs[p].learner.minibatch_log())
    # per-policy checkpoints and summaries directories
    mp.cfg.train_for_env_steps = mp.env_steps          # run(): no more iterations, final save
    assert mp.run() == 0
    d = os.path.join("/tmp/sfb200_tests", "multi_policy_a")
    assert os.path.isdir(os.path.join(d, "checkpoint_p0")) and os.path.isdir(os.path.join(d, "checkpoint_p1"))
    assert os.path.isdir(os.path.join(d, ".summary", "0")) and os.path.isdir(os.path.join(d, ".summary", "1"))


def test_pbt_replaces_the_worst_policy_and_mutates_hyperparameters():
    """PBT on three policies: with forced objectives the worst member takes the best member's weights / normaliser / optimiser
    state (policy version advanced by max_policy_lag + 1), gets mutated hyper-parameters that reach its learner's kernels
    (graphs re-captured), the best member is left alone, policy 0 is never mutated; the json files are the reference's."""
    from collections import deque

    from sample_factory_b200.multi_policy import MultiPolicyRunner

    random.seed(7)
    N = 256
    _register_tape_population("tape_population3", N)
    cfg = _cfg("tape_population3", [f"--batch_size={N * T_ROLL // 2}", "--num_batches_per_epoch=2", "--num_policies=3",
                                    "--with_pbt=True", "--pbt_period_env_steps=1000", "--pbt_start_mutation=1000",
                                    "--pbt_mutation_rate=0.9", "--pbt_optimize_gamma=True", "--pbt_replace_fraction=0.3",
                                    "--pbt_target_objective=true_objective", "--learner_cuda_graph=True"], "pbt_a")
    mp = MultiPolicyRunner(cfg)
    assert mp.init() == 0
    d = os.path.join("/tmp/sfb200_tests", "pbt_a")
    cfg0 = json.load(open(os.path.join(d, "policy_00_cfg.json")))
    assert cfg0["learning_rate"] == cfg.learning_rate and "gamma" in cfg0        # policy 0 starts from the defaults
    assert json.load(open(os.path.join(d, "policy_01_cfg.json"))) != cfg0         # the others start mutated (rate 0.9)
    assert mp.subs[1].learner.cfg.learning_rate == mp.pbt.policy_cfg[1]["learning_rate"]
    for _ in range(3):
        mp.iteration()
    torch.cuda.synchronize()
    # objectives: policy 2 best, policy 1 worst, policy 0 in the middle
    mp.policy_avg_stats["true_objective"] = [deque([1.0]), deque([-5.0]), deque([4.0])]
    before = [s.model.flat.clone() for s in mp.subs]
    steps_before = [s.learner.train_step for s in mp.subs]
    lr1_before = mp.subs[1].learner.cfg.learning_rate
    mp.pbt.on_training_step()
    torch.cuda.synchronize()
    assert mp.pbt.num_replacements == 1
    assert torch.equal(mp.subs[1].model.flat, before[2]) and torch.equal(mp.subs[2].model.flat, before[2])
    assert torch.equal(mp.subs[1].model.exp_avg_sq, mp.subs[2].model.exp_avg_sq)
    assert torch.equal(mp.subs[1].model.obs_mean, mp.subs[2].model.obs_mean)
    assert mp.subs[1].learner.train_step == steps_before[1] + cfg.max_policy_lag + 1
    assert mp.subs[1].learner.opt_step == mp.subs[2].learner.opt_step
    assert torch.equal(mp.subs[0].model.flat, before[0])                          # the middle policy keeps its weights ...
    assert mp.pbt.policy_cfg[0] == cfg0                                           # ... and policy 0 is never mutated
    new1 = json.load(open(os.path.join(d, "policy_01_cfg.json")))
    assert new1 == mp.pbt.policy_cfg[1] and mp.subs[1].learner.cfg.learning_rate == new1["learning_rate"]
    assert new1 != mp.pbt.policy_cfg[2] or lr1_before != new1["learning_rate"]
    assert 0.0 < new1["gamma"] < 1.0
    # training goes on with the new weights and hyper-parameters (graphs re-captured), all experience of the replaced policy that
    # was collected before the swap is invalid (policy lag)
    for _ in range(2):
        mp.iteration()
    torch.cuda.synchronize()
    for s in mp.subs:
        assert torch.isfinite(s.model.flat).all()
    assert not torch.equal(mp.subs[1].model.flat, mp.subs[2].model.flat)


class _TwoAgentEnv:
    """the game of sf_examples/train_custom_multi_env.py (both agents get -1 unless they pick the same action; random
    inactive phases reported through info["is_active"]), vector observations"""

    def __init__(self, episode_len=10, seed=0):
        from gymnasium import spaces

        self.num_agents, self.is_multiagent = 2, True
        self.observation_space = spaces.Box(0.0, 1.0, (8,), np.float32)
        self.action_space = spaces.Discrete(2)
        self.rng = np.random.RandomState(seed)
        self.episode_len, self.t = episode_len, 0
        self.inactive_steps = [3, 0]
        self.total_steps = 0
        self.shaping = dict(rew=-1.0)

    def _obs(self):
        return [self.rng.rand(8).astype(np.float32) for _ in range(2)]

    def reset(self, **kwargs):
        self.t = 0
        return self._obs(), [dict(), dict()]

    def step(self, actions):
        infos = []
        for j in range(2):
            if self.inactive_steps[j] > 0:
                self.inactive_steps[j] -= 1
            elif j == 1 and self.total_steps == 5:
                self.inactive_steps[j] = 2          # agent 1 reports "inactive" at its steps 5 and 6
            infos.append(dict(is_active=self.inactive_steps[j] <= 0))
        self.total_steps += 1
        self.t += 1
        r = 0.0 if actions[0] == actions[1] else self.shaping["rew"]
        rewards = [r if infos[j]["is_active"] else 0.0 for j in range(2)]
        done = self.t >= self.episode_len
        obs = self.reset()[0] if done else self._obs()
        return obs, rewards, [done] * 2, [done] * 2, infos

    def get_default_reward_shaping(self):
        return dict(self.shaping)

    def set_reward_shaping(self, shaping, agent_idx):
        self.shaping = dict(shaping)


def test_multi_agent_host_env_rows_inactive_agents_and_reward_shaping():
    """a plain multi-agent CPU env is adapted automatically: agent j of env i is row i * A + j, the env auto-resets, steps of
    inactive agents are stamped with policy id -1 (masked by the learner), PBT's reward shaping reaches every instance"""
    from sample_factory_b200.envs import register_env
    from sample_factory_b200.train import Runner

    made = []

    def make(name, cfg, env_config, render_mode=None):
        made.append(_TwoAgentEnv(seed=len(made)))
        return made[-1]

    register_env("two_agent_env", make)
    n_envs = 8
    cfg = _cfg("two_agent_env", ["--num_workers=2", f"--num_envs_per_worker={n_envs // 2}", f"--batch_size={2 * n_envs * T_ROLL}",
                                 "--num_batches_per_epoch=1", "--encoder_mlp_layers", "32", "32"], "multi_agent_a")
    r = Runner(cfg)
    r.init()
    assert len(made) == n_envs and r.env.num_agents == 2 * n_envs and r.env.multi_agent
    r.training_info["reward_shaping"] = dict(rew=-2.0)
    r.iteration()
    torch.cuda.synchronize()
    assert all(e.shaping == dict(rew=-2.0) for e in made)
    pid = r.traj["policy_id"].cpu()
    rew = r.traj["rewards"].cpu()
    dones = r.traj["dones"].cpu()
    assert set(pid.unique().tolist()) == {-1, 0}
    # agent 0 of env 0 reports is_active=False after its first two steps: the status an agent reported at step t - 1 marks
    # step t (non_batched_sampling.py:197-203)
    assert pid[0, 0] == 0 and (pid[0, 1:3] == -1).all() and pid[0, 3] == 0
    assert (pid[1, :6] == 0).all() and (pid[1, 6:8] == -1).all() and (pid[1, 8:] == 0).all()
    assert set(rew.unique().tolist()) <= {0.0, -2.0 * cfg.reward_scale}
    assert dones[:, 9].all() and not dones[:, :9].any()           # episode_len 10, the env resets itself
    # both agents of an env share the payout on steps where both REPORT active (steps 2-4 and 7-9 of the first episode)
    assert (rew[0::2, 2:5] == rew[1::2, 2:5]).all() and (rew[0::2, 7:10] == rew[1::2, 7:10]).all()
    assert (rew[0::2, :2] == 0).all() and (rew[1::2, 5:7] == 0).all()             # the env pays nothing to an inactive agent
    st = r.learner.fetch_stats()
    assert np.isfinite(st["loss"])
    assert int((pid == -1).sum()) == n_envs * 4
