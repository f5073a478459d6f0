"""Multi-policy training (cfg.num_policies > 1) and population-based training on the device engine -- SURVEY section 8(f)
row 4.  In the reference this is the non-batched sampler's territory (non_batched_sampling.py:313-703: per-agent policy
ids, one inference worker and one learner PROCESS per policy, runner.py:626-678) plus the PBT observer
(pbt/population_based_training.py).  Here a population is P single-policy runners in ONE process per GPU:

  * agent -> policy mapping: the reference's sync-mode rule (agent_policy_mapping.py:39-45: global env index % num_policies,
    "deterministic mapping ensures we always collect the same amount of experience per policy per iteration"): policy p owns
    1/P of the env instances for the whole run -- its own batched env, trajectory buffers, sampler, model and learner
    (train.Runner with population=(p, P)); every trajectory is stamped with its policy id and the learner's valid mask
    (learner.py:946-953) works unchanged;
  * execution: the P rollout+train iterations are enqueued on P CUDA streams (fork / join per iteration).  A population
    member at N/P envs is even more latency-bound than the single policy at N, so the members' kernel chains fill each
    other's gaps; with CUDA graphs an iteration is two launches per member.  Under data parallelism (world_size > 1) the
    members run one after the other on one stream: every rank must reach the members' gradient exchanges in the same order;
  * statistics, summaries (one tensorboard writer per policy, runner.py:199-205), checkpoints (checkpoint_p<id>/) per policy;
  * PBT (pbt.py): decisions from cfg.pbt_target_objective, weight replacement as device-to-device copies between two
    members (plus the donor's checkpoint on disk, as the reference writes it), hyper-parameters through Learner.set_new_cfg,
    reward shaping through the env's RewardShapingInterface.

Inference for every member is always batched on the device; what the reference's non-batched sampler adds for CPU envs --
stepping heterogeneous env instances one by one, multi-agent envs, inactive agents -- is host_env.BatchedHostEnv."""
from __future__ import annotations

import copy
import time
from collections import deque
from typing import Dict, List, Optional

import torch

from .checkpoint import save_best, save_checkpoint
from .pbt import PopulationBasedTraining
from .train import Runner, StatusCode, experiment_dir


class MultiPolicyRunner:
    def __init__(self, cfg):
        assert cfg.num_policies > 1
        self.cfg = cfg
        self.subs: List[Runner] = []
        self.policy_avg_stats: Dict[str, List[deque]] = {}
        self.writers: Dict[int, object] = {}
        self.observers: List = []
        self.total_train_seconds = 0.0
        self.pbt: Optional[PopulationBasedTraining] = PopulationBasedTraining(cfg, self) if cfg.with_pbt else None
        self.initialized = False

    # ---- what PBT and the reference's observers read --------------------------------------------------------------
    @property
    def env_steps_per_policy(self) -> List[int]:
        return [s.env_steps for s in self.subs]

    @property
    def env_steps(self) -> int:
        return sum(s.env_steps for s in self.subs)

    def register_observer(self, observer) -> None:
        self.observers.append(observer)

    def member_cfg(self, p: int):
        c = copy.deepcopy(self.cfg)
        c.policy_id = p
        c.num_policies = 1          # (a member IS a single-policy runner; the population size travels as Runner.population)
        return c

    # ---------------------------------------------------------------------------------------------------------------
    def init(self) -> int:
        cfg = self.cfg
        P = cfg.num_policies
        for p in range(P):
            sub = Runner(self.member_cfg(p), population=(p, P))
            status = sub.init()
            if status != StatusCode.SUCCESS:
                return status
            self.subs.append(sub)
            self.writers.update(sub.writers)
        s0 = self.subs[0]
        self.rank, self.world_size, self.device = s0.rank, s0.world_size, s0.device
        # one stream per member on a single GPU; one shared stream under data parallelism (ordered collectives)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(P)] if self.world_size == 1 else None
        if self.pbt is not None:
            env0 = s0.env
            default_shaping = env0.get_default_reward_shaping() if hasattr(env0, "get_default_reward_shaping") else None
            if self.rank == 0:
                self.pbt.on_init(experiment_dir(cfg), default_shaping)
            else:
                self.pbt.dir, self.pbt.default_reward_shaping = experiment_dir(cfg), default_shaping
            self._pbt_broadcast_all()
            self.pbt.on_start()
        self.initialized = True
        return StatusCode.SUCCESS

    def iteration(self) -> None:
        """one rollout + train of every member"""
        if self.streams is None:
            for sub in self.subs:
                sub.iteration()
            return
        main = torch.cuda.current_stream()
        for sub, st in zip(self.subs, self.streams):           # fork
            st.wait_stream(main)
            with torch.cuda.stream(st):
                sub.iteration()
        for st in self.streams:                                # join
            main.wait_stream(st)

    # ---- PBT callbacks (pbt.PopulationBasedTraining) ----------------------------------------------------------------
    def update_policy_cfg(self, p: int, new_cfg: Dict) -> None:
        self.subs[p].learner.set_new_cfg(new_cfg)

    def update_reward_shaping(self, p: int, shaping: Optional[Dict]) -> None:
        if shaping is not None:     # forwarded to the member's envs before its next rollout (envs.set_training_info)
            self.subs[p].training_info["reward_shaping"] = shaping

    def replace_policy(self, p: int, donor: int) -> None:
        src, dst = self.subs[donor], self.subs[p]
        torch.cuda.synchronize()
        if self.rank == 0:
            save_checkpoint(src.cfg, src.model, src.learner)     # "force replacement policy learner to save its model" (:372-374)
        dst.learner.load_policy_from(src.learner)
        if dst.async_rl:
            dst.sampler_model.copy_weights_from(dst.model)
            dst.snapshot_version = dst.learner.train_step
        torch.cuda.synchronize()

    def pbt_decide(self, pbt: PopulationBasedTraining, p: int) -> Optional[int]:
        """rank 0 ranks the population and mutates; the decision and the mutated parameters go to every rank"""
        decision, have = None, False
        if self.rank == 0:
            obj = pbt.objectives()
            if obj is not None:
                decision = pbt.decide(p, obj)
                have = True
        if self.world_size > 1:
            box = [(have, decision, pbt.policy_cfg[p], pbt.policy_reward_shaping[p])]
            torch.distributed.broadcast_object_list(box, src=0)
            have, decision, pbt.policy_cfg[p], pbt.policy_reward_shaping[p] = box[0]
        return decision if have else None

    def _pbt_broadcast_all(self) -> None:
        if self.world_size > 1:
            box = [(self.pbt.policy_cfg, self.pbt.policy_reward_shaping)]
            torch.distributed.broadcast_object_list(box, src=0)
            self.pbt.policy_cfg, self.pbt.policy_reward_shaping = box[0]

    # ---------------------------------------------------------------------------------------------------------------
    def _collect_stats(self, fps: float) -> None:
        cfg = self.cfg
        P = cfg.num_policies
        for p, sub in enumerate(self.subs):
            ep = sub.sampler.pop_episode_stats()
            if "reward" in ep and "true_objective" not in ep:
                ep["true_objective"] = ep["reward"]      # non_batched_sampling.py:300-301: info["true_objective"], else the reward
            st = sub.learner.fetch_stats()
            for key in ("version_diff_min", "version_diff_avg", "version_diff_max"):
                if key in st:
                    sub.policy_lag[0][key] = st[key]
            for key, val in ep.items():
                self.policy_avg_stats.setdefault(key, [deque(maxlen=cfg.stats_avg) for _ in range(P)])[p].append(val)
                sub.policy_avg_stats.setdefault(key, [deque(maxlen=cfg.stats_avg)])[0].append(val)
            if self.rank == 0:
                sub._report_experiment_summaries(fps / P, st)
                print(f"[sf_b200] policy {p}: env_steps {sub.env_steps} loss {st.get('loss', float('nan')):.4f} "
                      f"reward {ep.get('reward', float('nan')):.3f} episodes {ep.get('episodes', 0)}", flush=True)

    def run(self) -> int:
        cfg = self.cfg
        assert self.initialized
        t_start = last_report = last_save = last_best = time.time()
        steps_at_report = self.env_steps
        status = StatusCode.SUCCESS
        try:
            while self.env_steps < cfg.train_for_env_steps and not self.subs[0]._time_is_up(t_start):
                self.iteration()
                if self.pbt is not None:
                    self.pbt.on_training_step()
                now = time.time()
                if now - last_report >= cfg.experiment_summaries_interval:
                    torch.cuda.synchronize()
                    now = time.time()
                    fps = (self.env_steps - steps_at_report) / (now - last_report)
                    self._collect_stats(fps)
                    if self.rank == 0:
                        print(f"[sf_b200] env_steps {self.env_steps} fps {fps:.0f} ({cfg.num_policies} policies)", flush=True)
                    last_report, steps_at_report = now, self.env_steps
                if now - last_save >= cfg.save_every_sec and self.rank == 0:
                    for sub in self.subs:
                        save_checkpoint(sub.cfg, sub.model, sub.learner)
                    last_save = now
                if now - last_best >= cfg.save_best_every_sec and self.rank == 0:
                    last_best = now
                    hist = self.policy_avg_stats.get(cfg.save_best_metric)
                    for p, sub in enumerate(self.subs):
                        vals = [v for v in hist[p] if v == v] if hist else []
                        if vals and sub.env_steps >= cfg.save_best_after:
                            save_best(sub.cfg, sub.model, sub.learner, cfg.save_best_metric, float(sum(vals) / len(vals)))
        except KeyboardInterrupt:
            status = StatusCode.INTERRUPTED
        torch.cuda.synchronize()
        self.total_train_seconds = time.time() - t_start
        if self.rank == 0:
            for sub in self.subs:
                save_checkpoint(sub.cfg, sub.model, sub.learner)
            fps = self.env_steps / max(self.total_train_seconds, 1e-9)
            collected = {p: s.env_steps for p, s in enumerate(self.subs)}
            print(f"[sf_b200] Collected {collected}, FPS: {fps:.1f}", flush=True)            # runner.py:763-764
            for w in self.writers.values():
                w.close()
        return status
