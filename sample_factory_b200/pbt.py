"""Population-based training over the policies of a MultiPolicyRunner -- the reference's PopulationBasedTraining observer
(pbt/population_based_training.py:107-415, after Jaderberg et al. 2018) with the same decision rules and the same files:

  * every policy owns a dict of the tunable hyper-parameters (HYPERPARAMS_TO_TUNE :56-64, + gamma with
    --pbt_optimize_gamma) stored in <experiment>/policy_XX_cfg.json and a reward-shaping scheme in
    policy_XX_reward_shaping.json; policy 0 starts from the defaults, the others from one mutation (:146-180);
  * every pbt_period_env_steps (after pbt_start_mutation) a policy is ranked by the mean of its recent
    cfg.pbt_target_objective values; the top ceil(replace_fraction * P) are left alone; everybody else mutates its own
    parameters, and a policy among the worst additionally takes weights AND parameters of a random top policy if the gap
    is larger than pbt_replace_reward_gap (relative) and pbt_replace_reward_gap_absolute (:302-374);
  * a replacement goes through a checkpoint of the donor (save, then load with load_progress=False, learner.py:300-310,
    415-428) and advances the receiver's policy version by max_policy_lag + 1 so that experience in flight is dropped.

The reference spreads this over signal/slot messages between runner, learners and PBT (save_model -> saved_model ->
load_model); here learners live in one process, so the same sequence is three calls.  Randomness comes from `random`
like the reference's; under data parallelism rank 0 decides and broadcasts (multi_policy.py)."""
from __future__ import annotations

import copy
import json
import math
import os
import random
import time
from typing import Dict, List, Optional, SupportsFloat

EPS = 1e-8      # algo/utils/misc.py:20


def perturb_float(x, perturb_amount=1.2):
    """:25-28: multiply or divide, with equal probability"""
    return x / perturb_amount if random.random() < 0.5 else x * perturb_amount


def perturb_vtrace(x, _cfg):
    return perturb_float(x, perturb_amount=1.005)


def perturb_exponential_decay(x, _cfg, perturb_amount_min=1.01, perturb_amount_max=1.2):
    """:35-42: gamma-like parameters move in (1 - x) space, conservatively"""
    perturb_amount = random.uniform(perturb_amount_min, perturb_amount_max)
    return max(EPS, 1.0 - perturb_float(1.0 - x, perturb_amount=perturb_amount))


def perturb_batch_size(x, cfg):
    """:45-58 (not in the tuned set: it would change the learner's static shapes, as it would the reference's batcher)"""
    new_value = min(perturb_float(x, perturb_amount=1.2), cfg.batch_size * 1.5)
    new_value = (int(new_value) // cfg.rollout) * cfg.rollout
    return max(new_value, cfg.rollout)


HYPERPARAMS_TO_TUNE = {"learning_rate", "exploration_loss_coeff", "value_loss_coeff", "max_grad_norm", "ppo_clip_ratio",
                       "ppo_clip_value"}
REWARD_CATEGORIES_TO_TUNE = {"doom_": ["delta", "selected_weapon"]}
SPECIAL_PERTURBATION = dict(gamma=perturb_exponential_decay, adam_beta1=perturb_exponential_decay,
                            vtrace_rho=perturb_vtrace, vtrace_c=perturb_vtrace, batch_size=perturb_batch_size)


def policy_cfg_file(experiment_dir: str, policy_id: int) -> str:
    return os.path.join(experiment_dir, f"policy_{policy_id:02d}_cfg.json")


def policy_reward_shaping_file(experiment_dir: str, policy_id: int) -> str:
    return os.path.join(experiment_dir, f"policy_{policy_id:02d}_reward_shaping.json")


def _iter_leaves(d: dict, d_default):
    """(dict, key, value, default value) over the leaves of two identically shaped nested dicts (utils/dicts.py)"""
    for k, v in list(d.items()):
        dv = d_default[k] if isinstance(d_default, dict) else getattr(d_default, k)
        if isinstance(v, dict):
            yield from _iter_leaves(v, dv)
        else:
            yield d, k, v, dv


class PopulationBasedTraining:
    def __init__(self, cfg, runner, log=print):
        self.cfg, self.runner, self.log = cfg, runner, log
        self.tuned = set(HYPERPARAMS_TO_TUNE)
        if cfg.pbt_optimize_gamma:
            self.tuned.add("gamma")
        P = cfg.num_policies
        self.last_update = [0] * P
        self.policy_cfg: List[dict] = [dict() for _ in range(P)]
        self.policy_reward_shaping: List[Optional[dict]] = [None] * P
        self.default_reward_shaping: Optional[dict] = None
        self.last_pbt_summaries = 0.0
        self.reward_categories_to_tune: List[str] = []
        for env_prefix, categories in REWARD_CATEGORIES_TO_TUNE.items():
            if str(cfg.env).startswith(env_prefix):
                self.reward_categories_to_tune = categories
        self.num_replacements = 0

    # ---- start-up (:140-196) -----------------------------------------------------------------------------------
    def on_init(self, experiment_dir: str, default_reward_shaping: Optional[dict]) -> None:
        self.dir = experiment_dir
        self.default_reward_shaping = default_reward_shaping
        for p in range(self.cfg.num_policies):
            f = policy_cfg_file(experiment_dir, p)
            if os.path.exists(f):
                with open(f) as fh:
                    self.policy_cfg[p] = json.load(fh)
            else:
                self.policy_cfg[p] = {name: getattr(self.cfg, name) for name in sorted(self.tuned)}
                if p > 0:                           # keep one policy with default settings in the beginning
                    self.policy_cfg[p] = self._perturb_cfg(self.policy_cfg[p])
            f = policy_reward_shaping_file(experiment_dir, p)
            if os.path.exists(f):
                with open(f) as fh:
                    self.policy_reward_shaping[p] = json.load(fh)
            else:
                self.policy_reward_shaping[p] = copy.deepcopy(default_reward_shaping)
                if p > 0:
                    self.policy_reward_shaping[p] = self._perturb_reward(self.policy_reward_shaping[p])
        for p in range(self.cfg.num_policies):
            self._save(p)

    def on_start(self) -> None:
        for p in range(self.cfg.num_policies):
            self.runner.update_policy_cfg(p, self.policy_cfg[p])
            self.runner.update_reward_shaping(p, self.policy_reward_shaping[p])

    def _save(self, p: int) -> None:
        with open(policy_cfg_file(self.dir, p), "w") as fh:
            json.dump(self.policy_cfg[p], fh)
        with open(policy_reward_shaping_file(self.dir, p), "w") as fh:
            json.dump(self.policy_reward_shaping[p], fh)

    # ---- mutation (:210-275) -----------------------------------------------------------------------------------
    def _perturb_param(self, param, param_name, default_param):
        if random.random() > self.cfg.pbt_mutation_rate:          # toss a coin whether we perturb the parameter at all
            return param
        if param != default_param and random.random() < 0.01:     # small chance to go back to the default value
            return default_param
        if param_name in SPECIAL_PERTURBATION:
            return SPECIAL_PERTURBATION[param_name](param, self.cfg)
        if type(param) is bool:
            return not param
        if isinstance(param, SupportsFloat):
            return perturb_float(float(param), perturb_amount=random.uniform(self.cfg.pbt_perturb_min, self.cfg.pbt_perturb_max))
        raise RuntimeError("Unsupported parameter type")

    def _perturb(self, old_params: dict, default_params) -> dict:
        params = copy.deepcopy(old_params)
        for d, key, value, value_default in _iter_leaves(params, default_params):
            if isinstance(value, (tuple, list)):     # reward shaping "delta" parameters: (negative change, positive change)
                d[key] = tuple(self._perturb_param(x, f"{key}_{i}", value_default[i]) for i, x in enumerate(value))
            else:
                d[key] = self._perturb_param(value, key, value_default)
        return params

    def _perturb_cfg(self, original_cfg: dict) -> dict:
        return self._perturb(copy.deepcopy(original_cfg), default_params=self.cfg)

    def _perturb_reward(self, original: Optional[dict]) -> Optional[dict]:
        if original is None:
            return None
        shaping = copy.deepcopy(original)
        if self.reward_categories_to_tune:
            for category in self.reward_categories_to_tune:
                if category in shaping:
                    shaping[category] = self._perturb(shaping[category], self.default_reward_shaping[category])
            return shaping
        return self._perturb(shaping, self.default_reward_shaping)

    # ---- selection (:302-374) ----------------------------------------------------------------------------------
    def decide(self, policy_id: int, target_objectives: List[float]) -> Optional[int]:
        """The reference's _update_policy up to the point where messages are sent: mutates self.policy_cfg /
        policy_reward_shaping of `policy_id` and returns the policy whose weights it takes (itself = keep its own), or None
        if `policy_id` is among the best and is left alone."""
        cfg = self.cfg
        P = cfg.num_policies
        ranked = [p for _obj, p in sorted(zip(target_objectives, range(P)), reverse=True)]
        replace_number = math.ceil(cfg.pbt_replace_fraction * P)
        best, worst = ranked[:replace_number], ranked[-replace_number:]
        if policy_id in best:
            return None
        replacement = policy_id
        if policy_id in worst:
            candidate = random.choice(best)
            delta = target_objectives[candidate] - target_objectives[policy_id]
            delta_rel = abs(delta / (target_objectives[candidate] + EPS))
            if abs(delta) > cfg.pbt_replace_reward_gap_absolute and delta_rel > cfg.pbt_replace_reward_gap:
                replacement = candidate
                self.log(f"[pbt] policy {policy_id} ({target_objectives[policy_id]:.4f}) takes the weights of policy "
                         f"{candidate} ({target_objectives[candidate]:.4f})")
        if policy_id == 0:       # never mutate the first policy (kept as the reference point); it may still be replaced
            self.policy_cfg[0] = self.policy_cfg[replacement]
            self.policy_reward_shaping[0] = self.policy_reward_shaping[replacement]
        else:
            self.policy_cfg[policy_id] = self._perturb_cfg(self.policy_cfg[replacement])
            self.policy_reward_shaping[policy_id] = self._perturb_reward(self.policy_reward_shaping[replacement])
        return replacement

    def apply(self, policy_id: int, replacement: int) -> None:
        """save_model(replacement) -> on_saved_model -> load_model / update_cfg / update_reward_shaping (:376-397)"""
        if replacement != policy_id:
            self.runner.replace_policy(policy_id, replacement)
            self.num_replacements += 1
        self._save(policy_id)
        self.runner.update_policy_cfg(policy_id, self.policy_cfg[policy_id])
        self.runner.update_reward_shaping(policy_id, self.policy_reward_shaping[policy_id])

    # ---- per training iteration (:399-424) ---------------------------------------------------------------------
    def objectives(self) -> Optional[List[float]]:
        stats = self.runner.policy_avg_stats.get(self.cfg.pbt_target_objective)
        if stats is None:
            return None
        out = []
        for p in range(self.cfg.num_policies):
            vals = [v for v in stats[p] if v == v]
            if not vals:
                return None                        # not enough data to perform PBT yet
            out.append(float(sum(vals) / len(vals)))
        return out

    def on_training_step(self) -> None:
        cfg = self.cfg
        if not cfg.with_pbt or cfg.num_policies <= 1:
            return
        env_steps = self.runner.env_steps_per_policy
        for p in range(cfg.num_policies):
            if env_steps[p] < cfg.pbt_start_mutation:
                continue
            if env_steps[p] - self.last_update[p] > cfg.pbt_period_env_steps:
                decision = self.runner.pbt_decide(self, p)         # (rank 0 decides, every rank applies)
                if decision is not None:
                    self.apply(p, decision)
                self._write_summaries(p, env_steps[p])
                self.last_update[p] = env_steps[p]
        now = time.time()
        if now - self.last_pbt_summaries > 5 * 60:
            for p in range(cfg.num_policies):
                self._write_summaries(p, env_steps[p])
            self.last_pbt_summaries = now

    def _write_summaries(self, p: int, env_steps: int) -> None:
        w = self.runner.writers.get(p)
        if w is None:
            return
        for name, d in (("cfg", self.policy_cfg[p]), ("rew", self.policy_reward_shaping[p])):
            if d is None:
                continue
            for _d, key, value, _dv in _iter_leaves(d, d):
                if isinstance(value, bool):
                    value = int(value)
                if isinstance(value, (int, float)):
                    w.add_scalar(f"zz_pbt/{name}_{key}", value, env_steps)
                elif isinstance(value, (tuple, list)):
                    for i, x in enumerate(value):
                        w.add_scalar(f"zz_pbt/{name}_{key}_{i}", x, env_steps)
