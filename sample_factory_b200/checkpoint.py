"""Checkpoint / resume in the reference's format (algo/learning/learner.py:257-386):

    <train_dir>/<experiment>/checkpoint_p<policy_id>/checkpoint_<train_step:09d>_<env_steps>.pth
    torch.save(dict(train_step, env_steps, best_performance, model=state_dict, optimizer=Adam.state_dict, curr_lr))

written to a *_temp file and renamed (:334-360), keeping the last --keep_checkpoints files.  `model` uses the
reference's state_dict keys (PolicyModel.state_dict), so reference tooling (enjoy / eval) can load our runs and we can
warm-start from reference checkpoints.
"""
from __future__ import annotations

import glob
import os
from typing import Optional

import torch


def checkpoint_dir(cfg, policy_id: int = 0) -> str:
    d = os.path.join(cfg.train_dir, cfg.experiment, f"checkpoint_p{policy_id}")
    os.makedirs(d, exist_ok=True)
    return d


def get_checkpoints(directory: str, pattern: str = "checkpoint_*"):
    return sorted(f for f in glob.glob(os.path.join(directory, pattern)) if not f.endswith("_temp"))


def save_checkpoint(cfg, model, learner, prefix: str = "checkpoint", name_suffix: str = "", keep: Optional[int] = None) -> str:
    """Learner._save_impl (learner.py:334-360)"""
    d = checkpoint_dir(cfg, learner.policy_id)
    ck = dict(
        train_step=learner.train_step,
        env_steps=learner.env_steps,
        best_performance=getattr(learner, "best_performance", -1e9),
        model={k: v.cpu() for k, v in model.state_dict().items()},
        optimizer=_cpu(model.optimizer_state_dict(learner.opt_step, learner.curr_lr, (cfg.adam_beta1, cfg.adam_beta2),
                                                  cfg.adam_eps)),
        curr_lr=learner.curr_lr,
    )
    name = f"{prefix}_{learner.train_step:09d}_{learner.env_steps}{name_suffix}.pth"
    tmp = os.path.join(d, name + "_temp")
    torch.save(ck, tmp)
    final = os.path.join(d, name)
    os.rename(tmp, final)
    files = get_checkpoints(d, f"{prefix}_*")
    while len(files) > (cfg.keep_checkpoints if keep is None else keep):
        os.remove(files.pop(0))
    return final


def save_best(cfg, model, learner, metric: str, metric_value: float) -> Optional[str]:
    """Learner.save_best (learner.py:376-386): a `best_*` file whenever the metric improved by more than 1e-3"""
    p = 3
    if metric_value - getattr(learner, "best_performance", -1e9) > 1 / 10 ** p:
        learner.best_performance = metric_value
        return save_checkpoint(cfg, model, learner, "best", f"_{metric}_{metric_value:.{p}f}", keep=1)
    return None


def _cpu(osd: dict) -> dict:
    for st in osd["state"].values():
        for k, v in st.items():
            if torch.is_tensor(v):
                st[k] = v.cpu()
    return osd


def load_checkpoint(cfg, model, device, kind: str = "latest", policy_id: int = 0) -> Optional[dict]:
    """Learner.load_from_checkpoint (learner.py:257-310) / enjoy.load_state_dict (enjoy.py:92-100): newest
    `checkpoint_*` file, or newest `best_*` file for kind="best"."""
    d = checkpoint_dir(cfg, policy_id)
    files = get_checkpoints(d, dict(latest="checkpoint", best="best")[kind] + "_*")
    if not files:
        return None
    ck = torch.load(files[-1], map_location="cpu", weights_only=False)
    ours = set(model.state_dict().keys())
    missing, unexpected = sorted(ours - set(ck["model"])), sorted(set(ck["model"]) - ours)
    if missing or unexpected:     # strict=False like the reference's enjoy; say what did not line up instead of hiding it
        print(f"[sf_b200] checkpoint {os.path.basename(files[-1])}: keys missing in the file {missing[:6]}"
              f"{'...' if len(missing) > 6 else ''}, keys the model does not have {unexpected[:6]}{'...' if len(unexpected) > 6 else ''}",
              flush=True)
    model.load_state_dict(ck["model"], strict=False)
    opt_step = model.load_optimizer_state_dict(ck["optimizer"]) if "optimizer" in ck else 0
    return dict(train_step=int(ck["train_step"]), env_steps=int(ck["env_steps"]), opt_step=opt_step,
                curr_lr=ck.get("curr_lr", cfg.learning_rate))
