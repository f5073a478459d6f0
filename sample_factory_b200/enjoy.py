"""`enjoy(cfg)`: run a trained policy and report the mean episode reward -- the reference's sample_factory/enjoy.py:103-295
over the device sampler.

Same contract: the experiment's saved `config.json` is loaded and overridden by explicitly passed CLI flags
(`load_from_checkpoint`, cfg/arguments.py:227-260), the latest (or `--load_checkpoint_kind=best`) checkpoint is loaded
into the policy (enjoy.py:92-100; RuntimeError if there is none), the env is stepped until `max_num_episodes` episodes
or `max_num_frames` vector steps are done, actions are sampled -- or, with `--eval_deterministic=True`, the argmax of the
action probabilities / the Gaussian means (enjoy.py:165-171) -- and the result is (status, mean reward of the finished
episodes).  Rendering, video and hub upload are host-side tooling outside this path.
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Tuple

import numpy as np
import torch

from .checkpoint import load_checkpoint
from .sampling_api import StatusCode, _DeviceSamplingLoop


def cfg_file(cfg) -> str:
    """cfg/arguments.py:215-218"""
    return os.path.join(cfg.train_dir, cfg.experiment, "config.json")


def load_from_checkpoint(cfg) -> argparse.Namespace:
    """cfg/arguments.py:227-260: saved parameters, overridden by what was passed on the command line, completed by any
    parameter the saved file does not know."""
    filename = cfg_file(cfg)
    if not os.path.isfile(filename):
        raise Exception(f"Could not load saved parameters for experiment {cfg.experiment} (file {filename} not found). "
                        "Check that you have the correct experiment name and --train_dir is set correctly.")
    with open(filename, "r") as f:
        loaded = json.load(f)
    for key, value in getattr(cfg, "cli_args", {}).items():
        if key in loaded and loaded[key] != value:
            loaded[key] = value
    for key, value in vars(cfg).items():
        if key not in loaded:
            loaded[key] = value
    return argparse.Namespace(**loaded)


def checkpoint_override_defaults(cfg, parser):
    """cfg/arguments.py:278-295: make the saved experiment configuration the parser's defaults"""
    from .cfg import AttrDict

    filename = cfg_file(cfg)
    if not os.path.isfile(filename):
        raise Exception(f"Could not load saved parameters for experiment {cfg.experiment} (file {filename} not found). "
                        "Check that you have the correct experiment name and --train_dir is set correctly.")
    with open(filename, "r") as f:
        loaded = AttrDict(json.load(f))
    parser.set_defaults(**loaded)
    return loaded


def enjoy(cfg) -> Tuple[int, float]:
    cfg = load_from_checkpoint(cfg)
    eval_frameskip = cfg.env_frameskip if getattr(cfg, "eval_env_frameskip", None) is None else cfg.eval_env_frameskip
    assert cfg.env_frameskip % eval_frameskip == 0, f"{cfg.env_frameskip=} must be divisible by {eval_frameskip=}"
    cfg.env_frameskip = cfg.eval_env_frameskip = eval_frameskip
    loop = _DeviceSamplingLoop(cfg, None, None, record_episodes=True,
                               deterministic=bool(getattr(cfg, "eval_deterministic", False)))
    ck = load_checkpoint(cfg, loop.model, loop.device, kind=getattr(cfg, "load_checkpoint_kind", "latest"),
                         policy_id=getattr(cfg, "policy_index", 0))
    if ck is None:
        raise RuntimeError("Could not load checkpoint")
    loop.start(ck["train_step"])
    max_frames = getattr(cfg, "max_num_frames", None)
    max_episodes = getattr(cfg, "max_num_episodes", int(1e9))
    rewards = []
    num_frames = 0
    while (max_frames is None or num_frames <= max_frames) and len(rewards) < max_episodes:
        loop.rollout()
        num_frames += cfg.rollout                     # one "frame" per vector-env step, as the reference counts them
        ret, _ = loop.sampler.finished_episodes()
        rewards.extend(float(r) for r in ret)
    torch.cuda.synchronize()
    if hasattr(loop.env, "close"):
        loop.env.close()
    rewards = rewards[:max_episodes]
    if not rewards:
        return StatusCode.SUCCESS, float("nan")
    avg = float(np.mean(rewards))
    print(f"[sf_b200] enjoy: {len(rewards)} episodes, {num_frames} frames, avg episode reward {avg:.3f}", flush=True)
    return StatusCode.SUCCESS, avg
