"""Phase timing inside the persistent rollout kernel (csrc/rollout_fused.cu): %globaltimer stamps of CTA (0,0)'s first epilogue
thread at the phase boundaries of every step -> mean duration of each phase (ns)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import make_cfg, N_ENVS, OBS_DIM, N_ACTIONS, ROLLOUT
from sample_factory_b200 import ops
from sample_factory_b200._lib import lib
from sample_factory_b200.envs import TapeVecEnv, register_env
from sample_factory_b200.train import Runner

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ops.bind_device(dev)
tape = torch.randn(97, N_ENVS, OBS_DIM, device=dev)
register_env("synthetic_tape", lambda name, cfg, env_config, render_mode=None: TapeVecEnv(tape, N_ACTIONS))
runner = Runner(make_cfg("synthetic_tape", "auto", False))
runner.init()
assert runner.sampler.fused_rollout
for _ in range(3):
    runner.sampler.rollout()
trace = torch.zeros(ROLLOUT * 16, dtype=torch.int64, device=dev)
lib().call("sfb200_rollout_set_trace", trace.data_ptr())
runner.sampler.rollout()
torch.cuda.synchronize()
lib().call("sfb200_rollout_set_trace", None)
tr = trace.view(ROLLOUT, 16).cpu().double()
names = ["tile1: start -> accumulator complete", "tile1: drain", "tile1: bias + activation", "tile1: h1 store + fence",
         "tile1: cluster barrier", "tile2: barrier -> accumulator complete", "tile2: drain", "tile2: bias + activation",
         "tile2: head partial dot products", "tile2: partials store", "tile2: cluster barrier", "tail: compute + stores + fence",
         "tail: cluster barrier"]
idx = [(0, 1), (1, 2), (2, 11), (11, 3), (3, 4), (4, 5), (5, 6), (6, 12), (12, 13), (13, 7), (7, 8), (8, 9), (9, 10)]
tot = 0.0
for n, (i, j) in zip(names, idx):
    d = (tr[1:, j] - tr[1:, i]).mean().item()
    tot += d
    print(f"{n:45s} {d:9.0f} ns")
print(f"{'step total':45s} {tot:9.0f} ns   (kernel: {(tr[-1, 10] - tr[0, 0]).item() / 1e3:.1f} us for {ROLLOUT} steps)")
