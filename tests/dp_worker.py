"""Worker for tests/test_gpu_multi.py (launched under torch.distributed.run, one process per GPU, NCCL).

Checks the data-parallel contract of DESIGN.md section 6:  G ranks x N/G envs  ==  1 process x N envs.
Every rank trains on its env shard for two iterations; rank 0 additionally trains a non-parallel learner on the
concatenated batch and compares parameters, normalizer statistics and loss terms."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import appo_oracle as O  # noqa: E402  (only to build inputs: init weights + an oracle rollout)
from sample_factory_b200 import ops  # noqa: E402
from sample_factory_b200.dist_utils import init_from_env  # noqa: E402
from sample_factory_b200.learner import Learner  # noqa: E402
from sample_factory_b200.model import ModelSpec, PolicyModel  # noqa: E402
from sample_factory_b200.trajectory import alloc_trajectory_tensors  # noqa: E402
from tests.test_gpu_engine import make_cfg  # noqa: E402


def main():
    rank, local_rank, world = init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ops.bind_device(dev)
    N, T, NMB = 512, 16, 4
    engine = ops.GEMM_TC_3XTF32 if ops.tc_available() else ops.GEMM_SIMT
    ocfg_full = O.OracleCfg(obs_dim=64, num_actions=8, encoder_mlp_layers=[256, 256], rollout=T, recurrence=1,
                            batch_size=N * T // NMB, num_batches_per_epoch=NMB, kl_loss_coeff=0.1)
    ocfg_loc = O.OracleCfg(obs_dim=64, num_actions=8, encoder_mlp_layers=[256, 256], rollout=T, recurrence=1,
                           batch_size=N * T // NMB // world, num_batches_per_epoch=NMB, kl_loss_coeff=0.1)
    st0 = O.init_state(ocfg_full, seed=7)
    gen = torch.Generator().manual_seed(5)
    # two (three with the graph variant) iterations of data from an oracle rollout (identical on every rank: same seeds)
    tape = torch.randn(2 * T + 1, N, 64, generator=gen)
    env = O.TapeVecEnv(tape, 8)
    last = env.reset()
    batches = []
    dp_graph = os.environ.get("SFB200_DP_GRAPH", "0") == "1"     # opt-in: also check the graph-captured DP learner
    for it in range(3 if dp_graph else 2):
        traj = O.alloc_trajectories(ocfg_full, N)
        noise = torch.empty(T, N, 8).exponential_(generator=gen)
        last = O.rollout(ocfg_full, st0, env, last, traj, noise, 0)
        traj["policy_id"][torch.rand(N, T, generator=gen) < 0.1] = -1     # some invalid samples
        batches.append(traj)

    # env shard of this rank: the single-process minibatch b is envs [b*N/NMB, (b+1)*N/NMB); each rank takes its slice
    per_mb = N // NMB
    per_rank = per_mb // world
    idx = torch.cat([torch.arange(b * per_mb + rank * per_rank, b * per_mb + (rank + 1) * per_rank) for b in range(NMB)])

    spec = ModelSpec(64, 8, [256, 256])

    def run(n_envs, ocfg, sel, data_parallel, graph=False):
        model = PolicyModel(spec, dev)
        model.load_state_dict(st0)
        traj_dev = alloc_trajectory_tensors(64, 8, n_envs, T, dev)
        learner = Learner(make_cfg(ocfg, learner_cuda_graph=graph), model, n_envs, engine=engine, data_parallel=data_parallel)
        assert learner.use_graph == graph
        logs = []
        for b in batches:
            for k, v in b.items():
                traj_dev[k].copy_((v if sel is None else v[sel]).view(traj_dev[k].shape))
            learner.train(traj_dev)
            logs.append(learner.minibatch_log().numpy().copy())
        torch.cuda.synchronize()
        return model, learner, logs

    model_dp, learner_dp, logs_dp = run(N // world, ocfg_loc, idx, True)
    assert learner_dp.world_size == world
    # replicas must be bit-identical across ranks
    flat = model_dp.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref), "replicas diverged"
    stats = torch.cat([model_dp.obs_mean, model_dp.obs_var, model_dp.ret_mean, model_dp.ret_var])
    ref_s = stats.clone()
    dist.broadcast(ref_s, src=0)
    assert torch.equal(stats, ref_s)

    # the same data-parallel training with Learner.train() captured as ONE CUDA graph (kernels + NCCL all-reduces): call 1
    # runs eagerly, call 2 captures and replays, call 3 replays -- same kernels, so the replicas must match the eager run
    if dp_graph:
        model_g, learner_g, logs_g = run(N // world, ocfg_loc, idx, True, graph=True)
        assert learner_g._graph is not None and learner_g.graph_replay_launches > 0
        np.testing.assert_allclose(model_g.flat.cpu().numpy(), model_dp.flat.cpu().numpy(), atol=1e-7, rtol=0)
        assert torch.equal(torch.cat([model_g.obs_mean, model_g.obs_var, model_g.ret_mean, model_g.ret_var]), stats)
        for lg, ld in zip(logs_g, logs_dp):
            np.testing.assert_allclose(lg, ld, rtol=1e-6, atol=1e-7)
        assert learner_g.train_step == learner_dp.train_step and learner_g.env_steps == learner_dp.env_steps
        if rank == 0:
            print("DP_GRAPH_OK")

    if rank == 0:
        model_1, learner_1, logs_1 = run(N, ocfg_full, None, False)
        assert learner_1.world_size == 1
        np.testing.assert_allclose(model_dp.flat.cpu().numpy(), model_1.flat.cpu().numpy(), atol=2e-6)
        for a, b in [(model_dp.obs_mean, model_1.obs_mean), (model_dp.obs_var, model_1.obs_var),
                     (model_dp.ret_mean, model_1.ret_mean), (model_dp.ret_var, model_1.ret_var)]:
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-6, atol=1e-7)
        for l_dp, l_1 in zip(logs_dp, logs_1):
            for key in ["num_valid", "adv_mean", "adv_std"]:
                np.testing.assert_allclose(l_dp[:, ops.LS[key]], l_1[:, ops.LS[key]], rtol=1e-6, atol=1e-6, err_msg=key)
        assert learner_dp.env_steps == learner_1.env_steps
        print("DP_EQUIVALENCE_OK world", world, "engine", engine)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
