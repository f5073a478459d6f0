"""ncu target 3: the HBM-bound kernels of the path at cfg-2 sizes (N=4096 envs, T=32, 131072 samples, 300 553 parameters).
   ncu --set full -k regex:"normalize_kernel|gae_returns|post_pre_step|ppo_loss_kernel|clip_adam|heads_from_partials|moments_partial|vtrace" -s <warm-up launches> -c 8 python tools/ncu_target3.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sample_factory_b200 import ops

dev = torch.device("cuda", 0)
ops.bind_device(dev)
N, T, D, A = 4096, 32, 64, 8
E = N * T
f32 = dict(dtype=torch.float32, device=dev)
obs = torch.randn(N * (T + 1), D, **f32)
nobs = torch.empty_like(obs)
mean = torch.zeros(D, dtype=torch.float64, device=dev)
var = torch.ones(D, dtype=torch.float64, device=dev)
bmean, bvar = torch.empty(D, **f32), torch.empty(D, **f32)
mws = torch.empty(ops.moments_workspace_bytes(D) // 4 + 4, **f32)
rewards = torch.randn(N, T, **f32)
dones = torch.rand(N, T, device=dev) < 0.02
touts = torch.zeros(N, T, dtype=torch.bool, device=dev)
values = torch.randn(N, T + 1, **f32)
valids = torch.ones(N, T + 1, dtype=torch.bool, device=dev)
adv, ret = torch.empty(N, T, **f32), torch.empty(N, T, **f32)
rm, rv = torch.zeros(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)
B = E // 4
logits, logits_old = torch.randn(B, A, **f32), torch.randn(B, A, **f32)
vals, v_old = torch.randn(B, **f32), torch.randn(B, **f32)
actions = torch.randint(0, A, (B,), device=dev).float()
lp_old = -torch.rand(B, **f32) * 2
advb, tgt = torch.randn(B, **f32), torch.randn(B, **f32)
vb = torch.ones(B, dtype=torch.bool, device=dev)
dl, dv = torch.empty(B, A, **f32), torch.empty(B, **f32)
stats = torch.zeros(ops.LS_SIZE, dtype=torch.float64, device=dev)
lws = torch.empty(ops.loss_workspace_bytes(B) // 8 + 8, dtype=torch.float64, device=dev)
npar = 300553
p, g, m, v = (torch.randn(npar, **f32) * 0.01 for _ in range(4))
v = v.abs()
aws = torch.empty(1024, **f32)
# sampler step pieces
o = torch.randn(N, D, **f32)
traj_obs = torch.empty(N, T + 1, D, **f32)
traj_rnn = torch.zeros(N, T + 1, 1, **f32)
rnn = torch.zeros(N, 1, **f32)
xn = torch.empty(N, D, **f32)
rew = torch.randn(N, **f32)
tm = torch.zeros(N, dtype=torch.bool, device=dev)
tr = torch.zeros(N, dtype=torch.bool, device=dev)
t_rew, t_done, t_to = torch.empty(N, T, **f32), torch.empty(N, T, dtype=torch.bool, device=dev), torch.empty(N, T, dtype=torch.bool, device=dev)
t_pid = torch.empty(N, T, dtype=torch.int32, device=dev)
ep = [torch.zeros(N, **f32), torch.zeros(N, dtype=torch.int32, device=dev), torch.zeros(N, **f32), torch.zeros(N, **f32)]
est = torch.zeros(8, dtype=torch.float64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
part = torch.randn(8 * N * ops.HEAD_PART_PAD, **f32)
bvv, baa = torch.zeros(1, **f32), torch.zeros(A, **f32)
tv, tl = torch.empty(N, T + 1, **f32), torch.empty(N, T, A, **f32)
ta, tlp, tpv = torch.empty(N, T, 1, **f32), torch.empty(N, T, **f32), torch.empty(N, T, **f32)
ea = torch.empty(N, dtype=torch.int32, device=dev)
pvs = torch.zeros(1, **f32)
for it in range(2):   # iteration 0 = warm-up
    ops.batch_moments(obs, bmean, bvar, mws)
    ops.normalize_obs(obs, nobs, mean, var, 0.0, 1.0)
    ops.gae_returns(rewards, dones, touts, values, valids, 0.99, 0.95, False, rm, rv, adv, ret)
    ops.adv_stats(advb, vb, stats, None, lws)
    ops.ppo_loss_fwd_bwd(logits, vals, actions, lp_old, v_old, advb, tgt, vb, logits_old, 0.1, 1.0, 0.003, 0.5, 0.0, 1.0, dl,
                         dv, stats, lws)
    ops.clip_adam_step(p, g, m, v, 1 + it, 1e-4, 0.9, 0.999, 1e-6, 4.0, None, None, None, aws)
    ops.sampler_post_pre_step(rew, tm, tr, 1.0, 1000.0, 0, t_rew[:, 3], t_done[:, 3], t_to[:, 3], t_pid[:, 3], ep[0], ep[1],
                              ep[2], ep[3], 1, est, cnt, obs=o, traj_obs_next=traj_obs[:, 4], rnn=rnn,
                              traj_rnn_next=traj_rnn[:, 4], x_norm=xn, mean=mean, var=var, sub_mean=0.0, inv_scale=1.0)
    ops.heads_from_partials(part, 8, N, bvv, baa, values=tv[:, 3], values_stride=T + 1, logits=tl[:, 3], logits_stride=T * A,
                            philox_seed=1, philox_offset_dev=cnt, actions_f32=ta[:, 3], actions_stride=T, env_actions=ea,
                            log_prob=tlp[:, 3], log_prob_stride=T, policy_version_scalar=pvs, policy_version_out=tpv[:, 3],
                            pv_stride=T)
torch.cuda.synchronize()
