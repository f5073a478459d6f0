"""ncu target 2: the short-K / small-M GEMMs (learner layer-1 forward, the sampler's two GEMMs) and heads_backward.
   ncu --set full -k regex:"gemm_tc_ta|heads_backward_vec" -s 8 -c 4 python tools/ncu_target2.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sample_factory_b200 import ops

dev = torch.device("cuda", 0)
ops.bind_device(dev)
eng = ops.ENGINES["3xtf32"]
A = 8
flat = torch.randn(512 * 64 + 512 + 512 * 512 + 512, device=dev) / 16
lo = torch.empty_like(flat)
ops.register_tf32_lo(flat, lo)
W1, b1 = flat[: 512 * 64].view(512, 64), flat[512 * 64: 512 * 64 + 512]
o = 512 * 64 + 512
W2, b2 = flat[o: o + 512 * 512].view(512, 512), flat[o + 512 * 512:]
Wv = torch.randn(1, 512, device=dev)
Wa = torch.randn(A, 512, device=dev)
xL = torch.randn(32768, 64, device=dev)
yL = torch.empty(32768, 512, device=dev)
xS = torch.randn(4096, 64, device=dev)
h1S = torch.empty(4096, 512, device=dev)
P = ops.linear_heads_partials(512, A, eng)
partS = torch.empty(P * 4096 * ops.HEAD_PART_PAD, device=dev)
dl = torch.randn(32768, A, device=dev)
dv = torch.randn(32768, device=dev)
dz = torch.empty(32768, 512, device=dev)
g = [torch.empty(512, device=dev), torch.empty(1, device=dev), torch.empty(A, 512, device=dev), torch.empty(A, device=dev),
     torch.empty(512, device=dev)]
hws = torch.empty(ops.heads_backward_workspace_bytes(512, A) // 4 + 4, device=dev)
for _ in range(3):   # 4 profiled kernels per iteration; iterations 0-1 are warm-up (-s 8)
    ops.linear_act_forward(xL, W1, b1, yL, ops.ACT["elu"], eng)                                   # learner L1 fwd
    ops.linear_act_forward(xS, W1, b1, h1S, ops.ACT["elu"], eng)                                  # sampler GEMM 1
    ops.linear_act_heads_forward(h1S, W2, b2, None, ops.ACT["elu"], eng, Wv, Wa, partS)           # sampler GEMM 2 + heads
    ops.heads_backward(yL, Wv, Wa, dl, dv, ops.ACT["elu"], dz, g[0], g[1], g[2], g[3], g[4], hws)  # heads backward
torch.cuda.synchronize()
