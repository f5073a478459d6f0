"""ncu target: the learner's three layer-2 GEMMs (forward + fused heads, dW, dx) at the cfg-2 minibatch size, a few
launches each after warm-up.   ncu --set full -k regex:gemm_tc_ta -s 6 -c 3 python tools/ncu_target.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sample_factory_b200 import ops

dev = torch.device("cuda", 0)
ops.bind_device(dev)
M, N, K, A = 32768, 512, 512, 8
eng = ops.ENGINES["3xtf32"]
x = torch.randn(M, K, device=dev)
W = torch.randn(N, K, device=dev) / math.sqrt(K)
b = torch.zeros(N, device=dev)
y = torch.empty(M, N, device=dev)
Wv = torch.randn(1, N, device=dev)
Wa = torch.randn(A, N, device=dev)
P = ops.linear_heads_partials(N, A, eng)
part = torch.empty(P * M * ops.HEAD_PART_PAD, device=dev)
dz = torch.randn(M, N, device=dev)
dW = torch.empty(N, K, device=dev)
dx = torch.empty(M, K, device=dev)
ws = torch.empty(ops.linear_backward_workspace_bytes(M, N, K) // 4 + 4, device=dev)
for _ in range(3):   # launches 0..5 warm-up (2 tcgen05 launches per iteration + 1 fused), then the profiled ones
    ops.linear_act_heads_forward(x, W, b, y, ops.ACT["elu"], eng, Wv, Wa, part)
    ops.linear_backward(dz, x, W, ops.ACT["elu"], dW, dx, None, eng, ws)
torch.cuda.synchronize()
