"""ncu target 4: the heads' backward kernel at the learner's cfg-2 minibatch size (32768 x 512, Discrete(8), ELU).
   ncu --set full --import-source on -k regex:"heads_backward_(pipe|vec)" -s 1 -c 1 python tools/ncu_target4.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sample_factory_b200 import ops

dev = torch.device("cuda", 0)
ops.bind_device(dev)
B, H, A = 32768, 512, 8
f32 = dict(dtype=torch.float32, device=dev)
h = torch.nn.functional.elu(torch.randn(B, H, **f32))
Wv, Wa = torch.randn(H, **f32) / 22, torch.randn(A, H, **f32) / 22
dlogits, dvalues = torch.randn(B, A, **f32) / B, torch.randn(B, **f32) / B
dz = torch.empty(B, H, **f32)
dWv, dbv, dWa, dba, dbp = torch.empty(H, **f32), torch.empty(1, **f32), torch.empty(A, H, **f32), torch.empty(A, **f32), torch.empty(H, **f32)
ws = torch.empty(ops.heads_backward_workspace_bytes(H, A) // 4 + 4, **f32)
flush = torch.empty(256 * 1024 * 1024 // 4, **f32)
times = []
for it in range(6):
    flush.zero_()                      # evict h / dz from the 126 MB L2 between launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.heads_backward(h, Wv, Wa, dlogits, dvalues, ops.ACT["elu"], dz, dWv, dbv, dWa, dba, dbp, ws)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3)
alg = h.numel() * 4 * 2 + dlogits.numel() * 4 + B * 4
print(f"heads_backward (+reduce) us per call: {[round(t, 1) for t in times]}  algorithmic GB/s (best): {alg / min(times) / 1e3:.0f}")
