"""ncu target: the fused two-layer policy step (csrc/policy_step.cu) at the cfg-2 sampler shape, next to the per-layer launches.
   ncu --set full -k regex:"policy_mlp2|gemm_tc_ta" -s 6 -c 6 python tools/ncu_policy_step.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sample_factory_b200 import ops

dev = torch.device("cuda", 0)
ops.bind_device(dev)
eng = ops.ENGINES["3xtf32"]
M, K1, H1, H2, A = 4096, 64, 512, 512, 8
flat = torch.empty(H1 * K1 + H2 * H1, device=dev)
lo = torch.empty_like(flat)
flat[: H1 * K1] = (torch.randn(H1, K1) / math.sqrt(K1)).reshape(-1).to(dev)
flat[H1 * K1:] = (torch.randn(H2, H1) / math.sqrt(H1)).reshape(-1).to(dev)
ops.register_tf32_lo(flat, lo)
ops.refresh_tf32_lo(flat)
W1, W2 = flat[: H1 * K1].view(H1, K1), flat[H1 * K1:].view(H2, H1)
b1, b2 = torch.zeros(H1, device=dev), torch.zeros(H2, device=dev)
Wv, Wa = torch.randn(1, H2, device=dev), torch.randn(A, H2, device=dev)
x = torch.randn(M, K1, device=dev)
h1 = torch.empty(M, H1, device=dev)
P = ops.policy_mlp2_partials(W1, W2, A, eng)
part = torch.empty(P * M * ops.HEAD_PART_PAD, device=dev)
act = ops.ACT["elu"]
reps = int(os.environ.get("REPS", "6"))
for _ in range(reps):
    ops.policy_mlp2_heads_forward(x, W1, b1, W2, b2, act, eng, Wv, Wa, part)
torch.cuda.synchronize()
for _ in range(reps):
    ops.linear_act_forward(x, W1, b1, h1, act, eng)
    ops.linear_act_heads_forward(h1, W2, b2, None, act, eng, Wv, Wa, part)
torch.cuda.synchronize()
# event timing outside ncu (REPS=200 python tools/ncu_policy_step.py)
if reps >= 100:
    def t(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    def two():
        ops.linear_act_forward(x, W1, b1, h1, act, eng)
        ops.linear_act_heads_forward(h1, W2, b2, None, act, eng, Wv, Wa, part)
    print("fused us", t(lambda: ops.policy_mlp2_heads_forward(x, W1, b1, W2, b2, act, eng, Wv, Wa, part)), "per-layer us", t(two))
ops.unregister_tf32_lo(flat)
