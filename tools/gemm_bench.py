"""Micro-benchmark of the GEMM engines (CUDA events, L2 flushed between launches). GPU box only."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sample_factory_b200 import ops

dev = torch.device("cuda", 0)
ops.bind_device(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for (M, N, K) in [(4096, 512, 64), (4096, 512, 512), (32768, 512, 64), (32768, 512, 512)]:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / math.sqrt(K); b = torch.zeros(N, device=dev)
    y = torch.empty(M, N, device=dev)
    for eng in ("simt", "tf32", "3xtf32"):
        ms = timeit(lambda: ops.linear_act_forward(x, W, b, y, ops.ACT["elu"], ops.ENGINES[eng]))
        print(f"fwd  M={M:6d} N={N:4d} K={K:4d} {eng:7s} {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
    dz = torch.randn(M, N, device=dev); dW = torch.empty(N, K, device=dev); dx = torch.empty(M, K, device=dev)
    ws = torch.empty(ops.linear_backward_workspace_bytes(M, N, K) // 4 + 4, device=dev)
    for eng in ("simt", "tf32", "3xtf32"):
        ms = timeit(lambda: ops.linear_backward(dz, x, W, ops.ACT["elu"], dW, dx, None, ops.ENGINES[eng], ws))
        print(f"bwd  M={M:6d} N={N:4d} K={K:4d} {eng:7s} {ms*1e3:8.1f} us  {4*M*N*K/ms/1e9:8.1f} TFLOP/s (dW+dx)")
h = torch.randn(4096, 512, device=dev)
Wv = torch.randn(1, 512, device=dev); bv = torch.zeros(1, device=dev); Wa = torch.randn(8, 512, device=dev); ba = torch.zeros(8, device=dev)
vals = torch.empty(4096, device=dev); lg = torch.empty(4096, 8, device=dev); act = torch.empty(4096, device=dev)
ea = torch.empty(4096, dtype=torch.int32, device=dev); lp = torch.empty(4096, device=dev)
noise = torch.empty(4096, 8, device=dev).exponential_()
print("heads_fwd 4096 values+logits      %.1f us" % (1e3 * timeit(lambda: ops.heads_forward(h, Wv, bv, Wa, ba, vals, 1, lg, 8))))
print("heads_fwd 4096 sample (noise)     %.1f us" % (1e3 * timeit(lambda: ops.heads_forward(h, Wv, bv, Wa, ba, vals, 1, lg, 8, noise, 0, 0, None, act, 1, ea, lp, 1))))
print("heads_fwd 4096 sample (philox)    %.1f us" % (1e3 * timeit(lambda: ops.heads_forward(h, Wv, bv, Wa, ba, vals, 1, lg, 8, None, 1, 0, None, act, 1, ea, lp, 1))))
h2 = torch.randn(32768, 512, device=dev); v2 = torch.empty(32768, device=dev); l2 = torch.empty(32768, 8, device=dev)
print("heads_fwd 32768 values+logits     %.1f us" % (1e3 * timeit(lambda: ops.heads_forward(h2, Wv, bv, Wa, ba, v2, 1, l2, 8))))
