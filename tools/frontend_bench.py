#!/usr/bin/env python
"""Device-time breakdown of the native sub-sampling front-end (conv1+conv2 / linear) via CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tools.gemm_bench import timeit
from conformer_pytorch_lightning_b200 import ops, _native as N
B, Tin, idim, C = 64, 998, 80, 256
x = torch.randn(B, Tin, idim, device="cuda")
w1 = torch.randn(C, 9, device="cuda") * 0.3; b1 = torch.randn(C, device="cuda") * 0.1
w2 = (torch.randn(C, 9 * C, device="cuda") / 48).bfloat16(); b2 = torch.randn(C, device="cuda") * 0.1
T2, F2 = 248, 19
ws = torch.empty(ops.subsample_ws_bytes(B, Tin, idim, C), dtype=torch.uint8, device="cuda")
act = torch.empty(B * T2, F2 * C, dtype=torch.bfloat16, device="cuda")
t = timeit(lambda: ops.subsample_conv(x, w1, b1, w2, b2, ws, act), iters=5, per_graph=4)
fl = 2.0 * B * T2 * F2 * C * 9 * C
print(f"conv1+conv2: {t:8.1f} us  (conv2 = {fl / 1e9:.0f} GFLOP -> {fl / t / 1e6:6.1f} TF/s if conv1 were free)")
w3 = (torch.randn(C, F2 * C, device="cuda") / 70).bfloat16(); b3 = torch.randn(C, device="cuda")
out = torch.zeros(B * T2, C, device="cuda")
t = timeit(lambda: ops.gemm(act, w3, b3, out, N.EPI_RESIDUAL, residual=out, alpha=1.0), iters=5, per_graph=4)
print(f"linear K=4864: {t:8.1f} us ({2.0 * B * T2 * C * F2 * C / t / 1e6:6.1f} TF/s)")
# per-kernel device times (CUPTI)
from torch.profiler import profile, ProfilerActivity
for _ in range(2):
    ops.subsample_conv(x, w1, b1, w2, b2, ws, act)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.subsample_conv(x, w1, b1, w2, b2, ws, act)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if e.device_time_total > 0:
        print(f"{e.key[:70]:70s} {e.device_time_total / max(e.count, 1):9.1f} us x{e.count}")
