#!/usr/bin/env python
"""Split-K sweep of the training GEMMs (cfm_gemm_ex) at the C5 shard: weight gradients C(N_out, K_in) += dY^T A over 3968
tokens and input gradients dX = dY W."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from conformer_pytorch_lightning_b200 import _native as N, ops
dev = "cuda"
n = 3968
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
print("weight gradients (A_MN, B_MN):  N_out x K_in, splits -> us")
for nout, kin in ((2048, 256), (256, 2048), (768, 256), (256, 256), (512, 256)):
    dy = torch.randn(n, nout, device=dev).bfloat16(); a = torch.randn(n, kin, device=dev).bfloat16()
    gw = torch.zeros(nout, kin, device=dev)
    row = []
    for sp in (0, 1, 2, 3, 4, 6, 9, 12):
        try:
            row.append((sp, timeit(lambda: ops.gemm_ex(dy.t(), a.t(), gw, accumulate=True, splits=sp))))
        except Exception as ex:
            row.append((sp, str(ex)[:20]))
    print(f"  {nout:5d} x {kin:5d}: " + "  ".join(f"{sp}:{t if isinstance(t, str) else round(t, 1)}" for sp, t in row))
print("input gradients dX(n, K_in) = dY(n, N_out) W(N_out, K_in):")
for nout, kin in ((2048, 256), (256, 2048), (768, 256), (256, 256)):
    dy = torch.randn(n, nout, device=dev).bfloat16(); w = torch.randn(nout, kin, device=dev).bfloat16()
    dx = torch.empty(n, kin, device=dev, dtype=torch.bfloat16)
    row = []
    for sp in (0, 1, 2, 4):
        try:
            row.append((sp, timeit(lambda: ops.gemm_ex(dy, w.t(), dx, splits=sp))))
        except Exception as ex:
            row.append((sp, str(ex)[:20]))
    print(f"  {nout:5d} -> {kin:5d}: " + "  ".join(f"{sp}:{t if isinstance(t, str) else round(t, 1)}" for sp, t in row))
