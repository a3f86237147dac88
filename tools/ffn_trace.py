#!/usr/bin/env python
"""Timeline of one fused-FFN tile (CTA 0): clock64 stamps written by the MMA warp and one SiLU thread."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(6 * 64, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_FFN_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import _native as N, ops
M, d, F = 15872, 256, 2048
y = torch.randn(M, d, device="cuda").bfloat16()
w1 = (torch.randn(F, d, device="cuda") / 16).bfloat16(); b1 = torch.randn(F, device="cuda")
w2 = (torch.randn(d, F, device="cuda") / 45).bfloat16(); b2 = torch.randn(d, device="cuda")
x = torch.randn(M, d, device="cuda")
for _ in range(3):
    ops.ffn(y, w1, b1, w2, b2, x, alpha=0.5, engine=N.ENGINE_TC)
torch.cuda.synchronize()
t = trace.cpu().view(6, 64)
t0 = int(t[0, 0])
print("MMA warp: job (G1/G2, chunk): start-wait, issued-after-wait   [cycles since job 0]")
NC = F // 128
def job_of(jx):
    if jx < 2: return "G1", jx
    if jx >= 2 * NC - 2: return "G2", jx - NC
    return ("G1", (jx + 1) // 2) if jx & 1 else ("G2", (jx - 2) // 2)
for jx in range(2 * NC):
    k, c = job_of(jx)
    print(f"  job {jx:2d} {k}({c:2d}): start {int(t[0, jx]) - t0:7d}  waited-until {int(t[1, jx]) - t0:7d}")
print("SiLU thread: chunk: begin-wait s_full, got s_full, begin-wait h_empty, arrived h_full")
for c in range(NC):
    print(f"  chunk {c:2d}: {int(t[2, c]) - t0:7d} {int(t[3, c]) - t0:7d} {int(t[5, c]) - t0:7d} {int(t[4, c]) - t0:7d}")
