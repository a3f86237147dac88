#!/usr/bin/env python
"""Timeline of one fused-FFN tile (CTA 0): clock64 stamps written by the MMA warp and one SiLU thread."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(16 * 64 + 4 * 160, dtype=torch.int64, device="cuda")
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
t = trace.cpu()[:12 * 64].view(12, 64)
t0 = int(t[0, 0])
NC = F // 128
print("chunk | MMA warp: G1 start, after s_empty wait | G2 start, after h_full[0] wait, before/after h_full[1] wait |"
      " SiLU thread: begin-wait s_full, got s_full, first half computed, arrived h_full[1]   [cycles since G1(0)]")
for c in range(NC):
    r = lambda i: int(t[i, c]) - t0
    g = lambda i: int(t[i, c // 2]) - t0          # G1 stamps are per chunk PAIR
    print(f"  {c:2d} | {g(0):6d} {g(1):6d} | {r(6):6d} {r(7):6d} {r(8):6d} {r(9):6d} | {int(t[2, c - (c & 1)]) - t0:6d} {int(t[3, c - (c & 1)]) - t0:6d} {r(5):6d} {r(4):6d}")
