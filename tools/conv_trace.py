#!/usr/bin/env python
"""Timeline of one fused conv-module tile (CTA 70): clock64 stamps of the MMA warp and epilogue thread 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_CONV_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import _native as N, ops
B, T, d, k = 64, 248, 256, 15
M = B * T
dev = "cuda"
y = torch.randn(M, d, device=dev).bfloat16(); y2 = torch.empty_like(y)
w1 = (torch.randn(2 * d, d, device=dev) / 16).bfloat16(); b1 = torch.randn(2 * d, device=dev)
w2 = (torch.randn(d, d, device=dev) / 16).bfloat16(); b2 = torch.randn(d, device=dev)
dw = torch.randn(k, d, device=dev) * 0.3; db = torch.randn(d, device=dev)
g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
x = torch.randn(M, d, device=dev)
for _ in range(3):
    ops.conv_module(y, w1, b1, dw, db, w2, b2, x, B, T, ln={"y": y2, "g1": g, "b1": b}, engine=N.ENGINE_TC)
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[0]
names = {0: "mma: start", 1: "mma: y tile landed", 2: "mma: pw1 issued", 3: "mma: C tile ready", 4: "mma: pw2 issued",
         8: "epi: start", 9: "epi: pw1 accumulators complete", 10: "epi: GLU written (this thread)", 11: "epi: GLU barrier",
         12: "epi: depthwise done (this thread)", 13: "epi: pw2 accumulator complete", 14: "epi: residual/LN epilogue done"}
for i in sorted(names, key=lambda i: t[i]):
    print(f"{t[i] - t0:8d} clk  {names[i]}")
