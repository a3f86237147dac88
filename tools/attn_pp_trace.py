#!/usr/bin/env python
"""clock64 timeline of attention_pp CTA (0,0,0): MMA issuer and the first warp of each softmax group (C4 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_ATTN_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import ops
B, T, H = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 1498, 4
qkv = torch.randn(B, T, 3, H, 64, device="cuda").bfloat16()
out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
mask = torch.ones(B, 1, T, dtype=torch.bool, device="cuda")
for _ in range(3):
    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, mask=mask, scale=0.125)
torch.cuda.synchronize()
t = trace.cpu()
t0 = int(t[0])
r = lambda i: int(t[i]) - t0 if int(t[i]) else -1
print(f"issuer: start 0, Q/KV landed {r(1)}")
for j in range(6):
    for q in range(2):
        b = 16 + (j * 2 + q) * 2
        print(f"  issuer tile {j} q{q}: p_ready seen {r(b):6d}  PV + next S issued {r(b + 1):6d}")
for g in range(2):
    for j in range(6):
        b = 48 + (g * 6 + j) * 6
        print(f"  group {g} tile {j}: wait S {r(b):6d}  S ready {r(b+1):6d}  max done {r(b+2):6d}  P written {r(b+3):6d}  O ready {r(b+4):6d}  O accumulated {r(b+5):6d}")
