#!/usr/bin/env python
"""clock64 timeline of attention CTA (0,0,0): MMA/TMA warp and softmax warp 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(128, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_ATTN_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import ops
B, T, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 248, 4
qkv = torch.randn(B, T, 3, H, 64, device="cuda").bfloat16()
out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
mask = torch.ones(B, 1, T, dtype=torch.bool, device="cuda")
for _ in range(3):
    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, mask=mask, scale=0.125)
torch.cuda.synchronize()
t = trace.cpu()
t0 = int(t[0])
r = lambda i: int(t[i]) - t0 if int(t[i]) else -1
print(f"MMA warp: TMA issued 0, q_full {r(1)}")
n_kv = (T + 127) // 128
for j in range(min(n_kv, 6)):
    b = 8 + j * 8
    print(f"  tile {j}: k_full {r(b)}  S issued {r(b+1)}  s_full {r(b+2)}  p_ready {r(b+3)}  v_full {r(b+4)}  PV issued {r(b+5)}  o_full {r(b+6)}")
for j in range(min(n_kv, 6)):
    b = 64 + j * 8
    print(f"  softmax tile {j}: vis done {r(b)}  s_full {r(b+1)}  pass1 {r(b+2)}  P stored+arrive {r(b+3)}  o_full {r(b+4)}  O acc {r(b+5)}")
