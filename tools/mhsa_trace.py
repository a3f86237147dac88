#!/usr/bin/env python
"""Timeline of one fused attention + output-projection tile (CTA (0, 3)): clock64 stamps of softmax thread 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_MHSA_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import _native as N, ops
B, T, H, d = 64, 248, 4, 256
dev = "cuda"
qkv = torch.randn(B, T, 3, H, 64, device=dev).bfloat16()
wo = (torch.randn(d, d, device=dev) / 16).bfloat16(); bo = torch.randn(d, device=dev)
g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
x = torch.randn(B * T, d, device=dev); y = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)
for _ in range(3):
    ops.mhsa_out(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], wo, bo, x, mask=mask, scale=0.125, ln={"y": y, "g1": g, "b1": b},
                 engine=N.ENGINE_TC)
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[8]
for h in range(4):
    print(f"head {h}: wait S {t[8 + 4 * h] - t0:6d}  S ready {t[9 + 4 * h] - t0:6d}  max done / P tile free {t[10 + 4 * h] - t0:6d}  P written {t[11 + 4 * h] - t0:6d}")
print(f"PV_3 done {t[24] - t0:6d}  ctx written {t[25] - t0:6d}  projection done {t[26] - t0:6d}  residual/LN epilogue done {t[27] - t0:6d}")
e = [t[32 + i] - t[26] for i in range(13)]
print("epilogue (cycles after projection done): start %d | chunk begin/computed: %s | pass 1 done %d | final pass starts %d | final pass done %d | drained %d"
      % (e[0], " ".join(f"{e[1 + 2 * c]}/{e[2 + 2 * c]}" for c in range(4)), e[9], e[10], e[11], e[12]))
