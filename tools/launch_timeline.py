#!/usr/bin/env python
"""Where does the time of one kernel launch go?  %globaltimer stamps of every CTA of mhsa_fused_kernel (entry, prologue
done, griddepcontrol.wait passed, exit) for back-to-back launches inside one CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B, T, H, d = 64, 248, 4, 256
NCTA = 2 * B
gt = torch.zeros(64 * NCTA * 4, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_MHSA_GTRACE_PTR"] = str(gt.data_ptr())
from conformer_pytorch_lightning_b200 import _native as N, ops
dev = "cuda"
qkv = torch.randn(B, T, 3, H, 64, device=dev).bfloat16()
wo = (torch.randn(d, d, device=dev) / 16).bfloat16(); bo = torch.randn(d, device=dev)
g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
x = torch.randn(B * T, d, device=dev); y = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)
fn = lambda: ops.mhsa_out(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], wo, bo, x, mask=mask, scale=0.125,
                          ln={"y": y, "g1": g, "b1": b}, engine=N.ENGINE_TC)
for _ in range(3):
    fn()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    fn()
torch.cuda.current_stream().wait_stream(side)
n = 8
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(n):
        fn()
first_slot = (3 + 1) % 64            # launches so far: 3 warm-up + 1 side
gr.replay(); torch.cuda.synchronize()
gt.zero_()
gr.replay(); torch.cuda.synchronize()
t = gt.cpu().view(64, NCTA, 4).double()
t0 = None
for i in range(n):
    s = t[(first_slot + i) % 64]
    if t0 is None:
        t0 = s[:, 0].min()
    e, p, w, x_ = (s[:, k] - t0 for k in range(4))
    print(f"launch {i}: entry {e.min() / 1e3:7.2f}..{e.max() / 1e3:7.2f} us | prologue done ..{p.max() / 1e3:7.2f} | "
          f"dependency wait passed {w.min() / 1e3:7.2f}..{w.max() / 1e3:7.2f} | exit {x_.min() / 1e3:7.2f}..{x_.max() / 1e3:7.2f}"
          f" | body (wait->exit) mean {(x_ - w).mean() / 1e3:6.2f} us")
