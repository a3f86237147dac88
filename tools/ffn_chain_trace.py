#!/usr/bin/env python
"""Phase timeline of one chained feed-forward tile (CTA 0 of ffn_fused_kernel: FFN2(l) -> FFN1(l+1) -> QKV projection), clock64
stamps of the MMA warp and of one epilogue thread.  CFM_B200_FFN_PAIR=1 selects the cta_group::2 kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(16 * 64 + 4 * 160, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_FFN_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import _native as N, ops
M, d, F = 15872, 256, 2048
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
y = rn(M, d).bfloat16()
x = rn(M, d)
def module():
    return dict(w1=rn(F, d, sc=1 / 16).bfloat16(), b1=rn(F), w2=rn(d, F, sc=1 / 45).bfloat16(), b2=rn(d),
                g1=1 + 0.1 * rn(d), be1=0.1 * rn(d))
a, b = module(), module()
wp = rn(3 * d, d, sc=1 / 16).bfloat16(); bp = rn(3 * d)
y_out = torch.empty_like(y); P = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
a["alpha"], b["alpha"] = 0.5, 0.5
for _ in range(3):
    ops.ffn_chain(y, a, b, x, y_out, proj=(wp, bp, P), engine=N.ENGINE_TC)
torch.cuda.synchronize()
tr = trace.cpu()
t = tr[:12 * 64].view(12, 64)
t0 = int(t[10, 0])
e = lambda i: int(t[10, i]) - t0
print(f"pair kernel launches: {N.kernel_launches('ffn_fused_pair')}")
print("MMA warp   : input tile landed %d | module 2 input ready (a_ready) %d | last G2 issued: module 1 %d, module 2 %d | "
      "projection input ready %d" % (e(1), e(2), e(3), e(4), e(5)))
print("epilogue   : Y complete: module 1 %d, module 2 %d | epilogue done: module 1 %d, module 2 %d | projection tail done %d"
      % (e(8), e(9), e(10), e(11), e(12)))
g1 = [int(t[1, pr]) - t0 for pr in range(F // 256)]
print("module 2 main loop, G1 issue times per chunk pair:", g1, " period", (g1[-1] - g1[1]) // (len(g1) - 2))
for name, off in (("module 1 epilogue (X, y stay on chip)", 0), ("module 2 epilogue (X stored, y -> projection operand)", 32)):
    r = [int(t[11, off + i]) - t0 for i in range(13)]
    q = [int(t[11, off + i]) - t0 - r[0] for i in range(13, 17)]
    print("   chunk 1 detail: landed", r[3] - r[0], "tmem ld done", q[0], "math+smem done", q[1], "tmem st issued", q[2], "fence done", r[4] - r[0], "barrier passed", q[3])
    print(f"{name}: start {r[0]} | chunks (residual landed, done): " + " ".join(f"({r[1 + 2 * c] - r[0]}, {r[2 + 2 * c] - r[0]})" for c in range(4))
          + f" | pass 1 done {r[9] - r[0]} | final pass {r[10] - r[0]} .. {r[11] - r[0]} | end {r[12] - r[0]}")
import numpy as np
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
durs = []
for _ in range(5):
    flush.fill_(1)
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record(); ops.ffn_chain(y, a, b, x, y_out, proj=(wp, bp, P), engine=N.ENGINE_TC); e_.record(); e_.synchronize()
    durs.append(s_.elapsed_time(e_) * 1e3)
g = trace.cpu()[16 * 64:].view(160, 4).numpy()
g = g[g[:, 0] > 0]
t00 = g[:, 0].min()
print(f"event bracket {sorted(durs)[2]:.1f} us; {len(g)} CTAs (globaltimer, us since the first CTA's entry): entry max {1e-3 * (g[:, 0].max() - t00):.1f} | "
      f"prologue done min {1e-3 * (g[:, 1].min() - t00):.1f} max {1e-3 * (g[:, 1].max() - t00):.1f} | all warps done min {1e-3 * (g[:, 2].min() - t00):.1f} "
      f"median {1e-3 * (np.median(g[:, 2]) - t00):.1f} max {1e-3 * (g[:, 2].max() - t00):.1f}")
print("per-CTA busy time (prologue done -> done), us: min %.1f median %.1f max %.1f" % tuple(1e-3 * np.percentile(g[:, 2] - g[:, 1], [0, 50, 100])))
