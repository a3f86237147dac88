#!/usr/bin/env python
"""Why is the chained feed-forward launch ~20 % slower inside the layer stack than back to back?  Times the same
ffn_chain launch (C2 shapes) (a) back to back, (b) after an L2 flush, (c) after another large kernel (conv module) that
evicts the instruction cache but leaves the data in L2, (d) after both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformer_pytorch_lightning_b200 import _native as N, ops
M, d, F = 15872, 256, 2048
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
y = rn(M, d).bfloat16(); x = rn(M, d)
def module():
    return dict(w1=rn(F, d, sc=1 / 16).bfloat16(), b1=rn(F), w2=rn(d, F, sc=1 / 45).bfloat16(), b2=rn(d),
                g1=1 + 0.1 * rn(d), be1=0.1 * rn(d), alpha=0.5)
a, b = module(), module()
wp = rn(3 * d, d, sc=1 / 16).bfloat16(); bp = rn(3 * d)
y_out = torch.empty_like(y); P = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# an unrelated big kernel on unrelated data: a plain GEMM chain through gemm_tc (different code, 8 MB of data)
ga = rn(4096, 512).bfloat16(); gw = rn(512, 512, sc=1 / 22).bfloat16(); gb = rn(512); gc = torch.empty(4096, 512, device=dev, dtype=torch.bfloat16)
def other():
    ops.gemm(ga, gw, gb, gc, N.EPI_BIAS_SILU)
def run():
    ops.ffn_chain(y, a, b, x, y_out, proj=(wp, bp, P), engine=N.ENGINE_TC)
def timeit(pre, n=20):
    ts = []
    for _ in range(n):
        pre()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
for _ in range(5):
    run()
print("back to back            : %.1f us" % timeit(lambda: None))
print("after L2 flush          : %.1f us" % timeit(lambda: flush.fill_(1)))
print("after another kernel    : %.1f us" % timeit(other))
print("after flush + kernel    : %.1f us" % timeit(lambda: (flush.fill_(1), other())))
def touch():
    flush.fill_(1)
    for m in (a, b):
        m["w1"].add_(0); m["w2"].add_(0)       # weights back in L2, activations cold
    wp.add_(0)
print("flush, weights re-warmed: %.1f us" % timeit(touch))
def touch_act():
    flush.fill_(1)
    x.add_(0); y.add_(0)                        # activations in L2, weights cold
print("flush, x/y re-warmed    : %.1f us" % timeit(touch_act))
