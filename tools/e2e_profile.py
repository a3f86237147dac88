#!/usr/bin/env python
"""GPU timeline of the pipelined end-to-end path: per-kernel device time per batch and idle time of the compute stream."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from _util import build_encoder
from oracle import conformer_oracle as O
from conformer_pytorch_lightning_b200 import EncoderPipeline
import bench
cfg_name, feats_np, lens_np, T, audio = bench.make_inputs("C2")
cfg = O.conformer_cfg(cfg_name)
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16)
feats = torch.from_numpy(feats_np).pin_memory(); lens = torch.from_numpy(lens_np)
pipe = EncoderPipeline(enc, depth=2)
outs = [torch.empty((64, T, 256)).pin_memory() for _ in range(4)]
for _ in pipe.stream(((feats, lens) for _ in range(4)), outs):
    pass
torch.cuda.synchronize()
n = 8
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in pipe.stream(((feats, lens) for _ in range(n)), outs):
        pass
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
agg = collections.defaultdict(float); cnt = collections.Counter()
for e in ev:
    agg[e.name[:70]] += e.device_time; cnt[e.name[:70]] += 1
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print(f"wall per batch {(t1 - t0) / n:8.1f} us")
tot = 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:16]:
    print(f"{v / n:9.1f} us/batch {cnt[k] / n:6.1f}x  {k}")
    tot += v / n
print(f"sum of the listed kernels {tot:8.1f} us/batch (copies overlap compute)")
