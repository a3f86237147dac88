#!/usr/bin/env python
"""In-situ per-kernel device durations of one measured-path step (torch.profiler / CUPTI, eager launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import collections
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from _util import build_encoder
from oracle import conformer_oracle as O
import bench
cfg_name, feats_np, lens_np, T, audio = bench.make_inputs("C2")
cfg = O.conformer_cfg(cfg_name)
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16)
enc.use_cuda_graphs = os.environ.get("GRAPH", "0") == "1"
feats = torch.from_numpy(feats_np).cuda(); lens = torch.from_numpy(lens_np).cuda()
with torch.no_grad():
    pad = ~bench.enc_make_pad(lens, feats.size(1))
    x, pos, pad = enc.embed(feats, pad)
    for _ in range(3):
        enc.encode_layers(x, pad, pos, pad)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            flush.fill_(1)
            torch.cuda._sleep(5_000_000)
            enc.encode_layers(x, pad, pos, pad)
        torch.cuda.synchronize()
agg = collections.defaultdict(list)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0:
        agg[e.name[:70]].append(e.device_time)
tot = 0
rows = []
for k, v in agg.items():
    if "fill" in k.lower() or "sleep" in k.lower() or "spin" in k.lower():
        continue
    rows.append((sum(v) / 3, len(v) // 3, float(np.mean(v)), k)); tot += sum(v) / 3
for s, n, m, k in sorted(rows, reverse=True)[:14]:
    print(f"{s:9.1f} us/step {n:4d}x {m:7.1f} us  {k}")
print("sum of kernel durations per step: %.1f us" % tot)
