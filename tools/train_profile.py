#!/usr/bin/env python
"""Per-kernel device time of one C5 training step (torch.profiler / CUPTI): where the step goes, graph replay or eager."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import collections
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from _util import build_encoder
from oracle import conformer_oracle as O
import conformer_pytorch_lightning_b200 as C
B, tin, V, Lmax = 16, 998, 5002, 40
cfg = O.conformer_cfg("M", static_chunk_size=16, dropout=0.1, attention_dropout=0.1, pos_enc_dropout=0.1)
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16).train()
enc.use_cuda_graphs = os.environ.get("GRAPH", "1") == "1"
dec = C.CTCDecoder(V, 256, 0.0).cuda(); dec.compute_dtype = torch.bfloat16
opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=1e-4, fused=True)
rs = np.random.RandomState(0)
feats = torch.from_numpy(rs.standard_normal((B, tin, 80)).astype(np.float32)).cuda()
lens = torch.full((B,), tin, dtype=torch.int32, device="cuda")
labels = torch.from_numpy(rs.randint(1, V - 1, size=(B, Lmax)).astype(np.int64)).cuda()
lab_len = torch.full((B,), Lmax, dtype=torch.int64, device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    out, mask = enc(feats, lens)
    loss = dec(out.float(), mask.squeeze(1).sum(1), labels, lab_len)
    loss.backward()
    opt.step()
for _ in range(4):
    step()
torch.cuda.synchronize()
NS = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(list)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0:
        agg[e.name[:90]].append(e.device_time)
rows, tot = [], 0
for k, v in agg.items():
    rows.append((sum(v) / NS, len(v) / NS, float(np.mean(v)), k)); tot += sum(v) / NS
for s, n, m, k in sorted(rows, reverse=True)[:40]:
    print(f"{s:9.1f} us/step {n:6.1f}x {m:7.1f} us  {k}")
print("sum of kernel durations per step: %.1f us" % tot)
