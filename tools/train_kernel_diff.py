#!/usr/bin/env python
"""Per-kernel device time of a C5 training step with FlatAdam vs torch.optim.Adam (torch.profiler / CUPTI)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import conformer_pytorch_lightning_b200 as C
from oracle import conformer_oracle as O
from _util import build_encoder
dev = torch.device("cuda", 0)
cfg = O.conformer_cfg("M", static_chunk_size=16, dropout=0.1, attention_dropout=0.1, pos_enc_dropout=0.1)
rs = np.random.RandomState(0)
feats = torch.from_numpy(rs.standard_normal((16, 998, 80)).astype(np.float32)).to(dev)
lens = torch.full((16,), 998, dtype=torch.int32, device=dev)
labels = torch.from_numpy(rs.randint(1, 5000, size=(16, 40)).astype(np.int64)).to(dev)
lab_len = torch.full((16,), 40, dtype=torch.int64, device=dev)
res = {}
for name in ("torch", "flat"):
    enc = build_encoder(cfg, 0, device=dev, compute_dtype=torch.bfloat16).train()
    dec = C.CTCDecoder(5002, 256, 0.0).to(dev); dec.compute_dtype = torch.bfloat16
    ps = list(enc.parameters()) + list(dec.parameters())
    opt = C.FlatAdam(ps, lr=1e-4) if name == "flat" else torch.optim.Adam(ps, lr=1e-4, fused=True)
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, mask = enc(feats, lens)
        loss = dec(out.float(), mask.squeeze(1).sum(1), labels, lab_len)
        loss.backward()
        opt.step()
    for _ in range(6):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    agg = {}
    for e in prof.key_averages():
        agg[e.key] = (e.device_time_total / 3.0, e.count / 3.0)
    res[name] = agg
keys = sorted(set(res["torch"]) | set(res["flat"]), key=lambda k: -abs(res["flat"].get(k, (0, 0))[0] - res["torch"].get(k, (0, 0))[0]))
print("total us/step: torch %.0f  flat %.0f" % (sum(v[0] for v in res["torch"].values()), sum(v[0] for v in res["flat"].values())))
for k in keys[:22]:
    a, b = res["torch"].get(k, (0, 0)), res["flat"].get(k, (0, 0))
    print(f"{b[0] - a[0]:+9.1f} us  torch {a[0]:8.1f} ({a[1]:.0f}x)  flat {b[0]:8.1f} ({b[1]:.0f}x)  {k[:100]}")
print("---- FlatAdam step, kernels by device time")
for k, v in sorted(res["flat"].items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"{v[0]:9.1f} us ({v[1]:4.0f}x, {v[0] / max(v[1], 1):6.1f} us each)  {k[:110]}")
