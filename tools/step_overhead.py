#!/usr/bin/env python
"""How much of a measured-path step is the captured graph, and how much the eager plumbing around it?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import build_encoder
from oracle import conformer_oracle as O
import bench
cfg_name, feats_np, lens_np, T, audio = bench.make_inputs("C2")
cfg = O.conformer_cfg(cfg_name)
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16)
feats = torch.from_numpy(feats_np).cuda(); lens = torch.from_numpy(lens_np).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=30, do_flush=True):
    tot = 0.0
    for _ in range(n):
        if do_flush:
            flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1e3


with torch.no_grad():
    pad = ~bench.enc_make_pad(lens, feats.size(1))
    x, pos, pad = enc.embed(feats, pad)
    from conformer_pytorch_lightning_b200.utils import make_attn_mask
    attn = make_attn_mask(x, pad, False, False, 0, -1, -1)
    step = lambda: enc.encode_layers(x, attn, pos, pad)
    for _ in range(4):
        step()
    plan = [p for p in enc._plans.values() if p.get("graph") is not None][0]
    g = plan["graph"]
    print(f"full step (fill + graph + clone), L2 flushed : {timed(step):8.1f} us")
    print(f"full step, no flush                         : {timed(step, do_flush=False):8.1f} us")
    print(f"graph replay only, L2 flushed               : {timed(g.replay):8.1f} us")
    print(f"graph replay only, no flush                 : {timed(g.replay, do_flush=False):8.1f} us")
    def ten():
        for _ in range(10):
            g.replay()
    print(f"10 replays back to back / 10                : {timed(ten, n=5, do_flush=False) / 10:8.1f} us")
