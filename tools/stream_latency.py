#!/usr/bin/env python
"""Per-chunk latency of the streaming path (forward_chunk, B = 1, 16-frame chunks): scope row f3."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import build_encoder
from oracle import conformer_oracle as O
cfg = O.conformer_cfg("M")
for dt in (torch.bfloat16, torch.float32):
    enc = build_encoder(cfg, 0, compute_dtype=dt)
    feats = torch.randn(1, 998, 80, device="cuda")
    with torch.no_grad():
        for left in (4, -1):
            for _ in range(2):
                out, _ = enc.forward_chunk_by_chunk(feats, 16, left)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out, _ = enc.forward_chunk_by_chunk(feats, 16, left)
            torch.cuda.synchronize()
            dtm = time.perf_counter() - t0
            n_chunks = (998 - 7) // 64 + 1
            print(f"{str(dt):15s} left_chunks={left:2d}: {out.shape[1]} frames in {n_chunks} chunks, {1e3 * dtm / n_chunks:6.2f} ms per chunk "
                  f"(0.64 s of audio each)")

# where does a steady-state step go?  (bf16, 4 left chunks)
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16)
feats = torch.randn(1, 998, 80, device="cuda")
with torch.no_grad():
    for _ in range(3):
        enc.forward_chunk_by_chunk(feats, 16, 4)
    chunk = feats[:, 640:707]
    cache = torch.zeros(12, 4, 64, 128, device="cuda")
    cnn = torch.zeros(0, 0, 0, 0, device="cuda")
    def t(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / n
    mask = torch.ones(1, 1, 67, dtype=torch.bool, device="cuda")
    print(f"forward_chunk (steady state)      : {t(lambda: enc.forward_chunk(chunk, 160, 64, cache, cnn)):6.2f} ms")
    print(f"  embed (sub-sampling + pos-enc)  : {t(lambda: enc.embed(chunk, mask, 160)):6.2f} ms")
    x, pos, _ = enc.embed(chunk, mask, 160)
    pe = enc.embed.position_encoding(offset=160 - 64, size=80)
    print(f"  layer loop (graph replay path)  : {t(lambda: enc._graph_chunk(x, pe, cache, 16, torch.bfloat16)):6.2f} ms")
    print(f"  layer loop (eager)              : {t(lambda: enc._chunk_layers(x.float(), pe, cache, 16, torch.bfloat16)):6.2f} ms")
    from torch.profiler import profile, ProfilerActivity
    import collections
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            enc._graph_chunk(x, pe, cache, 16, torch.bfloat16)
        torch.cuda.synchronize()
    agg = collections.defaultdict(list)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0:
            agg[e.name[:80]].append(e.device_time)
    rows = sorted(((sum(v) / 3, len(v) // 3, k) for k, v in agg.items()), reverse=True)
    print("kernels of one steady-state step (CUPTI):")
    for tot, n, k in rows[:14]:
        print(f"  {tot:8.1f} us {n:4d}x  {k}")
    print(f"  sum {sum(r[0] for r in rows):8.1f} us in {sum(r[1] for r in rows)} launches")
