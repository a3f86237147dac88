#!/usr/bin/env python
"""Host and device cost of the optimizer step of the C5 training loop: FlatAdam vs torch.optim.Adam(fused=True)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import conformer_pytorch_lightning_b200 as C
from conformer_pytorch_lightning_b200 import _native
from oracle import conformer_oracle as O
from _util import build_encoder
dev = torch.device("cuda", 0)
cfg = O.conformer_cfg("M", static_chunk_size=16, dropout=0.1, attention_dropout=0.1, pos_enc_dropout=0.1)
rs = np.random.RandomState(0)
feats = torch.from_numpy(rs.standard_normal((16, 998, 80)).astype(np.float32)).to(dev)
lens = torch.full((16,), 998, dtype=torch.int32, device=dev)
labels = torch.from_numpy(rs.randint(1, 5000, size=(16, 40)).astype(np.int64)).to(dev)
lab_len = torch.full((16,), 40, dtype=torch.int64, device=dev)
for name in ("torch", "flat", "torch", "flat"):
    enc = build_encoder(cfg, 0, device=dev, compute_dtype=torch.bfloat16).train()
    dec = C.CTCDecoder(5002, 256, 0.0).to(dev); dec.compute_dtype = torch.bfloat16
    ps = list(enc.parameters()) + list(dec.parameters())
    opt = C.FlatAdam(ps, lr=1e-4) if name == "flat" else torch.optim.Adam(ps, lr=1e-4, fused=True)
    host, devt, whole, phases = [], [], [], []
    for it in range(12):
        torch.cuda.synchronize(); t_step = time.perf_counter()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, mask = enc(feats, lens)
        ev[1].record()
        loss = dec(out.float(), mask.squeeze(1).sum(1), labels, lab_len)
        ev[2].record()
        loss.backward()
        ev[3].record()
        t_q = time.perf_counter()                  # host is done enqueueing fwd + bwd
        torch.cuda.synchronize()
        t_sync = time.perf_counter()
        l0 = _native.kernel_launches("adam")
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); s.record(); opt.step(); e.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        if it >= 4:
            phases.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
            host.append(1e3 * (t1 - t0)); devt.append(s.elapsed_time(e)); whole.append((1e3 * (t_q - t_step), 1e3 * (t_sync - t_step)))
    print(f"{name:6s}: opt.step host {np.median(host):.2f} ms, device {np.median(devt):.2f} ms, adam launches {_native.kernel_launches('adam') - l0}; "
          f"device ms fwd / loss / bwd {np.median(np.array(phases), axis=0).round(2).tolist()}; "
          f"fwd+bwd host enqueue {np.median([w[0] for w in whole]):.2f} ms of {np.median([w[1] for w in whole]):.2f} ms device")
