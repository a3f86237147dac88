#!/usr/bin/env python
"""Where the HOST time of a C5 training step goes (cProfile over a few steady-state steps)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
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
enc = build_encoder(cfg, 0, device=dev, compute_dtype=torch.bfloat16).train()
dec = C.CTCDecoder(5002, 256, 0.0).to(dev); dec.compute_dtype = torch.bfloat16
ps = list(enc.parameters()) + list(dec.parameters())
opt = C.FlatAdam(ps, lr=1e-4) if os.environ.get("OPT", "flat") == "flat" else torch.optim.Adam(ps, lr=1e-4, fused=True)
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
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
