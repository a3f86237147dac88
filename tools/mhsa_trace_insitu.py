#!/usr/bin/env python
"""Timeline of the fused attention kernel (CTA (0,3), last layer) inside a real encoder step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
trace = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_MHSA_TRACE_PTR"] = str(trace.data_ptr())
from _util import build_encoder
from oracle import conformer_oracle as O
import bench
cfg_name, feats_np, lens_np, T, audio = bench.make_inputs("C2")
cfg = O.conformer_cfg(cfg_name)
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16)
enc.use_cuda_graphs = False
feats = torch.from_numpy(feats_np).cuda(); lens = torch.from_numpy(lens_np).cuda()
with torch.no_grad():
    pad = ~bench.enc_make_pad(lens, feats.size(1))
    x, pos, pad = enc.embed(feats, pad)
    from conformer_pytorch_lightning_b200.utils import make_attn_mask
    attn = make_attn_mask(x, pad, False, False, 0, -1, -1)
    print("attn mask", tuple(attn.shape), attn.dtype, "pad", tuple(pad.shape))
    for _ in range(3):
        enc.encode_layers(x, attn, pos, pad)
    torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[8]
for h in range(4):
    print(f"head {h}: wait S {t[8 + 4 * h] - t0:6d}  S ready {t[9 + 4 * h] - t0:6d}  max done / P tile free {t[10 + 4 * h] - t0:6d}  P written {t[11 + 4 * h] - t0:6d}")
print(f"PV_3 done {t[24] - t0:6d}  ctx written {t[25] - t0:6d}  projection done {t[26] - t0:6d}  residual/LN epilogue done {t[27] - t0:6d}")
e = [t[32 + i] - t[26] for i in range(13)]
print("epilogue (cycles after projection done): start %d | chunk begin/computed: %s | pass 1 done %d | final pass starts %d | final pass done %d | drained %d"
      % (e[0], " ".join(f"{e[1 + 2 * c]}/{e[2 + 2 * c]}" for c in range(4)), e[9], e[10], e[11], e[12]))
