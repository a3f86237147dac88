import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_train import _train_step, grad_sample
from _util import load_golden, max_rel
g = load_golden("m2_train_grad")
for dtype in (torch.float32, torch.bfloat16):
    enc, dec, out, loss, launches = _train_step(g, dtype)
    print(dtype, "loss", loss.item(), float(g["loss"]), "out err", max_rel(out.detach().cpu().numpy(), g["out"]), "launches", launches)
    rows = []
    for k, p in list(enc.named_parameters()) + [("ctc." + k, p) for k, p in dec.named_parameters()]:
        key = k.replace(".", "__")
        ref_n, ref_s = float(g["gn__" + key]), g["gs__" + key]
        got = p.grad.detach().float().cpu().numpy()
        gs = grad_sample(got)
        rows.append((k, ref_n, float(np.abs(ref_s).max()), float(np.abs(gs - ref_s).max()), float(np.linalg.norm(got.astype(np.float64)))))
    for r in rows:
        rel = r[3] / max(r[2], 1e-30)
        if rel > (1e-4 if dtype == torch.float32 else 2e-2):
            print(f"  {r[0]:55s} refnorm {r[1]:.3e} refmax {r[2]:.3e} abserr {r[3]:.3e} rel {rel:.3e} gotnorm {r[4]:.3e}")
