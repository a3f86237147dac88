"""Times cfm_gemm on the GEMM shapes of the BASELINE configs (CUDA events, L2 flushed between launches).
Run once per CFM_B200_GEMM_PAIR setting (the switch is read once per process):
    CFM_B200_GEMM_PAIR=0 python tools/gemm_shapes_bench.py ; CFM_B200_GEMM_PAIR=1 python tools/gemm_shapes_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conformer_pytorch_lightning_b200 import _native as N, ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [  # (label, M, N, K, epilogue)
    ("L w_1+SiLU", 15936, 2048, 512, N.EPI_BIAS_SILU),
    ("L w_2+resid", 15936, 512, 2048, N.EPI_RESIDUAL),
    ("L qkv", 15936, 1536, 512, N.EPI_BIAS),
    ("L out+resid", 15936, 512, 512, N.EPI_RESIDUAL),
    ("L pw1+GLU", 15936, 512, 512, N.EPI_BIAS_GLU),
    ("M w_1+SiLU", 15872, 2048, 256, N.EPI_BIAS_SILU),
    ("M w_2+resid", 15872, 256, 2048, N.EPI_RESIDUAL),
    ("M out+resid", 15872, 256, 256, N.EPI_RESIDUAL),
    ("C4 out+resid", 23968, 256, 256, N.EPI_RESIDUAL),
]
print("CFM_B200_GEMM_PAIR =", os.environ.get("CFM_B200_GEMM_PAIR", "(auto)"))
for label, M, Nn, K, epi in SHAPES:
    a = torch.randn(M, K, device=dev).bfloat16()
    wrows = 2 * Nn if epi == N.EPI_BIAS_GLU else Nn
    w = (torch.randn(wrows, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(wrows, device=dev)
    if epi == N.EPI_RESIDUAL:
        c = torch.randn(M, Nn, device=dev)
        run = lambda: ops.gemm(a, w, b, c, epi, residual=c, alpha=0.5)
    else:
        c = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
        run = lambda: ops.gemm(a, w, b, c, epi)
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    fl = 2.0 * M * wrows * K
    print(f"{label:14s} M={M} N={Nn} K={K}: {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s  pair launches {N.kernel_launches('gemm_tc_pair')}")
