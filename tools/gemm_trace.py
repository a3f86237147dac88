#!/usr/bin/env python
"""Cycle accounting of CTA 0 of gemm_tc_kernel (activation epilogues): where the MMA warp and the epilogue wait.
CFM_B200_GEMM_PAIR=0/1 selects single-CTA / CTA-pair tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["CFM_B200_GEMM_TRACE_PTR"] = str(trace.data_ptr())
from conformer_pytorch_lightning_b200 import _native as N, ops
for label, M, Nn, K, epi in (("L w_1+SiLU", 15936, 2048, 512, N.EPI_BIAS_SILU), ("L qkv", 15936, 1536, 512, N.EPI_BIAS),
                             ("M w_1+SiLU", 15872, 2048, 256, N.EPI_BIAS_SILU), ("K=2048 N=512 bias", 15936, 512, 2048, N.EPI_BIAS)):
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(Nn, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(Nn, device="cuda"); c = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, b, c, epi)
    trace.zero_()
    ops.gemm(a, w, b, c, epi)
    torch.cuda.synchronize()
    t = trace.cpu().tolist()
    tiles = max(t[3], 1)
    print(f"{label:18s} M={M} N={Nn} K={K}: MMA warp {t[0]} cycles for {t[3]} tiles ({t[0] // tiles} per tile; ideal {K // 16 * 128}): waiting for a drained "
          f"accumulator {t[1]}, for operands (TMA) {t[2]} | epilogue thread: waiting for the accumulator {t[4]}, for the store ring {t[5]}, busy {t[6]}")
