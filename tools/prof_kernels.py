#!/usr/bin/env python
"""Tiny driver for `ncu --set full`: launches each hot kernel a few times on the C2 shapes."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformer_pytorch_lightning_b200 import _native as N, ops

M, d, F = 15872, 256, 2048
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
torch.manual_seed(0)
if which == "gemm":
    y = torch.randn(M, d, device=dev).bfloat16()
    w1 = (torch.randn(F, d, device=dev) / 16).bfloat16(); b1 = torch.randn(F, device=dev)
    w2 = (torch.randn(d, F, device=dev) / 45).bfloat16(); b2 = torch.randn(d, device=dev)
    h = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    x = torch.randn(M, d, device=dev)
    wo = (torch.randn(d, d, device=dev) / 16).bfloat16()
    for _ in range(3):
        ops.gemm(y, w1, b1, h, N.EPI_BIAS_SILU)
        ops.gemm(h, w2, b2, x, N.EPI_RESIDUAL, residual=x, alpha=0.5)
        ops.gemm(y, wo, b2, x, N.EPI_RESIDUAL, residual=x, alpha=1.0)
elif which == "ffn":
    y = torch.randn(M, d, device=dev).bfloat16()
    w1 = (torch.randn(F, d, device=dev) / 16).bfloat16(); b1 = torch.randn(F, device=dev)
    w2 = (torch.randn(d, F, device=dev) / 45).bfloat16(); b2 = torch.randn(d, device=dev)
    g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
    x = torch.randn(M, d, device=dev)
    for _ in range(3):
        ops.ffn(y, w1, b1, w2, b2, x, alpha=0.5, ln={"y": y, "g1": g, "b1": b}, engine=N.ENGINE_TC)
elif which == "attn":
    B, T, H = 64, 248, 4
    qkv = torch.randn(B, T, 3, H, 64, device=dev).bfloat16()
    out = torch.empty(B, T, H * 64, device=dev, dtype=torch.bfloat16)
    mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)
    for _ in range(3):
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, mask=mask, scale=0.125)
elif which == "gemm512":
    # Conformer-L (C3) shapes: the feed-forward GEMMs at d = 512
    Ml, dl = 15936, 512
    y = torch.randn(Ml, dl, device=dev).bfloat16()
    w1 = (torch.randn(F, dl, device=dev) / 22).bfloat16(); b1 = torch.randn(F, device=dev)
    w2 = (torch.randn(dl, F, device=dev) / 45).bfloat16(); b2 = torch.randn(dl, device=dev)
    h = torch.empty(Ml, F, device=dev, dtype=torch.bfloat16)
    x = torch.randn(Ml, dl, device=dev)
    for _ in range(3):
        ops.gemm(y, w1, b1, h, N.EPI_BIAS_SILU)
        ops.gemm(h, w2, b2, x, N.EPI_RESIDUAL, residual=x, alpha=0.5)
elif which == "attn_long":
    B, T, H = 16, 1498, 4
    qkv = torch.randn(B, T, 3, H, 64, device=dev).bfloat16()
    out = torch.empty(B, T, H * 64, device=dev, dtype=torch.bfloat16)
    lens = torch.randint(T // 2, T + 1, (B,), device=dev)
    mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).unsqueeze(1)
    for _ in range(3):
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, mask=mask, scale=0.125)
elif which == "dwconv":
    B, T = 64, 248
    x = torch.randn(B, T, d, device=dev).bfloat16(); y = torch.empty_like(x)
    w = torch.randn(15, d, device=dev); b = torch.randn(d, device=dev)
    for _ in range(3):
        ops.dwconv(x, w, b, y)
elif which == "frontend":
    # sub-sampling front-end (subsample_fused_kernel + the Linear GEMM) on the C2 batch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from _util import build_encoder
    from oracle import conformer_oracle as O
    enc = build_encoder(O.conformer_cfg("M", encoder_num_layers=1), 0, compute_dtype=torch.bfloat16)
    feats = torch.randn(64, 998, 80, device=dev)
    pad = torch.ones(64, 1, 998, dtype=torch.bool, device=dev)
    with torch.no_grad():
        for _ in range(3):
            enc.embed(feats, pad)
elif which == "ctc":
    # CTC greedy head: ctc_lo GEMM with the argmax epilogue (gemm_tc_kernel<256, 100>), V = 5002
    x = torch.randn(M, d, device=dev).bfloat16()
    w = (torch.randn(5002, d, device=dev) / 16).bfloat16(); b = torch.randn(5002, device=dev)
    for _ in range(3):
        ops.ctc_argmax(x, w, b)
torch.cuda.synchronize()
print("done")
