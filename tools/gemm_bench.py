#!/usr/bin/env python
"""Micro-benchmark of the GEMM engine on the encoder's shapes vs torch (cuBLAS) on the same operands.
Warm-L2 timings with CUDA events (operands of one call fit in the 126 MB L2); used to steer kernel work."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformer_pytorch_lightning_b200 import _native as N, ops


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 15872
    dev = "cuda"
    shapes = [("ffn_w1+silu", 2048, 256, N.EPI_BIAS_SILU), ("ffn_w2+res", 256, 2048, N.EPI_RESIDUAL),
              ("qkv", 768, 256, N.EPI_BIAS), ("out/pw2+res", 256, 256, N.EPI_RESIDUAL), ("pw1+glu", 256, 256, N.EPI_BIAS_GLU),
              ("L:w1", 2048, 512, N.EPI_BIAS_SILU), ("L:w2", 512, 2048, N.EPI_RESIDUAL), ("L:qkv", 1536, 512, N.EPI_BIAS)]
    for name, Nn, K, epi in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        wrows = 2 * Nn if epi == N.EPI_BIAS_GLU else Nn
        w = (torch.randn(wrows, K, device=dev) / math.sqrt(K)).bfloat16()
        bias = torch.randn(wrows, device=dev)
        if epi == N.EPI_RESIDUAL:
            out = torch.randn(M, Nn, device=dev)
            fn = lambda: ops.gemm(a, w, bias, out, epi, residual=out, alpha=0.5)
        else:
            out = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, bias, out, epi)
        t = timeit(fn)
        tref = timeit(lambda: torch.nn.functional.linear(a, w))
        fl = 2.0 * M * wrows * K
        print(f"{name:14s} M={M} N={wrows} K={K}: ours {t:7.1f} us ({fl / t / 1e6:7.1f} TF/s)   torch.linear(no epilogue) {tref:7.1f} us ({fl / tref / 1e6:7.1f} TF/s)")


if __name__ == "__main__":
    main()
