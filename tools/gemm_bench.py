#!/usr/bin/env python
"""Micro-benchmark of the GEMM engine on the encoder's shapes vs torch (cuBLAS) on the same operands.
Warm-L2 timings with CUDA events (operands of one call fit in the 126 MB L2); used to steer kernel work."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformer_pytorch_lightning_b200 import _native as N, ops


def timeit(fn, iters=20, per_graph=10):
    """Average device time of one call, measured by replaying a CUDA graph of `per_graph` calls (Python /
    ctypes launch overhead is ~10 us per call and would otherwise hide every kernel shorter than that)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / (iters * per_graph) * 1e3  # us


def small_kernels():
    B, T, d, H = 64, 248, 256, 4
    dev = "cuda"
    x = torch.randn(B, T, d, device=dev).bfloat16(); y = torch.empty_like(x)
    w = torch.randn(15, d, device=dev); b = torch.randn(d, device=dev)
    t = timeit(lambda: ops.dwconv(x, w, b, y))
    print(f"dwconv k=15 (B={B},T={T},d={d}): {t:6.1f} us  -> {2 * x.numel() * 2 / t / 1e3:7.1f} GB/s algorithmic")
    qkv = torch.randn(B, T, 3, H, 64, device=dev).bfloat16()
    out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * B * T * T * d
    for name, m in (("pad mask (B,1,T)", torch.ones(B, 1, T, dtype=torch.bool, device=dev)), ("no mask", None),
                    ("full mask (B,T,T)", torch.ones(B, T, T, dtype=torch.bool, device=dev))):
        t = timeit(lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, mask=m, scale=0.125))
        print(f"attention {name:18s}: {t:6.1f} us  ({fl / t / 1e6:6.1f} TF/s)")
    xr = torch.randn(B * T, d, device=dev); g = torch.ones(d, device=dev); be = torch.zeros(d, device=dev)
    yb = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.layernorm(xr, g, be, y=yb))
    print(f"layernorm fp32->bf16: {t:6.1f} us -> {(xr.numel() * 6) / t / 1e3:7.1f} GB/s")


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


def ffn_bench(M=15872, d=256, F=2048):
    dev = "cuda"
    y = torch.randn(M, d, device=dev).bfloat16()
    w1 = (torch.randn(F, d, device=dev) / 16).bfloat16(); b1 = torch.randn(F, device=dev)
    w2 = (torch.randn(d, F, device=dev) / 45).bfloat16(); b2 = torch.randn(d, device=dev)
    g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
    x = torch.randn(M, d, device=dev)
    h = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * M * d * F
    for name, eng in (("fused", N.ENGINE_TC), ("two GEMMs", N.ENGINE_SIMT + 99)):
        for mode, ln in (("no LN", None), ("LN", {"y": y, "g1": g, "b1": b}), ("2xLN", {"y": y, "g1": g, "b1": b, "g2": g, "b2": b})):
            if eng == N.ENGINE_TC:
                fn = lambda: ops.ffn(y, w1, b1, w2, b2, x, alpha=0.5, ln=ln, hidden_ws=h, engine=N.ENGINE_TC)
            else:
                def fn():
                    ops.gemm(y, w1, b1, h, N.EPI_BIAS_SILU)
                    if ln is None:
                        ops.gemm(h, w2, b2, x, N.EPI_RESIDUAL, residual=x, alpha=0.5)
                    else:
                        ops.gemm_ln(h, w2, b2, x, y, alpha=0.5, g1=ln["g1"], b1=ln["b1"], g2=ln.get("g2"), b2=ln.get("b2"))
            t = timeit(fn)
            print(f"ffn {name:10s} {mode:6s} M={M} F={F}: {t:7.1f} us ({fl / t / 1e6:7.1f} TF/s)")
    a = {"w1": w1, "b1": b1, "w2": w2, "b2": b2, "alpha": 0.5, "g1": g, "be1": b, "g2": g, "be2": b}
    bm = {"w1": w1, "b1": b1, "w2": w2, "b2": b2, "alpha": 0.5, "g1": g, "be1": b}
    t = timeit(lambda: ops.ffn_chain(y, a, bm, x, y, engine=N.ENGINE_TC))
    print(f"ffn chain (2xLN module + LN module in one kernel): {t:7.1f} us ({2 * fl / t / 1e6:7.1f} TF/s)")


def conv_bench(B=64, T=248, d=256, k=15):
    dev = "cuda"
    M = B * T
    y = torch.randn(M, d, device=dev).bfloat16(); y2 = torch.empty_like(y)
    w1 = (torch.randn(2 * d, d, device=dev) / 16).bfloat16(); b1 = torch.randn(2 * d, device=dev)
    w2 = (torch.randn(d, d, device=dev) / 16).bfloat16(); b2 = torch.randn(d, device=dev)
    dw = torch.randn(k, d, device=dev) * 0.3; db = torch.randn(d, device=dev)
    g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
    x = torch.randn(M, d, device=dev)
    rv = torch.ones(M, dtype=torch.uint8, device=dev)
    gws, cws = torch.empty_like(y), torch.empty_like(y)
    fl = 2.0 * M * d * d * 3
    for name, eng in (("fused", N.ENGINE_TC), ("3 kernels", N.ENGINE_SIMT + 99)):
        for mode, ln in (("no LN", None), ("LN", {"y": y2, "g1": g, "b1": b})):
            if eng == N.ENGINE_TC:
                fn = lambda: ops.conv_module(y, w1, b1, dw, db, w2, b2, x, B, T, row_valid=rv, ln=ln, engine=N.ENGINE_TC)
            else:
                def fn():
                    ops.gemm(y, w1, b1, gws, N.EPI_BIAS_GLU)
                    ops.dwconv(gws.view(B, T, d), dw, db, cws.view(B, T, d))
                    if ln is None:
                        ops.gemm(cws, w2, b2, x, N.EPI_RESIDUAL, residual=x, alpha=1.0, row_valid=rv)
                    else:
                        ops.gemm_ln(cws, w2, b2, x, y2, alpha=1.0, g1=g, b1=b, row_valid=rv)
            t = timeit(fn)
            print(f"conv module {name:10s} {mode:6s} B={B} T={T}: {t:7.1f} us ({fl / t / 1e6:7.1f} TF/s)")


def mhsa_bench(B=64, T=248, H=4, d=256):
    dev = "cuda"
    qkv = torch.randn(B, T, 3, H, 64, device=dev).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    wo = (torch.randn(d, d, device=dev) / 16).bfloat16(); bo = torch.randn(d, device=dev)
    g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
    x = torch.randn(B * T, d, device=dev)
    y = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
    ctx = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
    mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)
    ln = {"y": y, "g1": g, "b1": b}
    fl = 4.0 * B * T * T * d + 2.0 * B * T * d * d
    t = timeit(lambda: ops.mhsa_out(q, k, v, wo, bo, x, mask=mask, scale=0.125, ln=ln, engine=N.ENGINE_TC))
    print(f"mhsa_out fused   (B={B},T={T}): {t:7.1f} us ({fl / t / 1e6:6.1f} TF/s)")
    def unfused():
        ops.attention(q, k, v, ctx.view(B, T, d), mask=mask, scale=0.125)
        ops.gemm_ln(ctx, wo, bo, x, y, alpha=1.0, g1=g, b1=b)
    t = timeit(unfused)
    print(f"attention + Wo/LN GEMM       : {t:7.1f} us ({fl / t / 1e6:6.1f} TF/s)")


def ctc_bench(M=15872, V=5002, d=256):
    dev = "cuda"
    x = torch.randn(M, d, device=dev).bfloat16()
    w = (torch.randn(V, d, device=dev) / 16).bfloat16(); b = torch.randn(V, device=dev) * 0.1
    fl = 2.0 * M * V * d
    t = timeit(lambda: ops.ctc_argmax(x, w, b), iters=5, per_graph=4)
    print(f"ctc_argmax fused (M={M}, V={V}): {t:7.1f} us ({fl / t / 1e6:6.1f} TF/s)")
    t = timeit(lambda: (torch.nn.functional.linear(x, w) + b).argmax(-1), iters=5, per_graph=4)
    print(f"torch linear + bias + argmax  : {t:7.1f} us")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        small_kernels()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ctc":
        ctc_bench()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mhsa":
        mhsa_bench()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "conv":
        conv_bench()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ffn":
        ffn_bench()
        sys.exit(0)
    main()
