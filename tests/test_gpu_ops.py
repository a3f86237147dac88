"""Op-level parity of every C-ABI kernel against a plain PyTorch fp32 restatement of the same
reference statements (level 1).  Runs on the B200 box: pytest -m gpu."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conformer_pytorch_lightning_b200 import _native as N
from conformer_pytorch_lightning_b200 import ops

DEV = "cuda"
TOL = {torch.float32: 2e-5, torch.bfloat16: 1.5e-2}     # max-abs-err / max-abs-ref


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rnd(*shape, dtype=torch.float32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


@pytest.mark.parametrize("d", [128, 256, 512])
@pytest.mark.parametrize("ydt", [torch.float32, torch.bfloat16])
def test_layernorm_variants(d, ydt):
    rows = 333
    x = rnd(rows, d, scale=3.0) + 0.5
    g1, b1, g2, b2 = rnd(d, seed=1) * 0.1 + 1, rnd(d, seed=2) * 0.1, rnd(d, seed=3) * 0.1 + 1, rnd(d, seed=4) * 0.1
    F = torch.nn.functional
    ref1 = F.layer_norm(x, (d,), g1, b1, 1e-5)
    ref2 = F.layer_norm(ref1, (d,), g2, b2, 1e-5)
    valid = (torch.arange(rows, device=DEV) % 5 != 0)
    # single LN -> y
    y = torch.empty(rows, d, dtype=ydt, device=DEV)
    ops.layernorm(x, g1, b1, y=y)
    assert rel_err(y.float(), ref1) < (1e-5 if ydt == torch.float32 else 6e-3)
    # chained LN, x_out in place, masked y
    xin = x.clone()
    ops.layernorm(xin, g1, b1, x_out=xin, g2=g2, b2=b2, y=y, row_valid=valid.to(torch.uint8))
    assert rel_err(xin, ref1) < 1e-5
    assert rel_err(y.float(), ref2 * valid[:, None]) < (1e-5 if ydt == torch.float32 else 6e-3)
    assert float(y[~valid].abs().max()) == 0.0


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,Nn,K", [(1, 256, 256), (77, 256, 2048), (300, 2048, 256), (129, 768, 256), (64, 512, 512),
                                    (16, 256, 2048), (16, 2048, 256), (16, 768, 256), (7, 255, 264), (13, 512, 1000)])
def test_gemm_epilogues_simt(dt, M, Nn, K):
    a = rnd(M, K, dtype=dt)
    w = rnd(Nn, K, dtype=dt, scale=1 / math.sqrt(K), seed=1)
    w2 = rnd(2 * Nn, K, dtype=dt, scale=1 / math.sqrt(K), seed=2)
    bias = rnd(Nn, seed=3)
    bias2 = rnd(2 * Nn, seed=4)
    af, wf, w2f = a.float(), w.float(), w2.float()
    lin = af @ wf.t() + bias
    tol = TOL[dt]
    for eng in (N.ENGINE_SIMT,):
        out = torch.empty(M, Nn, dtype=dt, device=DEV)
        ops.gemm(a, w, bias, out, N.EPI_BIAS, engine=eng)
        assert rel_err(out.float(), lin) < tol
        ops.gemm(a, w, bias, out, N.EPI_BIAS_SILU, engine=eng)
        assert rel_err(out.float(), torch.nn.functional.silu(lin)) < tol
        lin2 = af @ w2f.t() + bias2
        ops.gemm(a, w2, bias2, out, N.EPI_BIAS_GLU, engine=eng)
        assert rel_err(out.float(), lin2[:, :Nn] * torch.sigmoid(lin2[:, Nn:])) < tol
        res = rnd(M, Nn, seed=5)
        valid = (torch.arange(M, device=DEV) % 3 != 1)
        x = res.clone()
        ops.gemm(a, w, bias, x, N.EPI_RESIDUAL, residual=x, alpha=0.5, row_valid=valid.to(torch.uint8), engine=eng)
        assert rel_err(x, res + 0.5 * lin * valid[:, None]) < tol
        # no bias, strided A (row stride > K)
        abig = rnd(M, K + 64, dtype=dt, seed=6)
        ops.gemm(abig[:, :K], w, None, out, N.EPI_BIAS, engine=eng)
        assert rel_err(out.float(), abig[:, :K].float() @ wf.t()) < tol


def _attn_ref(q, k, v, mask, key_bias, scale):
    # attention.py:84-97 in fp32 on (B,T,H,64) layouts
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2)
    if key_bias is not None:
        s = s + key_bias[:, :, None, :]
    s = s * scale
    if mask is not None:
        m = mask.unsqueeze(1).eq(0)
        s = s.masked_fill(m, -float("inf"))
        p = torch.softmax(s, dim=-1).masked_fill(m, 0.0)
        p = torch.nan_to_num(p, nan=0.0)
    else:
        p = torch.softmax(s, dim=-1)
    return (p @ vf).permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[1], -1)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,Tq,Tk,kind", [(2, 4, 49, 49, "pad"), (3, 4, 74, 74, "chunk"), (1, 4, 16, 48, "none"),
                                           (2, 8, 130, 130, "left"), (1, 4, 16, 32, "bias"), (2, 4, 200, 200, "pad")])
def test_attention_simt(dt, B, H, Tq, Tk, kind):
    qkv = rnd(B, max(Tq, Tk), 3, H, 64, dtype=dt)
    q, k, v = qkv[:, :Tq, 0], qkv[:, :Tk, 1], qkv[:, :Tk, 2]
    mask, kb = None, None
    if kind == "pad":
        lens = torch.tensor([Tk, Tk - 13, Tk // 2][:B], device=DEV)
        mask = (torch.arange(Tk, device=DEV)[None, :] < lens[:, None]).unsqueeze(1)        # (B,1,Tk)
    elif kind in ("chunk", "left"):
        from conformer_pytorch_lightning_b200 import subsequent_chunk_mask
        cm = subsequent_chunk_mask(Tk, 16, 1 if kind == "left" else -1, torch.device(DEV))
        lens = torch.tensor([Tk, Tk - 20, 7][:B], device=DEV)
        pad = (torch.arange(Tk, device=DEV)[None, :] < lens[:, None]).unsqueeze(1)
        mask = pad & cm.unsqueeze(0)                                                     # (B,Tq,Tk) with empty rows
    elif kind == "bias":
        kb = rnd(B, H, Tk, seed=9)
    out = torch.empty(B, Tq, H * 64, dtype=dt, device=DEV)
    ops.attention(q, k, v, out, mask=mask, key_bias=kb, scale=0.125, engine=N.ENGINE_SIMT)
    ref = _attn_ref(q, k, v, mask, kb, 0.125)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref) < TOL[dt]
    if kind == "left":
        empty = ~mask.any(dim=-1)                      # fully masked rows -> exactly 0
        if empty.any():
            assert float(out.float()[empty].abs().max()) == 0.0


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,d,k", [(2, 49, 256, 15), (3, 130, 512, 31), (1, 16, 256, 15), (2, 70, 128, 7)])
def test_dwconv(dt, B, T, d, k):
    x = rnd(B, T, d, dtype=dt)
    w = rnd(d, 1, k, scale=0.3, seed=1)
    b = rnd(d, seed=2)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w, b, padding=(k - 1) // 2, groups=d).transpose(1, 2)
    y = torch.empty_like(x)
    ops.dwconv(x, w[:, 0].t().contiguous(), b, y, apply_silu=True)
    assert rel_err(y.float(), torch.nn.functional.silu(ref)) < TOL[dt]
    raw = torch.empty(B, T, d, dtype=torch.float32, device=DEV)
    ops.dwconv(x, w[:, 0].t().contiguous(), b, raw, apply_silu=False)
    assert rel_err(raw, ref) < 1e-5


def test_bn_stats_and_apply():
    rows, d = 1000, 256
    x = rnd(rows, d, scale=2.0) + 1.0
    st = torch.zeros(2, d, device=DEV)
    ops.bn_stats(x, st[0], st[1])
    assert rel_err(st[0], x.sum(0)) < 1e-5 and rel_err(st[1], (x * x).sum(0)) < 1e-5
    mean, var = x.mean(0), x.var(0, unbiased=False)
    g, b = rnd(d, seed=1), rnd(d, seed=2)
    for dt in (torch.float32, torch.bfloat16):
        y = torch.empty(rows, d, dtype=dt, device=DEV)
        ops.bn_apply_silu(x, mean, torch.rsqrt(var + 1e-5), g, b, y)
        ref = torch.nn.functional.silu((x - mean) * torch.rsqrt(var + 1e-5) * g + b)
        assert rel_err(y.float(), ref) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_relpos_keys(dt):
    B, Tk, H = 2, 37, 4
    k = rnd(B, Tk, 3, H, 64, dtype=dt)[:, :, 1]
    for Bp in (1, B):
        p = rnd(Bp, Tk, H * 64, dtype=dt, seed=3)
        u, vb = rnd(H, 64, seed=4), rnd(H, 64, seed=5)
        ko = torch.empty(B, Tk, H, 64, dtype=dt, device=DEV)
        kb = torch.empty(B, H, Tk, device=DEV)
        ops.relpos_keys(k, p, u, vb, ko, kb)
        p4 = p.float().view(Bp, Tk, H, 64)
        assert rel_err(ko.float(), k.float() + p4) < TOL[dt]
        ref_kb = torch.einsum("hc,bjhc->bhj", vb - u, p4).expand(B, H, Tk)
        assert rel_err(kb, ref_kb) < 1e-5


def test_errors_are_reported_not_thrown_across_abi():
    x = torch.zeros(4, 100, device=DEV)       # d=100 unsupported
    g = torch.ones(100, device=DEV)
    with pytest.raises(RuntimeError, match="cfm_layernorm"):
        ops.layernorm(x, g, g, y=torch.empty_like(x))
    before = N.launch_count()
    ops.layernorm(torch.zeros(4, 128, device=DEV), torch.ones(128, device=DEV), torch.ones(128, device=DEV),
                  y=torch.empty(4, 128, device=DEV))
    assert N.launch_count() == before + 1


@pytest.mark.parametrize("M,Nn,K", [(128, 256, 256), (200, 256, 2048), (1000, 2048, 256), (333, 768, 256), (4096, 512, 512),
                                    (15872, 2048, 256), (777, 384, 128), (64, 128, 64),
                                    # long-K shapes run on CTA pairs (cta_group::2, 256-row tiles): odd tile counts leave a
                                    # phantom half tile past M, N = 128 uses 64-row W boxes
                                    (300, 256, 1024), (1000, 512, 2048), (129, 128, 1024), (4000, 1536, 512)])
def test_gemm_tcgen05(M, Nn, K):
    """tcgen05/TMEM/TMA engine vs fp32 matmul of the same bf16 operands (and vs the SIMT engine)."""
    dt = torch.bfloat16
    pairs0 = N.kernel_launches("gemm_tc_pair")
    a = rnd(M, K, dtype=dt)
    w = rnd(Nn, K, dtype=dt, scale=1 / math.sqrt(K), seed=1)
    bias = rnd(Nn, seed=3)
    af, wf = a.float(), w.float()
    lin = af @ wf.t() + bias
    out = torch.empty(M, Nn, dtype=dt, device=DEV)
    ops.gemm(a, w, bias, out, N.EPI_BIAS, engine=N.ENGINE_TC)
    assert rel_err(out.float(), lin) < 6e-3
    ops.gemm(a, w, bias, out, N.EPI_BIAS_SILU, engine=N.ENGINE_TC)
    assert rel_err(out.float(), torch.nn.functional.silu(lin)) < 6e-3
    ops.gemm(a, w, None, out, N.EPI_BIAS, engine=N.ENGINE_TC)
    assert rel_err(out.float(), af @ wf.t()) < 6e-3
    res = rnd(M, Nn, seed=5)
    valid = (torch.arange(M, device=DEV) % 3 != 1)
    x = res.clone()
    ops.gemm(a, w, bias, x, N.EPI_RESIDUAL, residual=x, alpha=0.5, row_valid=valid.to(torch.uint8), engine=N.ENGINE_TC)
    assert rel_err(x, res + 0.5 * lin * valid[:, None]) < 1e-5      # fp32 output: only accumulation-order noise
    x2 = res.clone()
    ops.gemm(a, w, bias, x2, N.EPI_RESIDUAL, residual=x2, alpha=0.5, row_valid=valid.to(torch.uint8), engine=N.ENGINE_SIMT)
    assert rel_err(x, x2) < 1e-5
    if Nn % 256 == 0:   # GLU: W = [Wa;Wb] with Nn/2 outputs
        No = Nn // 2
        bias2 = rnd(Nn, seed=4)
        lin2 = af @ wf.t() + bias2
        o = torch.empty(M, No, dtype=dt, device=DEV)
        ops.gemm(a, w, bias2, o, N.EPI_BIAS_GLU, engine=N.ENGINE_TC)
        assert rel_err(o.float(), lin2[:, :No] * torch.sigmoid(lin2[:, No:])) < 6e-3
    # strided A / strided C views
    abig = rnd(M, K + 64, dtype=dt, seed=6)
    obig = torch.zeros(M, Nn + 8, dtype=dt, device=DEV)
    ops.gemm(abig[:, :K], w, bias, obig[:, :Nn], N.EPI_BIAS, engine=N.ENGINE_TC)
    assert rel_err(obig[:, :Nn].float(), abig[:, :K].float() @ wf.t() + bias) < 6e-3
    assert float(obig[:, Nn:].abs().max()) == 0.0
    if os.environ.get("CFM_B200_GEMM_PAIR") is None and M > 128:
        assert (N.kernel_launches("gemm_tc_pair") > pairs0) == (K >= 1024 or (K >= 512 and Nn >= 1024))


@pytest.mark.parametrize("B,H,Tq,Tk,kind", [(2, 4, 248, 248, "pad"), (3, 4, 74, 74, "chunk"), (2, 8, 130, 130, "left"),
                                           (1, 4, 300, 300, "none"), (2, 4, 128, 128, "pad"), (1, 2, 1498, 1498, "pad"),
                                           (2, 4, 129, 257, "none"), (2, 4, 64, 64, "chunk"), (1, 4, 40, 200, "padu8")])
def test_attention_tcgen05(B, H, Tq, Tk, kind):
    dt = torch.bfloat16
    qkv = rnd(B, max(Tq, Tk), 3, H, 64, dtype=dt)
    q, k, v = qkv[:, :Tq, 0], qkv[:, :Tk, 1], qkv[:, :Tk, 2]
    mask = None
    lens = torch.tensor([Tk, max(Tk - 13, 1), max(Tk // 2, 1)][:B], device=DEV)
    pad = (torch.arange(Tk, device=DEV)[None, :] < lens[:, None]).unsqueeze(1)
    if kind == "pad":
        mask = pad
    elif kind == "padu8":
        mask = pad.to(torch.uint8) * 7          # any non-zero byte is "visible" (.eq(0) semantics)
    elif kind in ("chunk", "left"):
        from conformer_pytorch_lightning_b200 import subsequent_chunk_mask
        cm = subsequent_chunk_mask(Tk, 16, 1 if kind == "left" else -1, torch.device(DEV))
        mask = pad & cm.unsqueeze(0)
    out = torch.full((B, Tq, H * 64), float("nan"), dtype=dt, device=DEV)
    ops.attention(q, k, v, out, mask=mask, scale=0.125, engine=N.ENGINE_TC)
    ref = _attn_ref(q, k, v, mask, None, 0.125)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref) < 1.5e-2
    out2 = torch.empty_like(out)
    ops.attention(q, k, v, out2, mask=mask, scale=0.125, engine=N.ENGINE_SIMT)
    assert rel_err(out.float(), out2.float()) < 1.5e-2
    if mask is not None and mask.shape[1] > 1:
        empty = ~mask.any(dim=-1)
        if empty.any():
            assert float(out.float()[empty].abs().max()) == 0.0


@pytest.mark.parametrize("M,K", [(128, 256), (333, 2048), (15872, 256), (200, 512)])
@pytest.mark.parametrize("mode", [1, 2])
def test_gemm_ln_fused_tcgen05(M, K, mode):
    """Residual GEMM with the following LayerNorm(s) in its epilogue vs fp32 torch (and vs the unfused path)."""
    dt, Nn = torch.bfloat16, 256
    a = rnd(M, K, dtype=dt)
    w = rnd(Nn, K, dtype=dt, scale=1 / math.sqrt(K), seed=1)
    bias = rnd(Nn, seed=3)
    g1, b1 = rnd(Nn, seed=4) * 0.1 + 1, rnd(Nn, seed=5) * 0.1
    g2, b2 = (rnd(Nn, seed=6) * 0.1 + 1, rnd(Nn, seed=7) * 0.1) if mode == 2 else (None, None)
    x0 = rnd(M, Nn, seed=8, scale=2.0)
    rv = (torch.arange(M, device=DEV) % 3 != 1).to(torch.uint8)
    yv = (torch.arange(M, device=DEV) % 5 != 2).to(torch.uint8)
    F = torch.nn.functional
    v = x0 + 0.5 * (a.float() @ w.float().t() + bias) * rv[:, None]
    if mode == 1:
        x_ref, y_ref = v, F.layer_norm(v, (Nn,), g1, b1, 1e-5)
    else:
        x_ref = F.layer_norm(v, (Nn,), g1, b1, 1e-5)
        y_ref = F.layer_norm(x_ref, (Nn,), g2, b2, 1e-5)
    y_ref = y_ref * yv[:, None]
    outs = {}
    for eng in (N.ENGINE_TC, N.ENGINE_SIMT):
        x = x0.clone()
        y = torch.full((M, Nn), float("nan"), dtype=dt, device=DEV)
        ops.gemm_ln(a, w, bias, x, y, alpha=0.5, g1=g1, b1=b1, g2=g2, b2=b2, row_valid=rv, y_row_valid=yv, engine=eng)
        assert rel_err(x, x_ref) < 2e-5, eng
        assert rel_err(y.float(), y_ref) < 6e-3, eng
        assert float(y.float()[yv == 0].abs().max()) == 0.0
        outs[eng] = (x, y)
    assert rel_err(outs[N.ENGINE_TC][0], outs[N.ENGINE_SIMT][0]) < 2e-5


def test_gemm_ln_fp32_unfused_path():
    M, K, Nn = 77, 256, 256
    a, w, bias = rnd(M, K), rnd(Nn, K, scale=1 / 16, seed=1), rnd(Nn, seed=2)
    g1, b1 = rnd(Nn, seed=4) * 0.1 + 1, rnd(Nn, seed=5) * 0.1
    x0 = rnd(M, Nn, seed=8)
    x = x0.clone()
    y = torch.empty(M, Nn, device=DEV)
    ops.gemm_ln(a, w, bias, x, y, alpha=1.0, g1=g1, b1=b1)
    v = x0 + a @ w.t() + bias
    assert rel_err(x, v) < 2e-5
    assert rel_err(y, torch.nn.functional.layer_norm(v, (Nn,), g1, b1, 1e-5)) < 2e-5


@pytest.mark.parametrize("M,F", [(128, 2048), (333, 2048), (15872, 2048), (200, 256), (20000, 512)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_ffn_fused_tcgen05(M, F, mode):
    """Whole feed-forward (+ LayerNorms) in one kernel vs fp32 torch on the same bf16 operands, and vs the
    unfused two-GEMM path of the library."""
    dt, d = torch.bfloat16, 256
    yin = rnd(M, d, dtype=dt)
    w1 = rnd(F, d, dtype=dt, scale=1 / 16, seed=1)
    b1 = rnd(F, seed=2) * 0.5
    w2 = rnd(d, F, dtype=dt, scale=1 / math.sqrt(F), seed=3)
    b2 = rnd(d, seed=4) * 0.5
    g1, be1 = rnd(d, seed=5) * 0.1 + 1, rnd(d, seed=6) * 0.1
    g2, be2 = rnd(d, seed=7) * 0.1 + 1, rnd(d, seed=8) * 0.1
    x0 = rnd(M, d, seed=9, scale=2.0)
    yv = (torch.arange(M, device=DEV) % 5 != 2).to(torch.uint8)
    Fn = torch.nn.functional
    h = Fn.silu(yin.float() @ w1.float().t() + b1).to(dt).float()       # hidden is rounded to bf16 on chip
    v = x0 + 0.5 * (h @ w2.float().t() + b2)
    ln = None
    if mode == 0:
        x_ref, y_ref = v, None
    elif mode == 1:
        x_ref, y_ref = v, Fn.layer_norm(v, (d,), g1, be1, 1e-5) * yv[:, None]
    else:
        x_ref = Fn.layer_norm(v, (d,), g1, be1, 1e-5)
        y_ref = Fn.layer_norm(x_ref, (d,), g2, be2, 1e-5) * yv[:, None]
    res = {}
    for eng in (N.ENGINE_TC, N.ENGINE_AUTO if False else N.ENGINE_SIMT):
        x = x0.clone()
        y = torch.full((M, d), float("nan"), dtype=dt, device=DEV)
        if mode == 1:
            ln = {"y": y, "g1": g1, "b1": be1, "y_row_valid": yv}
        elif mode == 2:
            ln = {"y": y, "g1": g1, "b1": be1, "g2": g2, "b2": be2, "y_row_valid": yv}
        hws = torch.empty(M, F, dtype=dt, device=DEV)
        ops.ffn(yin, w1, b1, w2, b2, x, alpha=0.5, ln=ln, hidden_ws=hws, engine=eng)
        assert torch.isfinite(x).all()
        assert rel_err(x, x_ref) < 2e-3, (eng, rel_err(x, x_ref))      # tanh.approx SiLU + bf16 hidden
        if y_ref is not None:
            assert rel_err(y.float(), y_ref) < 8e-3, eng
            assert float(y.float()[yv == 0].abs().max()) == 0.0
        res[eng] = x
    assert rel_err(res[N.ENGINE_TC], res[N.ENGINE_SIMT]) < 2e-3
    # in-place aliasing of the LayerNorm output with the FFN input (how the layer chain uses it)
    if mode == 1:
        x = x0.clone()
        yio = yin.clone()
        ops.ffn(yio, w1, b1, w2, b2, x, alpha=0.5, ln={"y": yio, "g1": g1, "b1": be1, "y_row_valid": yv}, engine=N.ENGINE_TC)
        assert rel_err(x, x_ref) < 2e-3 and rel_err(yio.float(), y_ref) < 8e-3


@pytest.mark.parametrize("M,F", [(128, 2048), (333, 2048), (15872, 2048), (200, 256), (20000, 512)])
@pytest.mark.parametrize("chained", [False, True])
def test_ffn_chain_projection_tcgen05(M, F, chained):
    """Feed-forward module(s) + LayerNorm + Q/K/V projection of the LayerNorm output in one kernel vs the separate library
    calls and vs fp32 torch."""
    dt, d, Np = torch.bfloat16, 256, 768
    yin = rnd(M, d, dtype=dt)
    mods = []
    for i in range(2):
        mods.append({"w1": rnd(F, d, dtype=dt, scale=1 / 16, seed=10 * i + 1), "b1": rnd(F, seed=10 * i + 2) * 0.5,
                     "w2": rnd(d, F, dtype=dt, scale=1 / math.sqrt(F), seed=10 * i + 3), "b2": rnd(d, seed=10 * i + 4) * 0.5,
                     "alpha": 0.5, "g1": rnd(d, seed=10 * i + 5) * 0.1 + 1, "be1": rnd(d, seed=10 * i + 6) * 0.1})
    mods[0].update(g2=rnd(d, seed=7) * 0.1 + 1, be2=rnd(d, seed=8) * 0.1)
    wp = rnd(Np, d, dtype=dt, scale=1 / 16, seed=31)
    bp = rnd(Np, seed=32) * 0.5
    x0 = rnd(M, d, seed=9, scale=2.0)
    Fn = torch.nn.functional

    def ref_mod(y, x, m):
        h = Fn.silu(y.float() @ m["w1"].float().t() + m["b1"]).to(dt).float()
        v = x + 0.5 * (h @ m["w2"].float().t() + m["b2"])
        if m.get("g2") is None:
            return v, Fn.layer_norm(v, (d,), m["g1"], m["be1"], 1e-5)
        xn = Fn.layer_norm(v, (d,), m["g1"], m["be1"], 1e-5)
        return xn, Fn.layer_norm(xn, (d,), m["g2"], m["be2"], 1e-5)
    x1, y1 = (x0, yin.float())
    if chained:
        x1, y1 = ref_mod(yin, x0, mods[0])
    x_ref, y_ref = ref_mod(y1.to(dt), x1, mods[1])
    p_ref = y_ref.to(dt).float() @ wp.float().t() + bp
    res = {}
    for eng in (N.ENGINE_TC, N.ENGINE_SIMT):
        x = x0.clone()
        y = yin.clone()
        pout = torch.full((M, Np), float("nan"), dtype=dt, device=DEV)
        hws = torch.empty(M, F, dtype=dt, device=DEV)
        ops.ffn_chain(y, mods[0] if chained else None, mods[1], x, y, proj=(wp, bp, pout), hidden_ws=hws, engine=eng)
        assert torch.isfinite(x).all() and torch.isfinite(pout.float()).all()
        assert rel_err(x, x_ref) < 4e-3, (eng, rel_err(x, x_ref))
        assert rel_err(pout.float(), p_ref) < 1.5e-2, (eng, rel_err(pout.float(), p_ref))
        res[eng] = (x, pout.float())
    assert rel_err(res[N.ENGINE_TC][0], res[N.ENGINE_SIMT][0]) < 4e-3
    assert rel_err(res[N.ENGINE_TC][1], res[N.ENGINE_SIMT][1]) < 1.5e-2


@pytest.mark.parametrize("M,F", [(128, 2048), (333, 2048), (15872, 2048), (200, 256), (20000, 512)])
@pytest.mark.parametrize("a_mode", [1, 2])
def test_ffn_chain_tcgen05(M, F, a_mode):
    """Two feed-forward modules chained in one kernel (X and the LayerNorm output in between stay on chip) vs the two
    cfm_ffn calls they replace and vs fp32 torch."""
    dt, d = torch.bfloat16, 256
    yin = rnd(M, d, dtype=dt)
    mods = []
    for i in range(2):
        mods.append({"w1": rnd(F, d, dtype=dt, scale=1 / 16, seed=10 * i + 1), "b1": rnd(F, seed=10 * i + 2) * 0.5,
                     "w2": rnd(d, F, dtype=dt, scale=1 / math.sqrt(F), seed=10 * i + 3), "b2": rnd(d, seed=10 * i + 4) * 0.5,
                     "alpha": 0.5, "g1": rnd(d, seed=10 * i + 5) * 0.1 + 1, "be1": rnd(d, seed=10 * i + 6) * 0.1})
    if a_mode == 2:
        mods[0].update(g2=rnd(d, seed=7) * 0.1 + 1, be2=rnd(d, seed=8) * 0.1)
    x0 = rnd(M, d, seed=9, scale=2.0)
    yv = (torch.arange(M, device=DEV) % 5 != 2).to(torch.uint8)
    Fn = torch.nn.functional

    def ref_mod(y, x, m, mode):
        h = Fn.silu(y.float() @ m["w1"].float().t() + m["b1"]).to(dt).float()
        v = x + 0.5 * (h @ m["w2"].float().t() + m["b2"])
        if mode == 1:
            return v, Fn.layer_norm(v, (d,), m["g1"], m["be1"], 1e-5)
        xn = Fn.layer_norm(v, (d,), m["g1"], m["be1"], 1e-5)
        return xn, Fn.layer_norm(xn, (d,), m["g2"], m["be2"], 1e-5)
    x1, y1 = ref_mod(yin, x0, mods[0], a_mode)
    x_ref, y_ref = ref_mod(y1.to(dt), x1, mods[1], 1)
    y_ref = y_ref * yv[:, None]
    res = {}
    for eng in (N.ENGINE_TC, N.ENGINE_SIMT):
        x = x0.clone()
        y = torch.full((M, d), float("nan"), dtype=dt, device=DEV)
        hws = torch.empty(M, F, dtype=dt, device=DEV)
        ops.ffn_chain(yin, mods[0], mods[1], x, y, y_row_valid=yv, hidden_ws=hws, engine=eng)
        assert torch.isfinite(x).all()
        assert rel_err(x, x_ref) < 4e-3, (eng, rel_err(x, x_ref))
        assert rel_err(y.float(), y_ref) < 1e-2, (eng, rel_err(y.float(), y_ref))
        assert float(y.float()[yv == 0].abs().max()) == 0.0
        res[eng] = x
    assert rel_err(res[N.ENGINE_TC], res[N.ENGINE_SIMT]) < 4e-3
    # in place on the LayerNorm buffer (how the layer chain uses it)
    x = x0.clone()
    yio = yin.clone()
    ops.ffn_chain(yio, mods[0], mods[1], x, yio, y_row_valid=yv, engine=N.ENGINE_TC)
    assert rel_err(x, x_ref) < 4e-3 and rel_err(yio.float(), y_ref) < 1e-2


@pytest.mark.parametrize("B,T", [(1, 64), (2, 128), (3, 129), (64, 248), (5, 256), (2, 100)])
@pytest.mark.parametrize("mask_kind", ["none", "pad", "full"])
@pytest.mark.parametrize("with_ln", [False, True])
def test_mhsa_out_fused_tcgen05(B, T, mask_kind, with_ln):
    """Attention over all heads + output projection + residual (+ LayerNorm) in one kernel vs fp32 torch on the same
    bf16 operands, and vs the unfused attention + GEMM chain of the library."""
    dt, H, d = torch.bfloat16, 4, 256
    qkv = rnd(B, T, 3, H, 64, dtype=dt, scale=1.5)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    wo = rnd(d, d, dtype=dt, scale=1 / 16, seed=1)
    bo = rnd(d, seed=2) * 0.5
    g1, be1 = rnd(d, seed=3) * 0.1 + 1, rnd(d, seed=4) * 0.1
    x0 = rnd(B * T, d, seed=5, scale=2.0)
    yv = (torch.arange(B * T, device=DEV) % 5 != 2).to(torch.uint8)
    mask = None
    if mask_kind == "pad":
        lens = torch.tensor([T - (7 * i) % max(T // 2, 1) for i in range(B)], device=DEV)
        mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None]).unsqueeze(1)            # (B,1,T)
    elif mask_kind == "full":
        idx = torch.arange(T, device=DEV)
        mask = ((idx[None, :] // 16) <= (idx[:, None] // 16)).unsqueeze(0).expand(B, T, T).contiguous()   # chunk-causal
        mask[:, T // 3, :] = False                                                                # a fully masked row
    Fn = torch.nn.functional
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    if mask is not None:
        m4 = mask.unsqueeze(1)
        s = s.masked_fill(~m4, float("-inf"))
        pr = torch.softmax(s, -1).masked_fill(~m4, 0.0)
        pr = torch.nan_to_num(pr, nan=0.0)
    else:
        pr = torch.softmax(s, -1)
    ctx = (pr @ vf).permute(0, 2, 1, 3).reshape(B * T, d)
    x_ref = x0 + ctx.to(dt).float() @ wo.float().t() + bo
    y_ref = Fn.layer_norm(x_ref, (d,), g1, be1, 1e-5) * yv[:, None]
    res = {}
    for eng in (N.ENGINE_TC, N.ENGINE_SIMT):
        x = x0.clone()
        y = torch.full((B * T, d), float("nan"), dtype=dt, device=DEV)
        ln = {"y": y, "g1": g1, "b1": be1, "y_row_valid": yv} if with_ln else None
        cws = torch.empty(B * T, d, dtype=dt, device=DEV)
        ops.mhsa_out(q, k, v, wo, bo, x, mask=mask, scale=0.125, ln=ln, ctx_ws=cws, engine=eng)
        assert torch.isfinite(x).all()
        assert rel_err(x, x_ref) < 4e-3, (eng, rel_err(x, x_ref))
        if with_ln:
            assert rel_err(y.float(), y_ref) < 8e-3, (eng, rel_err(y.float(), y_ref))
            assert float(y.float()[yv == 0].abs().max()) == 0.0
        res[eng] = x
    assert rel_err(res[N.ENGINE_TC], res[N.ENGINE_SIMT]) < 4e-3


@pytest.mark.parametrize("B,T", [(1, 64), (1, 114), (2, 115), (3, 57), (5, 15), (64, 248), (7, 333), (1, 1000)])
@pytest.mark.parametrize("with_ln", [False, True])
def test_conv_module_fused_tcgen05(B, T, with_ln):
    """Whole convolution module (+ LayerNorm) in one kernel vs fp32 torch on the same bf16 operands (intermediates
    rounded to bf16 as on chip), and vs the unfused three-kernel chain of the library."""
    dt, d, k = torch.bfloat16, 256, 15
    M = B * T
    Fn = torch.nn.functional
    rv = (torch.arange(M, device=DEV) % 7 != 3).to(torch.uint8)
    if B > 1:
        rv.view(B, T)[1, T // 2:] = 0                                    # a padded tail
    yin = rnd(M, d, dtype=dt) * rv[:, None].to(dt)
    w1 = rnd(2 * d, d, dtype=dt, scale=1 / 16, seed=1)
    b1 = rnd(2 * d, seed=2) * 0.5
    dw = rnd(d, 1, k, scale=0.3, seed=3)
    db = rnd(d, seed=4) * 0.5
    w2 = rnd(d, d, dtype=dt, scale=1 / 16, seed=5)
    b2 = rnd(d, seed=6) * 0.5
    g1, be1 = rnd(d, seed=7) * 0.1 + 1, rnd(d, seed=8) * 0.1
    x0 = rnd(M, d, seed=9, scale=2.0)
    u = yin.float() @ w1.float().t() + b1
    g = (u[:, :d] * torch.sigmoid(u[:, d:])).to(dt).float()
    c = Fn.conv1d(g.view(B, T, d).transpose(1, 2), dw, db, padding=(k - 1) // 2, groups=d).transpose(1, 2)
    c = Fn.silu(c).reshape(M, d).to(dt).float()
    x_ref = x0 + (c @ w2.float().t() + b2) * rv[:, None]
    y_ref = Fn.layer_norm(x_ref, (d,), g1, be1, 1e-5)
    dw_t = dw[:, 0].t().contiguous()
    res = {}
    for eng in (N.ENGINE_TC, N.ENGINE_SIMT):
        x = x0.clone()
        y = torch.full((M, d), float("nan"), dtype=dt, device=DEV)
        ln = {"y": y, "g1": g1, "b1": be1} if with_ln else None
        gws, cws = torch.empty(M, d, dtype=dt, device=DEV), torch.empty(M, d, dtype=dt, device=DEV)
        ops.conv_module(yin, w1, b1, dw_t, db, w2, b2, x, B, T, row_valid=rv, ln=ln, glu_ws=gws, dw_ws=cws, engine=eng)
        assert torch.isfinite(x).all()
        assert rel_err(x, x_ref) < 3e-3, (eng, rel_err(x, x_ref))
        if with_ln:
            assert rel_err(y.float(), y_ref) < 8e-3, (eng, rel_err(y.float(), y_ref))
        res[eng] = x
    assert rel_err(res[N.ENGINE_TC], res[N.ENGINE_SIMT]) < 3e-3


@pytest.mark.parametrize("B,Tin,C", [(2, 200, 256), (64, 998, 256), (3, 131, 512), (1, 67, 256)])
def test_native_subsampling_frontend(B, Tin, C):
    """conv1 (CUDA cores) + conv2 (tcgen05 implicit GEMM with TMA-gathered taps) vs torch conv2d on the same weights."""
    torch.manual_seed(0)
    idim = 80
    x = rnd(B, Tin, idim)
    conv1 = torch.nn.Conv2d(1, C, 3, 2).to(DEV)
    conv2 = torch.nn.Conv2d(C, C, 3, 2).to(DEV)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h = torch.relu(conv1(x.unsqueeze(1)))
        ref = torch.relu(torch.nn.functional.conv2d(h.bfloat16().float(), conv2.weight.bfloat16().float(), conv2.bias, stride=2))
    T2, F2 = ref.shape[2], ref.shape[3]
    w1 = conv1.weight.detach().reshape(C, 9).contiguous()
    w2 = conv2.weight.detach().permute(0, 2, 3, 1).reshape(C, 9 * C).bfloat16().contiguous()
    ws = torch.empty(ops.subsample_ws_bytes(B, Tin, idim, C), dtype=torch.uint8, device=DEV)
    out = torch.full((B, T2, F2, C), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.subsample_conv(x, w1, conv1.bias.detach(), w2, conv2.bias.detach(), ws, out)
    got = out.float().permute(0, 3, 1, 2)          # (B, C, T2, F2)
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 1.5e-2


@pytest.mark.parametrize("M,V,d", [(15872, 5002, 256), (333, 5002, 256), (64, 300, 256), (1000, 5002, 512), (37, 5002, 256)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_ctc_argmax(M, V, d, dt):
    """ctc_lo projection + frame argmax without materialising the logits (tcgen05) / through the chunked workspace
    (CUDA cores) vs torch on the same operands."""
    x = rnd(M, d, dtype=dt)
    w = rnd(V, d, dtype=dt, scale=1 / math.sqrt(d), seed=1)
    b = rnd(V, seed=2) * 0.1
    logits = x.float() @ w.float().t() + b
    ref_best, ref_ids = logits.max(dim=1)
    for eng in ((N.ENGINE_AUTO, N.ENGINE_SIMT) if dt == torch.bfloat16 else (N.ENGINE_SIMT,)):
        ids, best = ops.ctc_argmax(x, w, b, want_best=True, engine=eng)
        ids = ids.long()
        assert ids.min() >= 0 and ids.max() < V
        picked = logits.gather(1, ids[:, None])[:, 0]
        # the chosen column must be a maximiser up to accumulation-order / bf16-logit noise
        on_tc = dt == torch.bfloat16 and eng != N.ENGINE_SIMT and M >= 64      # else: logits rounded to the act dtype
        tol = 1e-4 if (dt == torch.float32 or on_tc) else 3e-2
        assert float((ref_best - picked).max()) <= tol * float(ref_best.abs().max() + 1), eng
        agree = float((ids == ref_ids).float().mean())
        assert agree > (0.999 if tol == 1e-4 else 0.9), (eng, agree)
        if on_tc or dt == torch.float32:
            assert rel_err(best, picked) < 1e-3


def test_ctc_greedy_head_matches_oracle_decode():
    from conformer_pytorch_lightning_b200 import CTCGreedyHead
    torch.manual_seed(0)
    head = CTCGreedyHead(256, 5002).to(DEV)
    hs = rnd(3, 50, 256, seed=7)
    lens = [50, 31, 7]
    with torch.no_grad():
        ids, hyps = head.greedy(hs, lens)
        logits = hs @ head.ctc_lo.weight.t() + head.ctc_lo.bias
    want_ids = logits.argmax(-1)
    assert float((ids == want_ids).float().mean()) > 0.999
    for b, n in enumerate(lens):
        seq, prev = [], -1
        for t in ids[b, :n].tolist():
            if t != prev and t != 0:
                seq.append(t)
            prev = t
        assert hyps[b] == seq


@pytest.mark.parametrize("engine", [N.ENGINE_TC, N.ENGINE_SIMT])
def test_gemm_fp32_output_without_residual(engine):
    """EPI_RESIDUAL with residual = None: X = alpha * rowmask(A W^T + b) in fp32 (the front-end Linear writes its fp32
    output this way instead of accumulating onto a zeroed buffer)."""
    M, K, Nn = 1000, 256, 256
    a = rnd(M, K, dtype=torch.bfloat16, seed=41)
    w = rnd(Nn, K, dtype=torch.bfloat16, seed=42)
    bias = rnd(Nn, dtype=torch.float32, seed=43)
    out = torch.full((M, Nn), float("nan"), dtype=torch.float32, device=DEV)
    rv = (torch.arange(M, device=DEV) % 9 != 4)
    ops.gemm(a, w, bias, out, N.EPI_RESIDUAL, residual=None, alpha=0.5, row_valid=rv, engine=engine)
    ref = 0.5 * (a.float() @ w.float().t() + bias) * rv[:, None]
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 1e-5
