"""Kernel-level tests of the training-path ops (general GEMM, LayerNorm / activation / conv / BatchNorm backward, softmax,
CTC loss) against PyTorch fp32 restatements of the same ops.  pytest -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conformer_pytorch_lightning_b200 import _native as N
from conformer_pytorch_lightning_b200 import ops


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _rand(*shape, dtype=torch.bfloat16, seed=0):
    g = torch.Generator().manual_seed(seed + int(np.prod(shape)) % 9973)
    return torch.randn(*shape, generator=g).to(dtype).cuda()


# ------------------------------------------------------------------------------------------------ general GEMM
@pytest.mark.parametrize("engine", [N.ENGINE_TC, N.ENGINE_SIMT])
@pytest.mark.parametrize("M,Nn,K", [(3968, 256, 2048), (3968, 2048, 256), (304, 768, 256), (248, 248, 64), (136, 64, 248)])
@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_ex_majors(engine, M, Nn, K, a_t, b_t):
    """All four operand-major combinations (K-major = nn.Linear layout, MN-major = transposed view), bf16 and fp32 out."""
    a = _rand(K, M, seed=1).t() if a_t else _rand(M, K, seed=1)
    b = _rand(K, Nn, seed=2).t() if b_t else _rand(Nn, K, seed=2)
    ref = a.float() @ b.float().t()
    out32 = torch.full((M, Nn), 7.0, dtype=torch.float32, device="cuda")
    ops.gemm_ex(a, b, out32, alpha=0.5, engine=engine)
    assert _rel(out32, 0.5 * ref) < 1e-5
    out16 = torch.empty((M, Nn), dtype=torch.bfloat16, device="cuda")
    ops.gemm_ex(a, b, out16, engine=engine)
    assert _rel(out16.float(), ref) < 6e-3


@pytest.mark.parametrize("engine", [N.ENGINE_TC, N.ENGINE_SIMT])
@pytest.mark.parametrize("splits", [0, 1, 7])
def test_gemm_ex_wgrad_accumulate(engine, splits):
    """dW += dC^T X with both operands as transposed views, fp32 accumulation on top of existing gradient values."""
    M, Nout, Kw = 3968 + 24, 2048, 256
    dc, x = _rand(M, Nout, seed=3), _rand(M, Kw, seed=4)
    grad = _rand(Nout, Kw, dtype=torch.float32, seed=5)
    ref = grad.double() + dc.double().t() @ x.double()
    ops.gemm_ex(dc.t(), x.t(), grad, accumulate=True, splits=splits, engine=engine)
    assert _rel(grad, ref) < 2e-5


@pytest.mark.parametrize("engine", [N.ENGINE_TC, N.ENGINE_SIMT])
@pytest.mark.parametrize("T", [248, 74, 311])
def test_gemm_ex_batched_attention_products(engine, T):
    """The six batched products of attention forward/backward on the (B, T, 3, H, 64) projection layout."""
    B, H = 3, 4
    qkv = _rand(B, T, 3, H, 64, seed=6)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))            # (B, H, T, 64) strided views
    Tp = (T + 7) // 8 * 8
    s = torch.zeros((B, H, T, Tp), dtype=torch.float32, device="cuda")
    ops.gemm_ex(q, k, s[..., :T], alpha=0.125, engine=engine)                    # S = Q K^T / sqrt(d_k)
    ref_s = 0.125 * q.float() @ k.float().transpose(-1, -2)
    assert _rel(s[..., :T], ref_s) < 1e-5 and float(s[..., T:].abs().max() if Tp > T else 0) == 0
    p = torch.zeros((B, H, T, Tp), dtype=torch.bfloat16, device="cuda")
    p[..., :T] = torch.softmax(ref_s, -1).to(torch.bfloat16)
    ctx = torch.empty((B, T, H, 64), dtype=torch.bfloat16, device="cuda")
    ops.gemm_ex(p[..., :T], v.transpose(-1, -2), ctx.permute(0, 2, 1, 3), engine=engine)   # O = P V
    ref_o = p[..., :T].float() @ v.float()
    assert _rel(ctx.permute(0, 2, 1, 3).float(), ref_o) < 6e-3
    do = _rand(B, T, H, 64, seed=7).permute(0, 2, 1, 3)
    dp = torch.empty((B, H, T, Tp), dtype=torch.float32, device="cuda")
    ops.gemm_ex(do, v, dp[..., :T], engine=engine)                               # dP = dO V^T
    assert _rel(dp[..., :T], do.float() @ v.float().transpose(-1, -2)) < 1e-5
    dqkv = torch.zeros((B, T, 3, H, 64), dtype=torch.bfloat16, device="cuda")
    dq, dk, dv = (dqkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ops.gemm_ex(p[..., :T].transpose(-1, -2), do.transpose(-1, -2), dv, engine=engine)       # dV = P^T dO
    assert _rel(dv.float(), p[..., :T].float().transpose(-1, -2) @ do.float()) < 6e-3
    ds = (p * 0.3).contiguous()
    ops.gemm_ex(ds[..., :T], k.transpose(-1, -2), dq, alpha=0.125, engine=engine)             # dQ = dS K / sqrt(d_k)
    assert _rel(dq.float(), 0.125 * ds[..., :T].float() @ k.float()) < 6e-3
    ops.gemm_ex(ds[..., :T].transpose(-1, -2), q.transpose(-1, -2), dk, alpha=0.125, engine=engine)   # dK = dS^T Q
    assert _rel(dk.float(), 0.125 * ds[..., :T].float().transpose(-1, -2) @ q.float()) < 6e-3


def test_gemm_ex_fp32_inputs():
    a, b = _rand(200, 96, dtype=torch.float32, seed=8), _rand(150, 96, dtype=torch.float32, seed=9)
    out = torch.empty((200, 150), dtype=torch.float32, device="cuda")
    ops.gemm_ex(a, b, out)
    assert _rel(out, a @ b.t()) < 1e-5
    ops.gemm_ex(a.t().contiguous().t(), b, out, accumulate=True)
    assert _rel(out, 2 * (a @ b.t())) < 1e-5


# ------------------------------------------------------------------------------------------------ memory-bound training kernels
from conformer_pytorch_lightning_b200 import train_ops as TO

DTYPES = [torch.float32, torch.bfloat16]


def _seed(v):
    return torch.tensor([v], dtype=torch.int64, device="cuda")


def _tol(dtype):
    return 2e-5 if dtype == torch.float32 else 1.2e-2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,d", [(1000, 256), (333, 512)])
def test_layernorm_fwd_bwd(dtype, rows, d):
    x = _rand(rows, d, dtype=torch.float32, seed=11).requires_grad_()
    g = (_rand(d, dtype=torch.float32, seed=12) * 0.1 + 1).requires_grad_()
    b = (_rand(d, dtype=torch.float32, seed=13) * 0.1).requires_grad_()
    valid = (torch.arange(rows, device="cuda") % 7 != 3)
    y = torch.empty((rows, d), dtype=dtype, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    TO.ln_fwd(x.detach(), g.detach(), b.detach(), y, mean, rstd, row_valid=valid)
    ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5) * valid[:, None]
    assert _rel(y.float(), ref) < (1e-5 if dtype == torch.float32 else 5e-3)
    dy = _rand(rows, d, dtype=dtype, seed=14)
    dx_in = _rand(rows, d, dtype=torch.float32, seed=15)
    (ref * dy.float()).sum().backward()
    dx = torch.empty_like(dx_in)
    dg, db = torch.ones(d, device="cuda"), torch.ones(d, device="cuda")        # accumulate on top of existing values
    TO.ln_bwd(dy, x.detach(), mean, rstd, g.detach(), dx, dg, db, dx_in=dx_in, row_valid=valid)
    assert _rel(dx, x.grad + dx_in) < 2e-5
    assert _rel(dg - 1, g.grad) < 2e-5 and _rel(db - 1, b.grad) < 2e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_silu_dropout_fwd_bwd(dtype):
    rows, cols = 777, 2048
    h = _rand(rows, cols, dtype=dtype, seed=16)
    a = torch.empty_like(h)
    TO.silu_dropout_fwd(h, a)
    hr = h.float().requires_grad_()
    ref = torch.nn.functional.silu(hr)
    assert _rel(a.float(), ref) < _tol(dtype)
    da = _rand(rows, cols, dtype=dtype, seed=17)
    ref.backward(da.float())
    dh, dbias = torch.empty_like(h), torch.zeros(cols, device="cuda")
    TO.silu_dropout_bwd(da, h, dh, dbias)
    assert _rel(dh.float(), hr.grad) < _tol(dtype)
    assert _rel(dbias, dh.float().sum(0)) < 1e-5
    # dropout: same mask in forward and backward, right rate and scale, deterministic in (seed, site)
    p = 0.1
    TO.silu_dropout_fwd(h, a, p=p, seed=_seed(1234), site=5)
    a2 = torch.empty_like(a)
    TO.silu_dropout_fwd(h, a2, p=p, seed=_seed(1234), site=5)
    assert torch.equal(a, a2)
    TO.silu_dropout_fwd(h, a2, p=p, seed=_seed(1234), site=6)
    assert not torch.equal(a, a2)
    nz = ref.detach().abs() > 1e-3
    kept = (a.float() != 0) & nz
    rate = 1 - kept.sum().item() / nz.sum().item()
    assert abs(rate - p) < 0.005
    assert _rel(a.float()[kept], (ref.detach() / (1 - p))[kept]) < _tol(dtype)
    TO.silu_dropout_bwd(da, h, dh, None, p=p, seed=_seed(1234), site=5)
    big = nz & (da.float().abs() > 1e-2) & (hr.grad.abs() > 1e-3)
    assert torch.equal((dh.float() != 0) & big, kept & big)


@pytest.mark.parametrize("dtype", DTYPES)
def test_residual_dropout_pair(dtype):
    rows, cols = 500, 256
    x0 = _rand(rows, cols, dtype=torch.float32, seed=18)
    f = _rand(rows, cols, dtype=dtype, seed=19)
    valid = (torch.arange(rows, device="cuda") % 5 != 0)
    x = x0.clone()
    TO.resid_dropout_add(x, f, alpha=0.5, row_valid=valid)
    assert _rel(x, x0 + 0.5 * f.float() * valid[:, None]) < 1e-6
    x = x0.clone()
    TO.resid_dropout_add(x, f, alpha=0.5, row_valid=valid, p=0.25, seed=_seed(7), site=3)
    mult = (x - x0) / (0.5 * f.float())                       # 0 or 1/(1-p) on valid rows
    m = mult[valid]
    assert ((m.abs() < 1e-3) | ((m - 1 / 0.75).abs() < 2e-2)).all()
    assert abs((m.abs() < 1e-3).float().mean().item() - 0.25) < 0.01
    dx = _rand(rows, cols, dtype=torch.float32, seed=20)
    df, dbias = torch.empty((rows, cols), dtype=dtype, device="cuda"), torch.zeros(cols, device="cuda")
    TO.scale_dropout_bwd(dx, df, dbias, alpha=0.5, row_valid=valid, p=0.25, seed=_seed(7), site=3)
    ref = 0.5 * dx * torch.where(mult.abs() < 1e-3, 0.0, 1 / 0.75) * valid[:, None]
    known = f.float().abs() > 1e-4                            # where f ~ 0 the forward does not reveal the mask (x - x0 underflows)
    assert _rel(df.float() * known, ref * known) < _tol(dtype)
    assert _rel(dbias, df.float().sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_glu_fwd_bwd(dtype):
    rows, d = 999, 256
    g = _rand(rows, 2 * d, dtype=dtype, seed=21)
    u = torch.empty((rows, d), dtype=dtype, device="cuda")
    TO.glu_fwd(g, u)
    gr = g.float().requires_grad_()
    ref = torch.nn.functional.glu(gr, dim=1)
    assert _rel(u.float(), ref) < _tol(dtype)
    du = _rand(rows, d, dtype=torch.float32, seed=22)
    ref.backward(du)
    dg, dbias = torch.empty_like(g), torch.zeros(2 * d, device="cuda")
    TO.glu_bwd(du, g, dg, dbias)
    assert _rel(dg.float(), gr.grad) < _tol(dtype)
    assert _rel(dbias, dg.float().sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_bn_silu_bwd(dtype):
    rows, d = 1984, 256
    raw = (_rand(rows, d, dtype=torch.float32, seed=23) * 2 + 0.3).requires_grad_()
    gamma = (_rand(d, dtype=torch.float32, seed=24) * 0.2 + 1).requires_grad_()
    beta = (_rand(d, dtype=torch.float32, seed=25) * 0.2).requires_grad_()
    out = torch.nn.functional.silu(torch.nn.functional.batch_norm(raw, None, None, gamma, beta, True, 0.1, 1e-5))
    dc = _rand(rows, d, dtype=dtype, seed=26)
    out.backward(dc.float())
    mean = raw.detach().mean(0)
    rstd = torch.rsqrt(raw.detach().var(0, unbiased=False) + 1e-5)
    sums = torch.empty(2 * d, device="cuda")
    draw = torch.empty((rows, d), dtype=dtype, device="cuda")
    TO.bn_silu_bwd(dc, raw.detach(), mean, rstd, gamma.detach(), beta.detach(), sums, draw)
    assert _rel(sums[:d], gamma.grad) < _tol(dtype) and _rel(sums[d:], beta.grad) < _tol(dtype)
    assert _rel(draw.float(), raw.grad) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k", [15, 31])
def test_dwconv_backward(dtype, k):
    """dgrad = cfm_dwconv with the tap-reversed filter; wgrad = cfm_dwconv_wgrad."""
    B, T, d = 3, 130, 256
    u = _rand(B, T, d, dtype=dtype, seed=27)
    w = _rand(k, d, dtype=torch.float32, seed=28) * 0.3
    ur = u.float().transpose(1, 2).requires_grad_()
    wr = w.t().reshape(d, 1, k).clone().requires_grad_()
    br = torch.zeros(d, device="cuda", requires_grad=True)
    out = torch.nn.functional.conv1d(ur, wr, br, padding=(k - 1) // 2, groups=d)
    dy = _rand(B, T, d, dtype=dtype, seed=29)
    out.backward(dy.float().transpose(1, 2))
    du = torch.empty((B, T, d), dtype=torch.float32, device="cuda")
    ops.dwconv(dy, w.flip(0).contiguous(), torch.zeros(d, device="cuda"), du, apply_silu=False)
    assert _rel(du, ur.grad.transpose(1, 2)) < 2e-5
    dw, dbias = torch.zeros(k, d, device="cuda"), torch.zeros(d, device="cuda")
    TO.dwconv_wgrad(dy, u, dw, dbias)
    assert _rel(dw, wr.grad.reshape(d, k).t()) < 2e-5 and _rel(dbias, br.grad) < 2e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("T", [74, 248])
def test_softmax_fwd_bwd(dtype, T):
    B, H = 2, 4
    Tp = (T + 7) // 8 * 8
    S = torch.zeros(B, H, T, Tp, device="cuda")
    S[..., :T] = _rand(B, H, T, T, dtype=torch.float32, seed=30) * 3
    i = torch.arange(T, device="cuda")
    mask = (i[None, :] < (i[:, None] // 16 + 1) * 16)[None].expand(B, T, T).clone()      # chunk-16 mask
    mask[1, :, T - 9:] = False                                                            # + key padding
    mask[1, 5] = False                                                                    # a fully masked row
    P = torch.empty((B, H, T, Tp), dtype=dtype, device="cuda")
    TO.softmax_fwd(S, P, None, mask, Tk=T)
    Sr = S[..., :T].clone().requires_grad_()
    m = ~mask[:, None]
    ref = torch.softmax(Sr.masked_fill(m, float("-inf")), -1).masked_fill(m, 0.0)
    ref = torch.nan_to_num(ref)                  # the reference yields NaN -> 0 by masked_fill for fully masked rows
    assert _rel(P[..., :T].float(), ref) < (1e-6 if dtype == torch.float32 else 5e-3)
    assert float(P[..., T:].float().abs().max()) == 0 if Tp > T else True
    assert float(P[1, :, 5].float().abs().max()) == 0
    dP = torch.zeros(B, H, T, Tp, device="cuda")
    dP[..., :T] = _rand(B, H, T, T, dtype=torch.float32, seed=31)
    Sq = S[..., :T].clone().requires_grad_()
    torch.softmax(Sq.masked_fill(m, float("-inf")), -1).masked_fill(m, 0.0)[:, :, [j for j in range(T) if j != 5] if True else slice(None)]
    Pq = torch.softmax(Sq.masked_fill(m, -1e30), -1) * (~m)
    (Pq * dP[..., :T]).sum().backward()
    dS = torch.empty_like(P)
    TO.softmax_bwd(P, dP, dS, Tk=T)
    good = torch.ones(B, H, T, dtype=torch.bool, device="cuda")
    good[1, :, 5] = False                                      # fully masked row: zero gradient here, NaN in torch
    assert _rel(dS[..., :T].float()[good], Sq.grad[good]) < (2e-5 if dtype == torch.float32 else 2e-2)
    assert float(dS[1, :, 5].float().abs().max()) == 0
    # dropout on the probabilities
    Pd = torch.empty_like(P)
    TO.softmax_fwd(S, P, Pd, mask, Tk=T, p=0.1, seed=_seed(99), site=2)
    nz = P.float() > 1e-3
    dropped = (Pd.float() == 0) & nz
    assert abs(dropped.sum().item() / nz.sum().item() - 0.1) < 0.01
    keep = nz & ~dropped
    assert _rel(Pd.float()[keep], P.float()[keep] / 0.9) < 1e-2


@pytest.mark.parametrize("dtype", DTYPES)
def test_colsum_strided(dtype):
    x = _rand(1000, 768, dtype=dtype, seed=32)
    out = torch.ones(256, device="cuda")
    TO.colsum(x[:, 256:512], out)
    assert _rel(out - 1, x[:, 256:512].float().sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_ctc_loss_fwd_bwd_vs_torch(dtype):
    """log_softmax + nn.CTCLoss(reduction='sum') / Lmax (decoder.py:18-23) and its gradient w.r.t. the logits; ragged
    input and label lengths, repeated labels."""
    B, T, V, Lmax = 5, 60, 5002, 12
    Vp = 5120
    g = torch.Generator().manual_seed(3)
    logits = torch.zeros(B * T, Vp, dtype=dtype, device="cuda")
    logits[:, :V] = (torch.randn(B * T, V, generator=g) * 2).to(dtype).cuda()
    labels = torch.randint(1, V, (B, Lmax), generator=g)
    labels[0, 3] = labels[0, 2]                                   # repeated label: needs a blank in between
    labels[2, :] = labels[2, 0]
    in_len = torch.tensor([60, 55, 60, 31, 47], dtype=torch.int32)
    lab_len = torch.tensor([12, 10, 12, 1, 7], dtype=torch.int32)
    lr = logits[:, :V].float().view(B, T, V).clone().requires_grad_()
    lp = lr.transpose(0, 1).log_softmax(2)
    loss_ref = torch.nn.functional.ctc_loss(lp, labels.cuda(), in_len.cuda().long(), lab_len.cuda().long(), blank=0,
                                            reduction="sum") / Lmax
    loss_ref.backward()
    nll = torch.empty(B, device="cuda")
    ws = TO.ctc_loss_ws(B, T, Lmax, "cuda")
    lab32, il, ll = labels.to(torch.int32).cuda(), in_len.cuda(), lab_len.cuda()
    TO.ctc_loss_fwd(logits, B, T, V, lab32, il, ll, nll, ws)
    assert abs(nll.sum().item() / Lmax - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    dl = torch.empty_like(logits)
    TO.ctc_loss_bwd(logits, B, T, V, lab32, il, ll, nll, ws, 1.0 / Lmax, dl)
    assert _rel(dl[:, :V].float().view(B, T, V), lr.grad) < (3e-4 if dtype == torch.float32 else 8e-3)   # torch.ctc_loss itself is an fp32 log-space recursion
    assert float(dl[:, V:].float().abs().max()) == 0
