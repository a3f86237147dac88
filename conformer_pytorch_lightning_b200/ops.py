"""Tensor-level wrappers over the C-ABI (include/cfm_b200.h).

PyTorch is used here for device memory and streams only: every function marshals
``data_ptr()``s and the *current* CUDA stream of the calling thread into one native
call.  CPU tensors raise (there is no CPU path).
"""
from __future__ import annotations

import torch

from . import _native as N

_DT = {torch.float32: N.F32, torch.bfloat16: N.BF16}


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t, name, dtype=None, contiguous=True):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the B200 Conformer kernels have no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


def ensure_init(t):
    """Per-device library setup on first use; kernels are launched on the calling thread's CURRENT device, so a tensor
    on another device is an error here rather than an illegal-address fault later."""
    cur = torch.cuda.current_device()
    idx = t.device.index if t.device.index is not None else cur
    if idx != cur:
        raise RuntimeError(f"tensor on cuda:{idx} but the current device is cuda:{cur}: wrap the call in "
                           f"`with torch.cuda.device({idx}):`")
    N.init(idx)


def layernorm(x, g1, b1, *, x_out=None, g2=None, b2=None, y=None, row_valid=None, eps=1e-5):
    """t = LN(x;g1,b1); x_out = t; t = LN(t;g2,b2) if g2; y = rowmask(t).  x (rows,d) fp32."""
    _req(x, "layernorm.x", torch.float32)
    rows, d = x.shape
    ensure_init(x)
    if y is not None:
        _req(y, "layernorm.y")
    N.check(N.lib().cfm_layernorm(x.data_ptr(), rows, d, g1.data_ptr(), b1.data_ptr(), _ptr(x_out),
                                  _ptr(g2), _ptr(b2), _ptr(y), _DT[y.dtype] if y is not None else N.F32,
                                  _ptr(row_valid), float(eps), _stream(x)))
    return y


def gemm(a, w, bias, out, epilogue, *, residual=None, alpha=1.0, row_valid=None, engine=N.ENGINE_AUTO):
    """out = epilogue(a @ w.T + bias); a (M,K) (row stride may exceed K), w (N or 2N, K)."""
    _req(a, "gemm.a", contiguous=False)
    _req(w, "gemm.w", a.dtype)
    _req(out, "gemm.out", contiguous=False)
    if a.stride(1) != 1 or out.stride(1) != 1:
        raise RuntimeError("gemm: inner dimension must be contiguous")
    M, K = a.shape
    Nn = out.shape[1]
    ensure_init(a)
    N.check(N.lib().cfm_gemm(a.data_ptr(), a.stride(0), w.data_ptr(), _ptr(bias), out.data_ptr(), out.stride(0),
                             M, Nn, K, _DT[a.dtype], epilogue, _ptr(residual), float(alpha), _ptr(row_valid),
                             engine, _stream(a)))
    return out


def gemm_ln(a, w, bias, x, y, *, alpha, g1, b1, g2=None, b2=None, row_valid=None, y_row_valid=None, eps=1e-5,
            engine=N.ENGINE_AUTO):
    """x (M,N) fp32 in place: v = x + alpha*rowmask(a @ w.T + bias); then the fused LayerNorm(s) -> x, y
    (see cfm_gemm_ln in include/cfm_b200.h)."""
    _req(a, "gemm_ln.a", contiguous=False)
    _req(w, "gemm_ln.w", a.dtype)
    _req(x, "gemm_ln.x", torch.float32, contiguous=False)
    _req(y, "gemm_ln.y", a.dtype, contiguous=False)
    M, K = a.shape
    Nn = x.shape[1]
    ensure_init(a)
    N.check(N.lib().cfm_gemm_ln(a.data_ptr(), a.stride(0), w.data_ptr(), _ptr(bias), x.data_ptr(), x.stride(0), M, Nn, K,
                                _DT[a.dtype], float(alpha), _ptr(row_valid), g1.data_ptr(), b1.data_ptr(), _ptr(g2),
                                _ptr(b2), y.data_ptr(), y.stride(0), _ptr(y_row_valid), float(eps), engine, _stream(a)))


def ffn(y_in, w1, b1, w2, b2, x, *, alpha, ln=None, hidden_ws=None, engine=N.ENGINE_AUTO):
    """x += alpha*(w2 silu(w1 y_in + b1) + b2) (+ fused LayerNorms, ln = dict(y,g1,b1,g2,b2,y_row_valid) or None);
    see cfm_ffn in include/cfm_b200.h."""
    _req(y_in, "ffn.y_in", contiguous=False)
    _req(w1, "ffn.w1", y_in.dtype)
    _req(w2, "ffn.w2", y_in.dtype)
    _req(x, "ffn.x", torch.float32, contiguous=False)
    M, d = y_in.shape
    F = w1.shape[0]
    ln = ln or {}
    yo = ln.get("y")
    ensure_init(x)
    N.check(N.lib().cfm_ffn(y_in.data_ptr(), y_in.stride(0), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                            x.data_ptr(), x.stride(0), M, d, F, _DT[y_in.dtype], float(alpha), _ptr(ln.get("g1")),
                            _ptr(ln.get("b1")), _ptr(ln.get("g2")), _ptr(ln.get("b2")), _ptr(yo),
                            yo.stride(0) if yo is not None else 0, _ptr(ln.get("y_row_valid")), 1e-5, _ptr(hidden_ws),
                            engine, _stream(x)))


def conv_module(y_in, w1, b1, dw_w, dw_b, w2, b2, x, B, T, *, row_valid=None, ln=None, glu_ws=None, dw_ws=None,
                engine=N.ENGINE_AUTO):
    """x += rowmask(w2 silu(dw(glu(w1 y_in + b1))) + b2) (+ fused LayerNorm, ln = dict(y,g1,b1) or None) with folded
    BatchNorm; see cfm_conv_module in include/cfm_b200.h."""
    _req(y_in, "conv_module.y_in")
    _req(w1, "conv_module.w1", y_in.dtype)
    _req(w2, "conv_module.w2", y_in.dtype)
    _req(dw_w, "conv_module.dw_w", torch.float32)
    _req(dw_b, "conv_module.dw_b", torch.float32)
    _req(x, "conv_module.x", torch.float32)
    n, d = y_in.shape
    k = dw_w.shape[0]
    if n != B * T or x.shape != (n, d) or w1.shape != (2 * d, d) or w2.shape != (d, d) or dw_w.shape != (k, d):
        raise RuntimeError("conv_module: shape mismatch")
    ln = ln or {}
    yo = ln.get("y")
    if yo is not None:
        _req(yo, "conv_module.y", y_in.dtype)
    ensure_init(x)
    N.check(N.lib().cfm_conv_module(y_in.data_ptr(), w1.data_ptr(), b1.data_ptr(), dw_w.data_ptr(), dw_b.data_ptr(),
                                    w2.data_ptr(), b2.data_ptr(), x.data_ptr(), B, T, d, k, _DT[y_in.dtype],
                                    _ptr(row_valid), _ptr(ln.get("g1")), _ptr(ln.get("b1")), _ptr(yo), 1e-5,
                                    _ptr(glu_ws), _ptr(dw_ws), engine, _stream(x)))


def ffn_chain(y_in, a, b, x, y_out, *, y_row_valid=None, proj=None, hidden_ws=None, engine=N.ENGINE_AUTO):
    """Up to two feed-forward modules back to back on the same rows (+ a projection of the final LayerNorm output); see
    cfm_ffn_chain.  a (or None), b: dicts with w1, b1, w2, b2, alpha, g1, be1 and optionally g2, be2; module a's
    LayerNorm output is module b's input.  proj = (w (Np,d), bias (Np), out (M,Np)) or None; with a projection on the
    fused path y_out is NOT written."""
    _req(y_in, "ffn_chain.y_in")
    _req(x, "ffn_chain.x", torch.float32)
    _req(y_out, "ffn_chain.y_out", y_in.dtype)
    M, d = y_in.shape
    F = b["w1"].shape[0]
    for m in (a, b):
        if m is None:
            continue
        _req(m["w1"], "ffn_chain.w1", y_in.dtype)
        _req(m["w2"], "ffn_chain.w2", y_in.dtype)
        if m["w1"].shape != (F, d) or m["w2"].shape != (d, F):
            raise RuntimeError("ffn_chain: both modules must have the same (F, d)")
    if x.shape != (M, d) or y_out.shape != (M, d):
        raise RuntimeError("ffn_chain: shape mismatch")

    def mod(m):
        if m is None:
            return [None, None, None, None, 0.0, None, None, None, None]
        return [m["w1"].data_ptr(), m["b1"].data_ptr(), m["w2"].data_ptr(), m["b2"].data_ptr(), float(m["alpha"]),
                _ptr(m.get("g1")), _ptr(m.get("be1")), _ptr(m.get("g2")), _ptr(m.get("be2"))]
    pw = pb = po = None
    npj = 0
    if proj is not None:
        pw, pb, po = proj
        _req(pw, "ffn_chain.proj.w", y_in.dtype)
        _req(pb, "ffn_chain.proj.bias", torch.float32)
        _req(po, "ffn_chain.proj.out", y_in.dtype)
        npj = pw.shape[0]
        if pw.shape != (npj, d) or po.shape != (M, npj) or pb.shape != (npj,):
            raise RuntimeError("ffn_chain: projection shape mismatch")
    ensure_init(x)
    N.check(N.lib().cfm_ffn_chain(y_in.data_ptr(), M, d, F, _DT[y_in.dtype], *mod(a), *mod(b), x.data_ptr(), y_out.data_ptr(),
                                  _ptr(y_row_valid), _ptr(pw), _ptr(pb), _ptr(po), npj, 1e-5, _ptr(hidden_ws), engine,
                                  _stream(x)))


def ctc_argmax(x, w, bias, *, want_best=False, engine=N.ENGINE_AUTO):
    """Frame-wise argmax of x w^T + bias over the vocabulary (see cfm_ctc_argmax): x (M,d), w (V,d) same dtype, bias (V)
    fp32 or None.  Returns ids (M,) int32 [, best (M,) fp32]."""
    _req(x, "ctc_argmax.x")
    _req(w, "ctc_argmax.w", x.dtype)
    M, d = x.shape
    V = w.shape[0]
    if w.shape != (V, d):
        raise RuntimeError("ctc_argmax: shape mismatch")
    if bias is not None:
        _req(bias, "ctc_argmax.bias", torch.float32)
    ensure_init(x)
    ids = torch.empty(M, dtype=torch.int32, device=x.device)
    best = torch.empty(M, dtype=torch.float32, device=x.device) if want_best else None
    ws = torch.empty(max(int(N.lib().cfm_ctc_ws_bytes(M, V, _DT[x.dtype])), 8), dtype=torch.uint8, device=x.device)
    N.check(N.lib().cfm_ctc_argmax(x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(bias), M, V, d, _DT[x.dtype], ids.data_ptr(),
                                   _ptr(best), ws.data_ptr(), engine, _stream(x)))
    return (ids, best) if want_best else ids


def attention(q, k, v, out, *, mask=None, key_bias=None, scale, engine=N.ENGINE_AUTO):
    """q (B,Tq,H,64), k/v (B,Tk,H,64) views with contiguous (H,64) tail; out (B,Tq,H*64) contiguous.
    mask: uint8/bool (Bm,R,Tk) with Bm in {1,B}, R in {1,Tq}; None = unmasked."""
    _req(q, "attention.q", contiguous=False)
    B, Tq, H, dk = q.shape
    Tk = k.shape[1]
    if dk != 64:
        raise RuntimeError(f"attention: head dim {dk} unsupported (kernels are specialised for d_k = 64)")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if t.stride(3) != 1 or t.stride(2) != 64 or t.dtype != q.dtype:
            raise RuntimeError(f"attention.{n}: need (H,64) contiguous tail and a common dtype")
    _req(out, "attention.out", q.dtype)
    mbs = mrs = 0
    if mask is not None:
        _req(mask, "attention.mask", contiguous=False)
        if mask.dtype not in (torch.uint8, torch.bool) or mask.stride(2) != 1 or mask.shape[2] != Tk:
            raise RuntimeError("attention.mask: need uint8/bool with contiguous key axis of length Tk")
        mbs = 0 if mask.shape[0] == 1 else mask.stride(0)
        mrs = 0 if mask.shape[1] == 1 else mask.stride(1)
    ensure_init(q)
    N.check(N.lib().cfm_attention(q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1),
                                  v.data_ptr(), v.stride(0), v.stride(1), out.data_ptr(), B, H, Tq, Tk,
                                  _ptr(mask), mbs, mrs, _ptr(key_bias), float(scale), _DT[q.dtype], engine,
                                  _stream(q)))
    return out


def mhsa_out(q, k, v, wo, bo, x, *, mask=None, key_bias=None, scale, ln=None, ctx_ws=None, engine=N.ENGINE_AUTO):
    """x += wo . attention(q, k, v, mask) + bo (+ fused LayerNorm, ln = dict(y, g1, b1, y_row_valid) or None);
    q (B,Tq,H,64), k/v (B,Tk,H,64) views as in ``attention``; x (B*Tq, H*64) fp32; see cfm_mhsa_out."""
    _req(q, "mhsa_out.q", contiguous=False)
    B, Tq, H, dk = q.shape
    Tk = k.shape[1]
    if dk != 64:
        raise RuntimeError(f"mhsa_out: head dim {dk} unsupported (kernels are specialised for d_k = 64)")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if t.stride(3) != 1 or t.stride(2) != 64 or t.dtype != q.dtype:
            raise RuntimeError(f"mhsa_out.{n}: need (H,64) contiguous tail and a common dtype")
    d = H * 64
    _req(wo, "mhsa_out.wo", q.dtype)
    _req(x, "mhsa_out.x", torch.float32)
    if x.shape != (B * Tq, d) or wo.shape != (d, d):
        raise RuntimeError("mhsa_out: shape mismatch")
    mbs = mrs = 0
    if mask is not None:
        _req(mask, "mhsa_out.mask", contiguous=False)
        if mask.dtype not in (torch.uint8, torch.bool) or mask.stride(2) != 1 or mask.shape[2] != Tk:
            raise RuntimeError("mhsa_out.mask: need uint8/bool with contiguous key axis of length Tk")
        mbs = 0 if mask.shape[0] == 1 else mask.stride(0)
        mrs = 0 if mask.shape[1] == 1 else mask.stride(1)
    ln = ln or {}
    yo = ln.get("y")
    if yo is not None:
        _req(yo, "mhsa_out.y", q.dtype)
    if ctx_ws is not None:
        _req(ctx_ws, "mhsa_out.ctx_ws", q.dtype)
    ensure_init(q)
    N.check(N.lib().cfm_mhsa_out(q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1),
                                 v.data_ptr(), v.stride(0), v.stride(1), B, H, Tq, Tk, _ptr(mask), mbs, mrs,
                                 _ptr(key_bias), float(scale), wo.data_ptr(), bo.data_ptr(), x.data_ptr(), _DT[q.dtype],
                                 _ptr(ln.get("g1")), _ptr(ln.get("b1")), _ptr(yo), _ptr(ln.get("y_row_valid")), 1e-5,
                                 _ptr(ctx_ws), engine, _stream(q)))


def relpos_keys(k, p, u, vb, k_out, key_bias):
    """k (B,Tk,H,64) view, p (Bp,Tk,H*64) contiguous with Bp in {1,B}; see cfm_relpos_keys."""
    B, Tk, H, _ = k.shape
    _req(p, "relpos_keys.p", k.dtype)
    p_bs = 0 if p.shape[0] == 1 else p.stride(0)
    ensure_init(k)
    N.check(N.lib().cfm_relpos_keys(k.data_ptr(), k.stride(0), k.stride(1), p.data_ptr(), p_bs, u.data_ptr(),
                                    vb.data_ptr(), k_out.data_ptr(), key_bias.data_ptr(), B, H, Tk, _DT[k.dtype],
                                    _stream(k)))


def dwconv(x, w, bias, y, *, apply_silu=True):
    """x (B,T,d) act dtype contiguous; w (k,d) fp32; bias (d) fp32; y (B,T,d) (fp32 if not apply_silu)."""
    _req(x, "dwconv.x")
    _req(y, "dwconv.y")
    B, T, d = x.shape
    ensure_init(x)
    N.check(N.lib().cfm_dwconv(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, T, d, w.shape[0],
                               _DT[x.dtype], 1 if apply_silu else 0, _stream(x)))
    return y


def bn_stats(x, s, q):
    _req(x, "bn_stats.x", torch.float32)
    rows, d = x.shape
    N.check(N.lib().cfm_bn_stats(x.data_ptr(), rows, d, s.data_ptr(), q.data_ptr(), _stream(x)))


def bn_apply_silu(x, mean, rstd, gamma, beta, y):
    _req(x, "bn_apply_silu.x", torch.float32)
    rows, d = x.shape
    N.check(N.lib().cfm_bn_apply_silu(x.data_ptr(), rows, d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                      beta.data_ptr(), y.data_ptr(), _DT[y.dtype], _stream(x)))
    return y


def subsample_conv(x, w1, b1, w2, b2, ws, out):
    """x (B,Tin,idim) fp32 -> out (B,T2,F2,C) bf16 = relu(conv2(relu(conv1(x)))); see cfm_subsample_conv."""
    _req(x, "subsample_conv.x", torch.float32)
    _req(w2, "subsample_conv.w2", torch.bfloat16)
    _req(out, "subsample_conv.out", torch.bfloat16)
    B, Tin, idim = x.shape
    C = w1.shape[0]
    ensure_init(x)
    need = N.lib().cfm_subsample_ws_bytes(B, Tin, idim, C)
    if ws.numel() * ws.element_size() < need:
        raise RuntimeError("subsample_conv: workspace too small")
    N.check(N.lib().cfm_subsample_conv(x.data_ptr(), B, Tin, idim, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                       b2.data_ptr(), C, ws.data_ptr(), out.data_ptr(), _stream(x)))
    return out


def subsample_ws_bytes(B, Tin, idim, C):
    return int(N.lib().cfm_subsample_ws_bytes(B, Tin, idim, C))


def _operand(t, name):
    """(nB, nH, MN, K) logical view -> (major flag, ld, head stride, batch stride); 2-D / 3-D views are unsqueezed."""
    while t.dim() < 4:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise RuntimeError(f"gemm_ex.{name}: expected a 2-D, 3-D or 4-D view")
    if t.stride(3) == 1 and (t.stride(2) >= t.shape[3] or t.shape[2] == 1):
        mn, ld = 0, t.stride(2) if t.shape[2] > 1 else max(t.shape[3], 1)
    elif t.stride(2) == 1:
        mn, ld = 1, t.stride(3) if t.shape[3] > 1 else max(t.shape[2], 1)
    else:
        raise RuntimeError(f"gemm_ex.{name}: one of the last two dimensions must have unit stride")
    return t, mn, ld, t.stride(1), t.stride(0)


def gemm_ex(a, b, c, *, alpha=1.0, accumulate=False, splits=0, engine=N.ENGINE_AUTO):
    """c (+)= alpha * a @ b^T over the last two dims, batched over up to two leading dims (see cfm_gemm_ex).
    a: (..., M, K) view, b: (..., N, K) view, c: (..., M, N) with unit stride along N.  Transposed operands are passed as
    transposed VIEWS (x.t(), x.transpose(-1, -2)): nothing is copied, the kernel reads them MN-major."""
    _req(a, "gemm_ex.a", contiguous=False)
    _req(b, "gemm_ex.b", a.dtype, contiguous=False)
    _req(c, "gemm_ex.c", contiguous=False)
    a4, a_mn, lda, a_hs, a_bs = _operand(a, "a")
    b4, b_mn, ldb, b_hs, b_bs = _operand(b, "b")
    c4 = c
    while c4.dim() < 4:
        c4 = c4.unsqueeze(0)
    nB, nH, M, K = a4.shape
    Nn = b4.shape[2]
    if b4.shape != (nB, nH, Nn, K) or c4.shape != (nB, nH, M, Nn):
        raise RuntimeError(f"gemm_ex: shape mismatch a{tuple(a4.shape)} b{tuple(b4.shape)} c{tuple(c4.shape)}")
    if Nn > 1 and c4.stride(3) != 1:
        raise RuntimeError("gemm_ex.c: the last dimension must have unit stride")
    if accumulate and c.dtype != torch.float32:
        raise RuntimeError("gemm_ex: accumulation needs an fp32 output")
    ensure_init(a)
    N.check(N.lib().cfm_gemm_ex(a4.data_ptr(), a_mn, lda, a_hs, a_bs, b4.data_ptr(), b_mn, ldb, b_hs, b_bs, c4.data_ptr(),
                                _DT[c.dtype], c4.stride(2) if M > 1 else max(Nn, 1), c4.stride(1), c4.stride(0),
                                1 if accumulate else 0, M, Nn, K, nH, nB, _DT[a.dtype], float(alpha), int(splits), engine,
                                _stream(a)))
    return c


def l2_prefetch(tensors, stream, blocks=16):
    """Warm up to 8 CUDA tensors into L2 from ``stream`` in one launch (see cfm_l2_prefetch_multi)."""
    import ctypes
    ts = list(tensors)[:8]
    ptrs = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    sizes = (ctypes.c_int64 * len(ts))(*[t.numel() * t.element_size() for t in ts])
    N.check(N.lib().cfm_l2_prefetch_multi(ptrs, sizes, len(ts), blocks, stream.cuda_stream))
