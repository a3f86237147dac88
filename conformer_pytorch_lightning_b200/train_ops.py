"""Tensor-level wrappers over the training-path entries of the C-ABI (include/cfm_b200.h, "TRAINING path").
Same conventions as ops.py: data_ptr()s + the calling thread's current CUDA stream, CPU tensors raise."""
from __future__ import annotations

import torch

from . import _native as N
from .ops import _DT, _ptr, _req, _stream, ensure_init


def _seed_ptr(seed, p):
    """Dropout seeds live on the device (one int64 / uint64 element) so that CUDA graphs can replay with new seeds."""
    if p <= 0.0 or seed is None:
        return 0
    if not (isinstance(seed, torch.Tensor) and seed.is_cuda and seed.dtype in (torch.int64, torch.uint64) and seed.numel() >= 1):
        raise RuntimeError("dropout seed must be a CUDA int64 tensor with one element")
    return seed.data_ptr()


def ln_fwd(x, g, b, y, mean, rstd, *, row_valid=None, eps=1e-5):
    """y = rowmask(LN(x)); mean / rstd (rows,) fp32 saved for the backward."""
    _req(x, "ln_fwd.x", torch.float32)
    _req(y, "ln_fwd.y")
    rows, d = x.shape
    ensure_init(x)
    N.check(N.lib().cfm_ln_fwd_train(x.data_ptr(), rows, d, g.data_ptr(), b.data_ptr(), y.data_ptr(), _DT[y.dtype],
                                     mean.data_ptr(), rstd.data_ptr(), _ptr(row_valid), float(eps), _stream(x)))
    return y


def ln_bwd(dy, x, mean, rstd, g, dx_out, dg, db, *, dx_in=None, row_valid=None):
    """dx_out = dx_in + LN-backward(rowmask(dy)); dg, db accumulated."""
    _req(dy, "ln_bwd.dy")
    _req(x, "ln_bwd.x", torch.float32)
    _req(dx_out, "ln_bwd.dx_out", torch.float32)
    rows, d = x.shape
    N.check(N.lib().cfm_ln_bwd(dy.data_ptr(), _DT[dy.dtype], x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), g.data_ptr(),
                               _ptr(row_valid), _ptr(dx_in), dx_out.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, d,
                               _stream(x)))
    return dx_out


def silu_dropout_fwd(h, a, *, p=0.0, seed=0, site=0):
    _req(h, "silu_dropout_fwd.h")
    _req(a, "silu_dropout_fwd.a", h.dtype)
    rows, cols = h.shape
    ensure_init(h)
    N.check(N.lib().cfm_silu_dropout_fwd(h.data_ptr(), a.data_ptr(), rows, cols, _DT[h.dtype], float(p), _seed_ptr(seed, p), int(site),
                                         _stream(h)))
    return a


def silu_dropout_bwd(da, h, dh, dbias, *, p=0.0, seed=0, site=0):
    _req(da, "silu_dropout_bwd.da")
    _req(h, "silu_dropout_bwd.h", da.dtype)
    rows, cols = h.shape
    N.check(N.lib().cfm_silu_dropout_bwd(da.data_ptr(), h.data_ptr(), dh.data_ptr(), _ptr(dbias), rows, cols, _DT[h.dtype],
                                         float(p), _seed_ptr(seed, p), int(site), _stream(h)))
    return dh


def resid_dropout_add(x, f, *, x_in=None, alpha=1.0, row_valid=None, p=0.0, seed=0, site=0):
    """x = (x_in if given else x) + alpha * rowmask * dropout(f)."""
    _req(x, "resid_dropout_add.x", torch.float32)
    _req(f, "resid_dropout_add.f")
    rows, cols = x.shape
    ensure_init(x)
    N.check(N.lib().cfm_resid_dropout_add(_ptr(x_in), x.data_ptr(), f.data_ptr(), rows, cols, _DT[f.dtype], float(alpha), _ptr(row_valid),
                                          float(p), _seed_ptr(seed, p), int(site), _stream(x)))
    return x


def scale_dropout_bwd(dx, df, dbias, *, alpha=1.0, row_valid=None, p=0.0, seed=0, site=0):
    _req(dx, "scale_dropout_bwd.dx", torch.float32)
    _req(df, "scale_dropout_bwd.df")
    rows, cols = dx.shape
    N.check(N.lib().cfm_scale_dropout_bwd(dx.data_ptr(), df.data_ptr(), _ptr(dbias), rows, cols, _DT[df.dtype], float(alpha),
                                          _ptr(row_valid), float(p), _seed_ptr(seed, p), int(site), _stream(dx)))
    return df


def glu_fwd(g, u):
    _req(g, "glu_fwd.g")
    _req(u, "glu_fwd.u", g.dtype)
    rows, d = u.shape
    ensure_init(g)
    N.check(N.lib().cfm_glu_fwd(g.data_ptr(), u.data_ptr(), rows, d, _DT[g.dtype], _stream(g)))
    return u


def glu_bwd(du, g, dg, dbias):
    _req(du, "glu_bwd.du", torch.float32)
    _req(g, "glu_bwd.g")
    _req(dg, "glu_bwd.dg", g.dtype)
    rows, d = du.shape
    N.check(N.lib().cfm_glu_bwd(du.data_ptr(), g.data_ptr(), dg.data_ptr(), _ptr(dbias), rows, d, _DT[g.dtype], _stream(g)))
    return dg


def bn_silu_bwd(dc, raw, mean, rstd, gamma, beta, sums, draw, *, batch_stats=True):
    _req(dc, "bn_silu_bwd.dc")
    _req(raw, "bn_silu_bwd.raw", torch.float32)
    _req(draw, "bn_silu_bwd.draw", dc.dtype)
    rows, d = raw.shape
    N.check(N.lib().cfm_bn_silu_bwd(dc.data_ptr(), raw.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                    beta.data_ptr(), sums.data_ptr(), draw.data_ptr(), rows, d, _DT[dc.dtype],
                                    1 if batch_stats else 0, _stream(dc)))
    return draw


def dwconv_wgrad(dy, u, dw, dbias):
    """dy, u (B,T,d) act dtype; dw (k,d), dbias (d) fp32 accumulated."""
    _req(dy, "dwconv_wgrad.dy")
    _req(u, "dwconv_wgrad.u", dy.dtype)
    B, T, d = dy.shape
    N.check(N.lib().cfm_dwconv_wgrad(dy.data_ptr(), u.data_ptr(), dw.data_ptr(), dbias.data_ptr(), B, T, d, dw.shape[0],
                                     _DT[dy.dtype], _stream(dy)))


def softmax_fwd(S, P, Pd, mask, *, Tk, p=0.0, seed=0, site=0):
    """S (B,H,Tq,Tp) fp32 -> P (and Pd when p > 0), same shape in the activation dtype.  mask uint8/bool (Bm,R,Tk)."""
    _req(S, "softmax_fwd.S", torch.float32)
    _req(P, "softmax_fwd.P")
    B, H, Tq, Tp = S.shape
    mbs = mrs = 0
    if mask is not None:
        mbs = 0 if mask.shape[0] == 1 else mask.stride(0)
        mrs = 0 if mask.shape[1] == 1 else mask.stride(1)
    ensure_init(S)
    N.check(N.lib().cfm_softmax_fwd(S.data_ptr(), P.data_ptr(), _ptr(Pd), _ptr(mask), mbs, mrs, B, H, Tq, Tk, Tp, _DT[P.dtype],
                                    float(p), _seed_ptr(seed, p), int(site), _stream(S)))


def softmax_bwd(P, dPd, dS, *, Tk, p=0.0, seed=0, site=0):
    _req(P, "softmax_bwd.P")
    _req(dPd, "softmax_bwd.dPd", torch.float32)
    _req(dS, "softmax_bwd.dS", P.dtype)
    B, H, Tq, Tp = P.shape
    N.check(N.lib().cfm_softmax_bwd(P.data_ptr(), dPd.data_ptr(), dS.data_ptr(), B, H, Tq, Tk, Tp, _DT[P.dtype], float(p),
                                    _seed_ptr(seed, p), int(site), _stream(P)))


def colsum(x, out):
    """out (cols,) fp32 += column sums of x (rows, cols) (row stride may exceed cols)."""
    _req(x, "colsum.x", contiguous=False)
    rows, cols = x.shape
    ensure_init(x)
    N.check(N.lib().cfm_colsum(x.data_ptr(), x.stride(0), out.data_ptr(), rows, cols, _DT[x.dtype], _stream(x)))
    return out


def ctc_loss_ws(B, T, Lmax, device):
    return torch.empty(int(N.lib().cfm_ctc_loss_ws_bytes(B, T, Lmax)), dtype=torch.uint8, device=device)


def ctc_loss_fwd(logits, B, T, V, labels, in_len, lab_len, nll, ws):
    _req(logits, "ctc_loss_fwd.logits", contiguous=False)
    _req(labels, "ctc_loss_fwd.labels", torch.int32)
    ensure_init(logits)
    N.check(N.lib().cfm_ctc_loss_fwd(logits.data_ptr(), logits.stride(0), B, T, V, labels.data_ptr(), labels.shape[1],
                                     in_len.data_ptr(), lab_len.data_ptr(), nll.data_ptr(), ws.data_ptr(), _DT[logits.dtype],
                                     _stream(logits)))
    return nll


def ctc_loss_bwd(logits, B, T, V, labels, in_len, lab_len, nll, ws, scale, dlogits):
    N.check(N.lib().cfm_ctc_loss_bwd(logits.data_ptr(), logits.stride(0), B, T, V, logits.shape[1], labels.data_ptr(),
                                     labels.shape[1], in_len.data_ptr(), lab_len.data_ptr(), nll.data_ptr(), ws.data_ptr(),
                                     float(scale), dlogits.data_ptr(), _DT[logits.dtype], _stream(logits)))
    return dlogits
