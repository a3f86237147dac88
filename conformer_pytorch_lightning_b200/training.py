"""Training path of the encoder layer stack: forward that saves what the backward needs, the hand-scheduled backward, and
the torch.autograd.Function that plugs both into PyTorch's autograd (reference: module.py:49-69 drives
loss.backward() through encoder.py:72-74 / encoder_layer.py:49-71 and the modules they call).

Compute dtype bf16 (tcgen05 GEMMs: forward on gemm_tc, dgrad / wgrad / attention products on the general transposed-
operand GEMM of gemm_gen.cu) or fp32 (CUDA-core engines, the gradient-parity path).  The residual stream, LayerNorm /
BatchNorm statistics, and every parameter gradient are fp32.  Dropout (feedforward.py:19, attention.py:95,
encoder_layer.py:58,62,66,69) is counter based: the backward regenerates the masks from (seed, site).

Per layer the backward finishes with ALL gradients of that layer in one flat fp32 bucket; ``grad_sync`` (ddp.GradSync) is
called with the bucket as soon as the layer's kernels are enqueued, so the NCCL all-reduce of layer i overlaps the
backward of layers i-1, ....
"""
from __future__ import annotations

import math
import weakref

import torch

from . import _native as N
from . import engine, ops
from . import train_ops as TO

_EPS = 1e-5
SITES_PER_LAYER = 8
(S_FFM_IN, S_FFM_OUT, S_ATT_P, S_ATT_OUT, S_CONV_OUT, S_FF_IN, S_FF_OUT) = range(7)


# ----------------------------------------------------------------------------------------------- parameter plumbing
def layer_param_list(layer):
    """Parameters of one ConformerEncoderLayer in the order the backward returns their gradients."""
    ps = []
    for ff in (layer.feed_forward_macaron, layer.feed_forward):
        ps += [ff.w_1.weight, ff.w_1.bias, ff.w_2.weight, ff.w_2.bias]
    a = layer.self_attn
    ps += [a.linear_q.weight, a.linear_q.bias, a.linear_k.weight, a.linear_k.bias, a.linear_v.weight, a.linear_v.bias,
           a.linear_out.weight, a.linear_out.bias]
    if hasattr(a, "pos_bias_u"):
        ps += [a.pos_bias_u, a.pos_bias_v, a.linear_pos.weight]
    c = layer.conv_module
    ps += [c.pointwise_conv1.weight, c.pointwise_conv1.bias, c.depthwise_conv.weight, c.depthwise_conv.bias,
           c.norm.weight, c.norm.bias, c.pointwise_conv2.weight, c.pointwise_conv2.bias]
    for ln in (layer.norm_ff_macaron, layer.norm_mha, layer.norm_conv, layer.norm_ff, layer.norm_final):
        ps += [ln.weight, ln.bias]
    return ps


class _Bucket:
    """One flat fp32 gradient buffer per layer with named views (derived-weight layout)."""

    def __init__(self, d, F, k, device):
        spec = []
        for ff in ("ffm", "ff"):
            spec += [(ff + "_w1", (F, d)), (ff + "_b1", (F,)), (ff + "_w2", (d, F)), (ff + "_b2", (d,))]
        spec += [("wqkv", (3 * d, d)), ("bqkv", (3 * d,)), ("wo", (d, d)), ("bo", (d,)),
                 ("pw1", (2 * d, d)), ("pw1_b", (2 * d,)), ("dw", (k, d)), ("dw_b", (d,)), ("bn_g", (d,)), ("bn_b", (d,)),
                 ("pw2", (d, d)), ("pw2_b", (d,))]
        for ln in ("ffm", "mha", "conv", "ff", "fin"):
            spec += [("ln_" + ln + "_g", (d,)), ("ln_" + ln + "_b", (d,))]
        total = sum(math.prod(s) for _, s in spec)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.v = {}
        off = 0
        for name, shape in spec:
            n = math.prod(shape)
            self.v[name] = self.flat[off:off + n].view(shape)
            off += n

    def grads_for(self, layer, d):
        """Views / reshapes matching layer_param_list(layer)."""
        v = self.v
        out = []
        for ff in ("ffm", "ff"):
            out += [v[ff + "_w1"], v[ff + "_b1"], v[ff + "_w2"], v[ff + "_b2"]]
        a = layer.self_attn
        H = a.num_heads
        out += [v["wqkv"][:d], v["bqkv"][:d], v["wqkv"][d:2 * d], v["bqkv"][d:2 * d], v["wqkv"][2 * d:], v["bqkv"][2 * d:],
                v["wo"], v["bo"]]
        if hasattr(a, "pos_bias_u"):
            # (q + u) is folded into the q bias: d(pos_bias_u) = d(q bias).  With one position row per batch element
            # (batched forward) the matrix_bd term is constant along keys, the softmax cancels it and the
            # gradients of pos_bias_v / linear_pos are exactly zero (the reference's autograd leaves ~1e-7 of rounding).
            out += [v["bqkv"][:d].view(H, d // H), torch.zeros_like(a.pos_bias_v), torch.zeros_like(a.linear_pos.weight)]
        k = v["dw"].shape[0]
        out += [v["pw1"].view(2 * d, d, 1), v["pw1_b"], v["dw"].t().reshape(d, 1, k), v["dw_b"], v["bn_g"], v["bn_b"],
                v["pw2"].view(d, d, 1), v["pw2_b"]]
        for ln in ("ffm", "mha", "conv", "ff", "fin"):
            out += [v["ln_" + ln + "_g"], v["ln_" + ln + "_b"]]
        return out


class TrainRun:
    """Everything one forward/backward pair of the stack shares."""

    def __init__(self, layers, after_norm, attn_mask, pad_mask, B, T, dtype, seed, grad_sync=None):
        self.layers, self.after_norm = layers, after_norm
        self.attn_mask = engine._mask_u8(attn_mask)
        self.row_valid = engine._row_valid(pad_mask, B, T)
        self.B, self.T, self.dtype, self.seed = B, T, dtype, seed        # seed: device int64 tensor, one element
        self.grad_sync = grad_sync
        self.saved = None


def _drop_p(module, p):
    return float(p) if module.training else 0.0


# ----------------------------------------------------------------------------------------------- forward
def _ffn_fwd(x_in, W, ln_g, ln_b, n, d, dtype, p_in, p_out, seed, site_in, site_out, sv, tag):
    dev = x_in.device
    y = torch.empty((n, d), dtype=dtype, device=dev)
    mean, rstd = torch.empty(n, device=dev), torch.empty(n, device=dev)
    TO.ln_fwd(x_in, ln_g, ln_b, y, mean, rstd)
    F = W["w1"].shape[0]
    h = torch.empty((n, F), dtype=dtype, device=dev)
    ops.gemm(y, W["w1"], W["b1"], h, N.EPI_BIAS)
    a = torch.empty((n, F), dtype=dtype, device=dev)
    TO.silu_dropout_fwd(h, a, p=p_in, seed=seed, site=site_in)
    x_out = torch.empty_like(x_in)
    if p_out == 0.0:
        ops.gemm(a, W["w2"], W["b2"], x_out, N.EPI_RESIDUAL, residual=x_in, alpha=0.5)
    else:
        f = torch.empty((n, d), dtype=dtype, device=dev)
        ops.gemm(a, W["w2"], W["b2"], f, N.EPI_BIAS)
        TO.resid_dropout_add(x_out, f, x_in=x_in, alpha=0.5, p=p_out, seed=seed, site=site_out)
    sv[tag] = dict(x=x_in, y=y, mean=mean, rstd=rstd, h=h, a=a)
    return x_out


def _mhsa_fwd(x_in, W, ln_g, ln_b, run, H, n, d, p_att, p_out, site_p, site_out, sv):
    dev, dtype, B, T = x_in.device, run.dtype, run.B, run.T
    y = torch.empty((n, d), dtype=dtype, device=dev)
    mean, rstd = torch.empty(n, device=dev), torch.empty(n, device=dev)
    TO.ln_fwd(x_in, ln_g, ln_b, y, mean, rstd)
    qkv = torch.empty((n, 3 * d), dtype=dtype, device=dev)
    ops.gemm(y, W["wqkv"], W["bqkv"], qkv, N.EPI_BIAS)
    q5 = qkv.view(B, T, 3, H, 64)
    q, k, v = (q5[:, :, i].permute(0, 2, 1, 3) for i in range(3))                 # (B, H, T, 64) strided views
    Tp = (T + 7) // 8 * 8
    ws = engine.thread_workspace()
    S = ws.get("train_scores", (B, H, T, Tp), torch.float32, dev)
    ops.gemm_ex(q, k, S[..., :T], alpha=1.0 / math.sqrt(64.0))                     # attention.py:84,88 (bd term: see D2)
    P = torch.empty((B, H, T, Tp), dtype=dtype, device=dev)
    Pd = torch.empty_like(P) if p_att > 0.0 else None
    TO.softmax_fwd(S, P, Pd, run.attn_mask, Tk=T, p=p_att, seed=run.seed, site=site_p)      # attention.py:89-95
    ctx = torch.empty((n, d), dtype=dtype, device=dev)
    pv = Pd if Pd is not None else P
    ops.gemm_ex(pv[..., :T], v.transpose(-1, -2), ctx.view(B, T, H, 64).permute(0, 2, 1, 3))    # attention.py:96-97
    x_out = torch.empty_like(x_in)
    if p_out == 0.0:
        ops.gemm(ctx, W["wo"], W["bo"], x_out, N.EPI_RESIDUAL, residual=x_in, alpha=1.0)
    else:
        f = torch.empty((n, d), dtype=dtype, device=dev)
        ops.gemm(ctx, W["wo"], W["bo"], f, N.EPI_BIAS)
        TO.resid_dropout_add(x_out, f, x_in=x_in, alpha=1.0, p=p_out, seed=run.seed, site=site_out)
    sv["mha"] = dict(x=x_in, y=y, mean=mean, rstd=rstd, qkv=qkv, P=P, Pd=Pd, ctx=ctx, Tp=Tp)
    return x_out


def _conv_fwd(x_in, W, ln_g, ln_b, module, run, n, d, p_out, site_out, sv):
    dev, dtype, B, T = x_in.device, run.dtype, run.B, run.T
    rv = run.row_valid
    y = torch.empty((n, d), dtype=dtype, device=dev)
    mean, rstd = torch.empty(n, device=dev), torch.empty(n, device=dev)
    TO.ln_fwd(x_in, ln_g, ln_b, y, mean, rstd, row_valid=rv)                       # + masked_fill of convolution.py:36-37
    g = torch.empty((n, 2 * d), dtype=dtype, device=dev)
    ops.gemm(y, W["w1"], W["b1"], g, N.EPI_BIAS)
    u = torch.empty((n, d), dtype=dtype, device=dev)
    TO.glu_fwd(g, u)
    raw = torch.empty((n, d), dtype=torch.float32, device=dev)
    bn = module.norm
    if module.training:
        ops.dwconv(u.view(B, T, d), W["dw_w"], W["dw_b"], raw.view(B, T, d), apply_silu=False)
        st = torch.zeros((2, d), dtype=torch.float32, device=dev)
        ops.bn_stats(raw, st[0], st[1])
        bmean = st[0] / n
        var = (st[1] / n - bmean * bmean).clamp_min_(0.0)
        with torch.no_grad():                                                      # running statistics (momentum 0.1)
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - mom).add_(bmean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var * (n / max(n - 1, 1)), alpha=mom)
            bn.num_batches_tracked += 1
        brstd = torch.rsqrt(var + bn.eps)
        batch_stats = True
    else:
        # eval-mode module inside a differentiated call (fine-tuning with frozen statistics): the derived filter of
        # engine.conv_weights has BatchNorm folded in, so use the raw parameters here
        dw = module.depthwise_conv.weight.detach().float().reshape(d, -1).t().contiguous()
        db = (module.depthwise_conv.bias.detach().float() if module.depthwise_conv.bias is not None
              else torch.zeros(d, device=dev))
        ops.dwconv(u.view(B, T, d), dw, db.contiguous(), raw.view(B, T, d), apply_silu=False)
        bmean = bn.running_mean.float()
        brstd = torch.rsqrt(bn.running_var.float() + bn.eps)
        batch_stats = False
    bmean, brstd = bmean.contiguous(), brstd.contiguous()
    c = torch.empty((n, d), dtype=dtype, device=dev)
    ops.bn_apply_silu(raw, bmean, brstd, W["gamma"], W["beta"], c)
    x_out = torch.empty_like(x_in)
    if p_out == 0.0:
        ops.gemm(c, W["w2"], W["b2"], x_out, N.EPI_RESIDUAL, residual=x_in, alpha=1.0, row_valid=rv)
    else:
        f = torch.empty((n, d), dtype=dtype, device=dev)
        ops.gemm(c, W["w2"], W["b2"], f, N.EPI_BIAS)
        TO.resid_dropout_add(x_out, f, x_in=x_in, alpha=1.0, row_valid=rv, p=p_out, seed=run.seed, site=site_out)
    sv["conv"] = dict(x=x_in, y=y, mean=mean, rstd=rstd, g=g, u=u, raw=raw, bmean=bmean, brstd=brstd, c=c,
                      batch_stats=batch_stats)
    return x_out


def _raw_conv_weights(layer, W, d):
    """Depthwise filter / bias WITHOUT BatchNorm folded in (training always; see _conv_fwd for eval)."""
    m = layer.conv_module
    if m.training:
        return W["dw_w"]
    return m.depthwise_conv.weight.detach().float().reshape(d, -1).t().contiguous()


def stack_forward(x_emb, run):
    """encoder.py:72-74 in training mode.  x_emb (B,T,d) fp32.  Returns out (B,T,d) fp32; run.saved holds the
    activations the backward needs."""
    B, T, d = x_emb.shape
    n = B * T
    dev = x_emb.device
    x = x_emb.reshape(n, d).float().contiguous()
    saved = []
    for li, layer in enumerate(run.layers):
        Wl = layer.derived_weights(run.dtype)
        H = layer.self_attn.num_heads
        sv = {}
        base = li * SITES_PER_LAYER
        p_l = _drop_p(layer, layer.dropout.p)
        x = _ffn_fwd(x, Wl["ffm"], Wl["ffm_g"], Wl["ffm_b"], n, d, run.dtype,
                     _drop_p(layer.feed_forward_macaron, layer.feed_forward_macaron.dropout.p), p_l, run.seed,
                     base + S_FFM_IN, base + S_FFM_OUT, sv, "ffm")
        x = _mhsa_fwd(x, Wl["mha"], Wl["mha_g"], Wl["mha_b"], run, H, n, d,
                      _drop_p(layer.self_attn, layer.self_attn.dropout.p), p_l, base + S_ATT_P, base + S_ATT_OUT, sv)
        x = _conv_fwd(x, Wl["conv"], Wl["conv_g"], Wl["conv_b"], layer.conv_module, run, n, d, p_l, base + S_CONV_OUT, sv)
        x = _ffn_fwd(x, Wl["ff"], Wl["ff_g"], Wl["ff_b"], n, d, run.dtype,
                     _drop_p(layer.feed_forward, layer.feed_forward.dropout.p), p_l, run.seed, base + S_FF_IN,
                     base + S_FF_OUT, sv, "ff")
        out = torch.empty_like(x)
        mean, rstd = torch.empty(n, device=dev), torch.empty(n, device=dev)
        TO.ln_fwd(x, Wl["fin_g"], Wl["fin_b"], out, mean, rstd)
        sv["fin"] = dict(x=x, mean=mean, rstd=rstd)
        sv["p"] = dict(layer=p_l, ffm=_drop_p(layer.feed_forward_macaron, layer.feed_forward_macaron.dropout.p),
                       ff=_drop_p(layer.feed_forward, layer.feed_forward.dropout.p),
                       att=_drop_p(layer.self_attn, layer.self_attn.dropout.p))
        saved.append(sv)
        x = out
    after = None
    if run.after_norm is not None:
        out = torch.empty_like(x)
        mean, rstd = torch.empty(n, device=dev), torch.empty(n, device=dev)
        TO.ln_fwd(x, engine._f32(run.after_norm.weight), engine._f32(run.after_norm.bias), out, mean, rstd)
        after = dict(x=x, mean=mean, rstd=rstd)
        x = out
    run.saved = (saved, after)
    return x.view(B, T, d)


# ----------------------------------------------------------------------------------------------- backward
def _linear_bwd(dy, a_in, w, gw, d_in):
    """dy (n, out) = gradient of a_in @ w.T: d_in (n, in) = dy @ w (dgrad), gw (out, in) += dy.T @ a_in (wgrad)."""
    ops.gemm_ex(dy, w.t(), d_in)
    ops.gemm_ex(dy.t(), a_in.t(), gw, accumulate=True)


def _ffn_bwd(dx, s, W, ln_g, bk, tag, n, d, dtype, p_in, p_out, seed, site_in, site_out):
    dev = dx.device
    F = W["w1"].shape[0]
    df = torch.empty((n, d), dtype=dtype, device=dev)
    TO.scale_dropout_bwd(dx, df, bk[tag + "_b2"], alpha=0.5, p=p_out, seed=seed, site=site_out)
    da = torch.empty((n, F), dtype=dtype, device=dev)
    _linear_bwd(df, s["a"], W["w2"], bk[tag + "_w2"], da)
    TO.silu_dropout_bwd(da, s["h"], da, bk[tag + "_b1"], p=p_in, seed=seed, site=site_in)
    dy = df                                                     # reuse: (n, d)
    _linear_bwd(da, s["y"], W["w1"], bk[tag + "_w1"], dy)
    lname = "ln_" + tag
    dx_out = torch.empty_like(dx)
    TO.ln_bwd(dy, s["x"], s["mean"], s["rstd"], ln_g, dx_out, bk[lname + "_g"], bk[lname + "_b"], dx_in=dx)
    return dx_out


def _mhsa_bwd(dx, s, W, ln_g, bk, run, H, n, d, p_att, p_out, site_p, site_out):
    dev, dtype, B, T = dx.device, run.dtype, run.B, run.T
    Tp = s["Tp"]
    do = torch.empty((n, d), dtype=dtype, device=dev)
    TO.scale_dropout_bwd(dx, do, bk["bo"], alpha=1.0, p=p_out, seed=run.seed, site=site_out)
    dctx = torch.empty((n, d), dtype=dtype, device=dev)
    _linear_bwd(do, s["ctx"], W["wo"], bk["wo"], dctx)
    q5 = s["qkv"].view(B, T, 3, H, 64)
    q, k, v = (q5[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    dqkv = torch.empty((n, 3 * d), dtype=dtype, device=dev)
    d5 = dqkv.view(B, T, 3, H, 64)
    dq, dk, dv = (d5[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    dc4 = dctx.view(B, T, H, 64).permute(0, 2, 1, 3)
    ws = engine.thread_workspace()
    dP = ws.get("train_scores", (B, H, T, Tp), torch.float32, dev)
    ops.gemm_ex(dc4, v, dP[..., :T])                                                    # dP = dO V^T
    pv = s["Pd"] if s["Pd"] is not None else s["P"]
    ops.gemm_ex(pv[..., :T].transpose(-1, -2), dc4.transpose(-1, -2), dv)               # dV = P^T dO
    dS = ws.get("train_ds", (B, H, T, Tp), dtype, dev)
    TO.softmax_bwd(s["P"], dP, dS, Tk=T, p=p_att, seed=run.seed, site=site_p)
    scale = 1.0 / math.sqrt(64.0)
    ops.gemm_ex(dS[..., :T], k.transpose(-1, -2), dq, alpha=scale)                      # dQ = dS K / sqrt(dk)
    ops.gemm_ex(dS[..., :T].transpose(-1, -2), q.transpose(-1, -2), dk, alpha=scale)    # dK = dS^T Q / sqrt(dk)
    TO.colsum(dqkv, bk["bqkv"])
    dy = do
    _linear_bwd(dqkv, s["y"], W["wqkv"], bk["wqkv"], dy)
    dx_out = torch.empty_like(dx)
    TO.ln_bwd(dy, s["x"], s["mean"], s["rstd"], ln_g, dx_out, bk["ln_mha_g"], bk["ln_mha_b"], dx_in=dx)
    return dx_out


def _conv_bwd(dx, s, W, ln_g, layer, bk, run, n, d, p_out, site_out):
    dev, dtype, B, T = dx.device, run.dtype, run.B, run.T
    rv = run.row_valid
    dz = torch.empty((n, d), dtype=dtype, device=dev)
    TO.scale_dropout_bwd(dx, dz, bk["pw2_b"], alpha=1.0, row_valid=rv, p=p_out, seed=run.seed, site=site_out)
    dc = torch.empty((n, d), dtype=dtype, device=dev)
    _linear_bwd(dz, s["c"], W["w2"], bk["pw2"], dc)
    sums = torch.empty(2 * d, dtype=torch.float32, device=dev)
    draw = dz                                                                             # reuse
    TO.bn_silu_bwd(dc, s["raw"], s["bmean"], s["brstd"], W["gamma"], W["beta"], sums, draw, batch_stats=s["batch_stats"])
    bk["bn_g"].add_(sums[:d])
    bk["bn_b"].add_(sums[d:])
    dw_raw = _raw_conv_weights(layer, W, d)
    du = torch.empty((n, d), dtype=torch.float32, device=dev)
    ops.dwconv(draw.view(B, T, d), dw_raw.flip(0).contiguous(), torch.zeros(d, dtype=torch.float32, device=dev),
               du.view(B, T, d), apply_silu=False)                                       # dgrad = tap-reversed filter
    TO.dwconv_wgrad(draw.view(B, T, d), s["u"].view(B, T, d), bk["dw"], bk["dw_b"])
    dg = torch.empty((n, 2 * d), dtype=dtype, device=dev)
    TO.glu_bwd(du, s["g"], dg, bk["pw1_b"])
    dy = dc
    _linear_bwd(dg, s["y"], W["w1"], bk["pw1"], dy)
    dx_out = torch.empty_like(dx)
    TO.ln_bwd(dy, s["x"], s["mean"], s["rstd"], ln_g, dx_out, bk["ln_conv_g"], bk["ln_conv_b"], dx_in=dx, row_valid=rv)
    return dx_out


def backward_head(run, dx):
    """after_norm backward (encoder.py:74).  Returns (dx, [g_weight, g_bias] or None)."""
    after = run.saved[1]
    if after is None:
        return dx, None
    d = dx.shape[-1]
    gg, gb = torch.zeros(d, device=dx.device), torch.zeros(d, device=dx.device)
    nxt = torch.empty_like(dx)
    TO.ln_bwd(dx, after["x"], after["mean"], after["rstd"], engine._f32(run.after_norm.weight), nxt, gg, gb)
    return nxt, [gg, gb]


def backward_layer(run, li, dx, release=True):
    """Backward of layer ``li``.  Returns (dx w.r.t. the layer input, the layer's flat gradient bucket)."""
    saved = run.saved[0]
    layer = run.layers[li]
    n, d = dx.shape
    dev = dx.device
    Wl = layer.derived_weights(run.dtype)
    sv = saved[li]
    p = sv["p"]
    H = layer.self_attn.num_heads
    F = Wl["ff"]["w1"].shape[0]
    k = Wl["conv"]["dw_w"].shape[0]
    bucket = _Bucket(d, F, k, dev)
    bk = bucket.v
    base = li * SITES_PER_LAYER
    nxt = torch.empty_like(dx)
    TO.ln_bwd(dx, sv["fin"]["x"], sv["fin"]["mean"], sv["fin"]["rstd"], Wl["fin_g"], nxt, bk["ln_fin_g"], bk["ln_fin_b"])
    dx = nxt
    dx = _ffn_bwd(dx, sv["ff"], Wl["ff"], Wl["ff_g"], bk, "ff", n, d, run.dtype, p["ff"], p["layer"], run.seed,
                  base + S_FF_IN, base + S_FF_OUT)
    dx = _conv_bwd(dx, sv["conv"], Wl["conv"], Wl["conv_g"], layer, bk, run, n, d, p["layer"], base + S_CONV_OUT)
    dx = _mhsa_bwd(dx, sv["mha"], Wl["mha"], Wl["mha_g"], bk, run, H, n, d, p["att"], p["layer"], base + S_ATT_P,
                   base + S_ATT_OUT)
    dx = _ffn_bwd(dx, sv["ffm"], Wl["ffm"], Wl["ffm_g"], bk, "ffm", n, d, run.dtype, p["ffm"], p["layer"], run.seed,
                  base + S_FFM_IN, base + S_FFM_OUT)
    if release:
        saved[li] = None                                          # release this layer's activations
    return dx, bucket


def stack_backward(dout, run):
    """Eager backward.  Returns (dx_emb (B,T,d) fp32, [per-layer gradient lists in layer_param_list order], after_norm
    grads)."""
    B, T = run.B, run.T
    d = dout.shape[-1]
    dx = dout.reshape(B * T, d).float().contiguous()
    dx, after_grads = backward_head(run, dx)
    layer_grads = [None] * len(run.layers)
    for li in range(len(run.layers) - 1, -1, -1):
        dx, bucket = backward_layer(run, li, dx)
        if run.grad_sync is not None:
            run.grad_sync.bucket_ready(bucket.flat)               # all-reduce overlaps the remaining layers' backward
        layer_grads[li] = bucket.grads_for(run.layers[li], d)
    if run.grad_sync is not None:
        if after_grads is not None:
            for g in after_grads:
                run.grad_sync.bucket_ready(g)
        run.grad_sync.finish()
    return dx.view(B, T, d), layer_grads, after_grads


# ----------------------------------------------------------------------------------------------- CUDA-graph plans
class _Token:
    pass


class TrainPlan:
    """A training step of the stack is ~780 small launches; issued one by one from Python the GPU idles most of the
    time.  Like the inference path, a (shape, dtype, mask layout, dropout configuration) is captured once into CUDA
    graphs -- one for the forward of the whole stack, one per layer for the backward (so that each layer's gradient
    bucket can be handed to the NCCL all-reduce between graph launches) -- and replayed afterwards.  Everything the
    graphs touch is static: inputs are copied into the plan's buffers, saved activations live in the graphs' memory
    pool, the dropout seed is a device scalar, the bf16 weight copies are re-derived INSIDE the forward graph."""

    def __init__(self):
        self.pool = None
        self.fwd = None
        self.bwd = None           # list of (graph, stage) in execution order
        self.pending = None       # weak reference to the autograd node of a forward whose backward has not run yet: it
                                  # owns the static buffers.  A node that died without a backward (a forward whose loss was
                                  # dropped) releases them.
        self.calls = 0

    @property
    def busy(self):
        return self.pending is not None and self.pending() is not None


def _plan_key(x_emb, layers, after_norm, attn_mask, pad_mask, dtype, grad_sync, params=()):
    # the captured graphs read the parameters through their addresses: a re-pointed parameter (p.data = ..., e.g. by
    # optim.FlatAdam when it moves the parameters into its flat buffers) must not replay a stale plan
    # ... and a plan captured while the bf16 mirrors of the parameters were current (no cast kernels inside) is only valid
    # while they are: any other in-place update of a parameter flips its bit and selects a plan that re-derives its weights
    where = hash(tuple((p.data_ptr(), engine.mirror_valid(p)) for p in params))
    cfg = (where,) + tuple((l.training, l.conv_module.training, l.dropout.p, l.self_attn.dropout.p, l.feed_forward.dropout.p,
                 l.feed_forward_macaron.dropout.p, l.conv_module.norm.momentum) for l in layers)
    return (tuple(x_emb.shape), x_emb.device, dtype, None if attn_mask is None else tuple(attn_mask.shape),
            None if pad_mask is None else tuple(pad_mask.shape), cfg, after_norm is not None, grad_sync is not None)


def _refresh_derived(layers, dtype):
    for layer in layers:
        for m in (layer.feed_forward_macaron, layer.feed_forward, layer.self_attn, layer.conv_module):
            m._derived.force = True
        layer.derived_weights(dtype)


def _capture_forward(plan, x_emb, layers, after_norm, attn_mask, pad_mask, dtype, grad_sync):
    dev = x_emb.device
    B, T, d = x_emb.shape
    plan.pool = torch.cuda.graph_pool_handle()
    plan.x = torch.empty((B, T, d), dtype=torch.float32, device=dev)
    am = engine._mask_u8(attn_mask)
    pm = engine._mask_u8(pad_mask)
    plan.am = None if am is None else torch.empty(am.shape, dtype=torch.bool, device=dev)
    plan.pm = None if pm is None else torch.empty(pm.shape, dtype=torch.bool, device=dev)
    plan.seed = torch.zeros(1, dtype=torch.int64, device=dev)
    _fill_plan(plan, x_emb, am, pm, None)
    plan.run = TrainRun(layers, after_norm, plan.am, plan.pm, B, T, dtype, plan.seed, grad_sync)
    g = torch.cuda.CUDAGraph()
    n0 = N.launch_count()
    with torch.cuda.graph(g, pool=plan.pool, capture_error_mode="thread_local"):
        _refresh_derived(layers, dtype)                # bf16 weight copies are re-derived at every replay
        plan.out = stack_forward(plan.x, plan.run)
    plan.fwd = g
    plan.fwd_launches = N.launch_count() - n0


def _fill_plan(plan, x_emb, am, pm, seed_host):
    plan.x.copy_(x_emb)
    if plan.am is not None:
        torch.ne(am, 0, out=plan.am)
    if plan.pm is not None:
        torch.ne(pm, 0, out=plan.pm)
    if seed_host is not None:
        plan.seed.copy_(seed_host, non_blocking=True)


def _capture_backward(plan, dout):
    run = plan.run
    B, T = run.B, run.T
    d = dout.shape[-1]
    plan.dout = torch.empty((B * T, d), dtype=torch.float32, device=dout.device)
    plan.dout.copy_(dout.reshape(B * T, d))
    stages = []
    n0 = N.launch_count()

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=plan.pool, capture_error_mode="thread_local"):
            res = fn()
        stages.append(g)
        return res
    if run.grad_sync is None:
        # one graph for the whole backward
        def whole():
            dx, after_grads = backward_head(run, plan.dout)
            buckets = [None] * len(run.layers)
            for li in range(len(run.layers) - 1, -1, -1):
                dx, buckets[li] = backward_layer(run, li, dx, release=False)
            return dx, after_grads, buckets
        plan.dx, plan.after_grads, plan.buckets = capture(whole)
        plan.stage_buckets = [None]
    else:
        dx, plan.after_grads = capture(lambda: backward_head(run, plan.dout))
        plan.stage_buckets = [None]
        plan.buckets = [None] * len(run.layers)
        for li in range(len(run.layers) - 1, -1, -1):
            dx, plan.buckets[li] = capture(lambda li=li, dx=dx: backward_layer(run, li, dx, release=False))
            plan.stage_buckets.append(plan.buckets[li])
        plan.dx = dx
    plan.bwd = stages
    plan.bwd_launches = N.launch_count() - n0


def _replay_backward(plan, dout):
    run = plan.run
    B, T = run.B, run.T
    d = dout.shape[-1]
    if plan.bwd is None:
        _capture_backward(plan, dout)
    else:
        plan.dout.copy_(dout.reshape(B * T, d))
    sync = run.grad_sync
    engine.GRAPH_REPLAYED_LAUNCHES[0] += plan.bwd_launches       # native kernels inside the replayed graphs
    for g, bucket in zip(plan.bwd, plan.stage_buckets):
        g.replay()
        if sync is not None and bucket is not None:
            sync.bucket_ready(bucket.flat)
    if sync is not None:
        if plan.after_grads is not None:
            for t in plan.after_grads:
                sync.bucket_ready(t)
        sync.finish()
    # the buckets are static graph memory: hand autograd private copies (param.grad may keep / alias what it is given)
    layer_grads = []
    for li, layer in enumerate(run.layers):
        b = plan.buckets[li]
        c = _Bucket.__new__(_Bucket)
        c.flat = b.flat.clone()
        c.v = {name: c.flat[v.storage_offset() - b.flat.storage_offset():][:v.numel()].view(v.shape) for name, v in b.v.items()}
        layer_grads.append(c.grads_for(layer, d))
    after = None if plan.after_grads is None else [t.clone() for t in plan.after_grads]
    return plan.dx.view(B, T, d).clone(), layer_grads, after


class EncoderStackFunction(torch.autograd.Function):
    """out = layers(x_emb) with the native forward / backward above.  ``params`` (the flattened parameter lists of the
    layers + after_norm) are passed so that autograd routes the returned gradients to them."""

    @staticmethod
    def forward(ctx, x_emb, run, plan, *params):
        ctx.n_params = len(params)
        ctx.plan = plan
        if plan is not None:
            plan.fwd.replay()
            engine.GRAPH_REPLAYED_LAUNCHES[0] += plan.fwd_launches
            ctx.token = _Token()                       # dies with the autograd node
            plan.pending = weakref.ref(ctx.token)
            ctx.run = plan.run
            return plan.out.clone()
        out = stack_forward(x_emb, run)
        ctx.run = run
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        run, plan = ctx.run, ctx.plan
        if plan is not None:
            if plan.pending is None or plan.pending() is not getattr(ctx, "token", None):
                raise RuntimeError("the native encoder backward can run only once per forward")
            dx, layer_grads, after_grads = _replay_backward(plan, dout.contiguous())
            plan.pending = None
        else:
            if run.saved is None:
                raise RuntimeError("the native encoder backward can run only once per forward (activations are released)")
            dx, layer_grads, after_grads = stack_backward(dout, run)
            run.saved = None
        flat = [g for lg in layer_grads for g in lg]
        if after_grads is not None:
            flat += after_grads
        assert len(flat) == ctx.n_params
        return (dx, None, None) + tuple(flat)


def run_stack(x_emb, layers, after_norm, attn_mask, pos_embed, pad_mask, dtype, grad_sync=None, owner=None):
    """Differentiable (and dropout-capable) replacement of engine.run_layers for the batched forward.  ``owner``: the
    module that keeps the CUDA-graph plans (None: always eager)."""
    B, T, d = x_emb.shape
    for layer in layers:
        a = layer.self_attn
        if not hasattr(a, "pos_bias_u") and a.training and a.dropout.p > 0.0:
            raise NotImplementedError("native training of the absolute-position attention with attention dropout > 0 "
                                      "(its extra output dropout, attention.py:177) is not implemented")
    if pos_embed is not None and pos_embed.numel() != B * d:
        raise NotImplementedError("the training path implements the batched forward (one position row per batch element)")
    seed_host = torch.randint(0, 2 ** 62, (1,))                   # CPU generator: follows torch.manual_seed
    params = [p for layer in layers for p in layer_param_list(layer)]
    if after_norm is not None:
        params += [after_norm.weight, after_norm.bias]
    plan = None
    if (owner is not None and getattr(owner, "use_cuda_graphs", False) and x_emb.is_cuda and torch.is_grad_enabled()
            and not torch.cuda.is_current_stream_capturing()
            and (pad_mask is None or pad_mask.dim() != 3 or pad_mask.size(2) == 0 or tuple(pad_mask.shape) == (B, 1, T))
            and all(l.conv_module.norm.momentum is not None for l in layers)):
        plans = owner.__dict__.setdefault("_train_plans", {})
        key = _plan_key(x_emb, layers, after_norm, attn_mask, pad_mask, dtype, grad_sync, params)
        cand = plans.get(key)
        if cand is None:
            while len(plans) >= 4:
                plans.pop(next(iter(plans)))
            plans[key] = cand = TrainPlan()
        cand.calls += 1
        if cand.calls >= 2 and not cand.busy:                     # first call of a configuration runs eagerly (warm-up)
            if cand.fwd is None:
                _capture_forward(cand, x_emb, layers, after_norm, attn_mask, pad_mask, dtype, grad_sync)
            cand.run.grad_sync = grad_sync
            _fill_plan(cand, x_emb, engine._mask_u8(attn_mask), engine._mask_u8(pad_mask), seed_host)
            plan = cand
    if plan is not None:
        return EncoderStackFunction.apply(x_emb, None, plan, *params)
    run = TrainRun(layers, after_norm, attn_mask, pad_mask, B, T, dtype, seed_host.to(x_emb.device), grad_sync)
    return EncoderStackFunction.apply(x_emb, run, None, *params)
