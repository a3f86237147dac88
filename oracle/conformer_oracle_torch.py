"""Multi-threaded CPU port of the measured path -- TEST / BASELINE INFRASTRUCTURE ONLY.

Same algorithm as ``conformer_oracle.py`` (which is the numpy restatement pinned against the
reference's golden outputs) but expressed with torch's CPU kernels (MKL GEMM, oneDNN conv,
vectorised elementwise), i.e. the *same ATen arithmetic the reference's nn.Modules execute on a
CPU*, with all host threads.  It exists so that ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` arm is a fair stand-in for "the reference on the host cores": the reference itself is
Python source that cannot travel to the GPU box.  It is pinned by tests/test_oracle_golden.py against
the same golden vectors.  The product package never imports it.

Functions follow reference file:line exactly like the numpy oracle:
encoder_layer.py:49-71, attention.py:54-100, convolution.py:34-49, feedforward.py:16-21,
encoder.py:72-74.
"""
import math

import torch
import torch.nn.functional as F


def to_torch_sd(sd):
    import numpy as np
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def feed_forward(x, sd, p):
    return F.linear(F.silu(F.linear(x, sd[p + "w_1.weight"], sd[p + "w_1.bias"])), sd[p + "w_2.weight"], sd[p + "w_2.bias"])


def rel_mhsa(x, attn_mask, pos_embed, sd, p, H):
    B, T, d = x.shape
    dk = d // H
    q = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, T, H, dk)
    k = F.linear(x, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(B, T, H, dk).transpose(1, 2)
    v = F.linear(x, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(B, T, H, dk).transpose(1, 2)
    pp = F.linear(pos_embed, sd[p + "linear_pos.weight"]).view(B, -1, H, dk).transpose(1, 2)
    ac = torch.matmul((q + sd[p + "pos_bias_u"]).transpose(1, 2), k.transpose(-2, -1))
    bd = torch.matmul((q + sd[p + "pos_bias_v"]).transpose(1, 2), pp.transpose(-2, -1))
    scores = (ac + bd) / math.sqrt(dk)
    if attn_mask is not None and attn_mask.size(2) > 0:
        m = attn_mask.unsqueeze(1).eq(0)
        attn = torch.softmax(scores.masked_fill(m, -float("inf")), dim=-1).masked_fill(m, 0.0)
    else:
        attn = torch.softmax(scores, dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, d)
    return F.linear(o, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def conv_module(x, pad_mask, sd, p):
    x = x.transpose(1, 2)
    if pad_mask is not None and pad_mask.size(2) > 0:
        x = x.masked_fill(~pad_mask, 0.0)
    y = F.glu(F.conv1d(x, sd[p + "pointwise_conv1.weight"], sd[p + "pointwise_conv1.bias"]), dim=1)
    w = sd[p + "depthwise_conv.weight"]
    y = F.conv1d(y, w, sd[p + "depthwise_conv.bias"], padding=(w.size(-1) - 1) // 2, groups=w.size(0))
    y = F.batch_norm(y, sd[p + "norm.running_mean"], sd[p + "norm.running_var"], sd[p + "norm.weight"],
                     sd[p + "norm.bias"], False, 0.1, 1e-5)
    y = F.conv1d(F.silu(y), sd[p + "pointwise_conv2.weight"], sd[p + "pointwise_conv2.bias"])
    if pad_mask is not None and pad_mask.size(2) > 0:
        y = y.masked_fill(~pad_mask, 0.0)
    return y.transpose(1, 2)


def _ln(x, sd, p):
    return F.layer_norm(x, (x.size(-1),), sd[p + "weight"], sd[p + "bias"], 1e-5)


def encoder_layers(x, attn_mask, pos_embed, pad_mask, sd, cfg):
    """encoder.py:72-74, relative-position model, eval mode."""
    H = cfg["num_heads"]
    with torch.no_grad():
        for i in range(cfg["encoder_num_layers"]):
            p = f"encoders.{i}."
            x = x + 0.5 * feed_forward(_ln(x, sd, p + "norm_ff_macaron."), sd, p + "feed_forward_macaron.")
            x = x + rel_mhsa(_ln(x, sd, p + "norm_mha."), attn_mask, pos_embed, sd, p + "self_attn.", H)
            x = x + conv_module(_ln(x, sd, p + "norm_conv."), pad_mask, sd, p + "conv_module.")
            x = x + 0.5 * feed_forward(_ln(x, sd, p + "norm_ff."), sd, p + "feed_forward.")
            x = _ln(x, sd, p + "norm_final.")
        return _ln(x, sd, "after_norm.")
