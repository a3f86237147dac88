"""CPU oracle for the Conformer encoder hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the algorithm implemented by the
reference (Lingeng56/conformer-pytorch-lightning) in

    src/encoder.py, src/encoder_layer.py, src/attention.py,
    src/convolution.py, src/feedforward.py and the mask helpers of src/utils.py.

It is the *checker* for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it;
the product package (``conformer_pytorch_lightning_b200``) never does, and it
fails loudly when its CUDA extension is missing instead of falling back here.

Parity status: PINNED.  The reference is Python and has no tests or golden
vectors of its own, so the oracle is pinned against
outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference/src``) and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays
every one of them through this file.

Every function cites the reference ``file:line`` it restates.  A model is a
plain ``dict[str, np.ndarray]`` keyed exactly like the reference encoder's
``state_dict`` plus a small ``cfg`` dict of constructor
arguments.  All arithmetic is done in ``dtype`` (float32 by default, float64
for a high-precision yard-stick).
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def silu(x):
    """nn.SiLU (feedforward.py:9, convolution.py:24)."""
    return x * _sigmoid(x)


def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm(d, eps=1e-5) (encoder_layer.py:41-45, encoder.py:48)."""
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + eps) * w + b


def linear(x, w, b=None):
    """nn.Linear: y = x @ w.T + b."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def softmax_lastdim(x):
    m = np.max(x, axis=-1, keepdims=True)
    # rows that are entirely -inf give nan in torch.softmax; reproduce that
    with np.errstate(invalid="ignore"):
        e = np.exp(x - m)
        return e / e.sum(axis=-1, keepdims=True)


# --------------------------------------------------------------------------
# positional encoding tables
# --------------------------------------------------------------------------


def rel_pos_table(max_len, d_model, dtype=np.float32):
    """RelativePositionalEncoding.__init__ (attention.py:12-16): sin/cos table
    of shape (max_len, 1, d_model), no sqrt(d) scaling anywhere."""
    position = np.arange(max_len, dtype=np.float32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2, dtype=np.float32)
                      * np.float32(-math.log(10000.0) / d_model)).astype(np.float32)
    pe = np.zeros((max_len, 1, d_model), dtype=np.float32)
    pe[:, 0, 0::2] = np.sin(position * div_term)
    pe[:, 0, 1::2] = np.cos(position * div_term)
    return pe.astype(dtype)


def abs_pos_table(max_len, d_model, dtype=np.float32):
    """PositionalEncoding.__init__ (attention.py:110-115): same table but the
    reference stores it in float16 (then casts to the input dtype on use)."""
    pe = rel_pos_table(max_len, d_model, np.float32)
    return pe.astype(np.float16).astype(dtype)


# --------------------------------------------------------------------------
# masks (src/utils.py) -- integer/bool work, must be bit exact
# --------------------------------------------------------------------------


def make_pad_mask(lengths, max_len):
    """utils.py:84-93: True where the frame is PADDING."""
    lengths = np.asarray(lengths).astype(np.int64)
    return np.arange(max_len, dtype=np.int64)[None, :] >= lengths[:, None]


def subsequent_chunk_mask(size, chunk_size, num_left_chunks):
    """utils.py:96-111 as the literal per-row loop."""
    ret = np.zeros((size, size), dtype=bool)
    for i in range(size):
        if num_left_chunks < 0:
            start = 0
        else:
            start = max((i // chunk_size - num_left_chunks) * chunk_size, 0)
        ending = min((i // chunk_size + 1) * chunk_size, size)
        ret[i, start:ending] = True
    return ret


def make_attn_mask(max_len, pad_mask, use_dynamic_chunk, use_dynamic_left_chunk,
                   decoding_chunk_size, static_chunk_size, num_decoding_left_chunks,
                   draws=None):
    """utils.py:115-160.  ``pad_mask`` is (B,1,T) bool, True = valid.

    The reference draws the dynamic chunk size / left chunks with
    ``torch.randint`` from torch's global generator (utils.py:131,139).  numpy
    cannot replay that stream, so the raw draws are passed in as
    ``draws = (first_randint, second_randint_or_None)``; the arithmetic applied
    to them is restated here.
    """
    if use_dynamic_chunk:
        if decoding_chunk_size < 0:
            chunk_size, num_left = max_len, -1
        elif decoding_chunk_size > 0:
            chunk_size, num_left = decoding_chunk_size, num_decoding_left_chunks
        else:
            chunk_size = int(draws[0])
            num_left = -1
            if chunk_size > max_len // 2:
                chunk_size = max_len
            else:
                chunk_size = chunk_size % 25 + 1
                if use_dynamic_left_chunk:
                    num_left = int(draws[1])
        cm = subsequent_chunk_mask(max_len, chunk_size, num_left)[None]
        return pad_mask & cm
    if static_chunk_size > 0:
        cm = subsequent_chunk_mask(max_len, static_chunk_size, num_decoding_left_chunks)[None]
        return pad_mask & cm
    return pad_mask


# --------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------


def feed_forward(x, sd, prefix):
    """PositionwiseFeedForwardModule.forward (feedforward.py:16-21), eval mode
    (dropout = identity)."""
    h = silu(linear(x, sd[prefix + "w_1.weight"], sd[prefix + "w_1.bias"]))
    return linear(h, sd[prefix + "w_2.weight"], sd[prefix + "w_2.bias"])


def rel_mhsa(x, attn_mask, pos_embed, cache, sd, prefix, num_heads):
    """RelativeMultiHeadSelfAttentionModule.forward (attention.py:54-100).

    x (B,T,d); attn_mask (B,1,Tk)/(B,Tq,Tk) or None for the empty (0,0,0)
    sentinel; pos_embed (P,1,d) with P == B (batched forward, attention.py:20:
    the table is sliced by *batch size*) or P == Tk (streaming, B == 1);
    cache None or (B,H,C,2*d_k).  Returns (out (B,T,d), new_cache (B,H,Tk,2*d_k)).
    There is no rel_shift in the reference (attention.py:84-88).
    """
    B, T, d = x.shape
    H = num_heads
    dk = d // H
    q = linear(x, sd[prefix + "linear_q.weight"], sd[prefix + "linear_q.bias"]).reshape(B, T, H, dk)
    k = linear(x, sd[prefix + "linear_k.weight"], sd[prefix + "linear_k.bias"]).reshape(B, T, H, dk)
    v = linear(x, sd[prefix + "linear_v.weight"], sd[prefix + "linear_v.bias"]).reshape(B, T, H, dk)
    k = k.transpose(0, 2, 1, 3)                      # (B,H,T,dk)
    v = v.transpose(0, 2, 1, 3)
    if cache is not None and cache.shape[0] > 0:     # attention.py:70-74
        kc, vc = cache[..., :dk], cache[..., dk:]
        k = np.concatenate([kc, k], axis=2)
        v = np.concatenate([vc, v], axis=2)
    new_cache = np.concatenate([k, v], axis=-1)      # attention.py:76
    # attention.py:78-79: view(batch_size, -1, H, dk) of a (P,1,d) tensor
    p = linear(pos_embed, sd[prefix + "linear_pos.weight"]).reshape(B, -1, H, dk).transpose(0, 2, 1, 3)
    u = sd[prefix + "pos_bias_u"]
    vb = sd[prefix + "pos_bias_v"]
    q_u = (q + u).transpose(0, 2, 1, 3)              # (B,H,T,dk)
    q_v = (q + vb).transpose(0, 2, 1, 3)
    ac = q_u @ k.transpose(0, 1, 3, 2)               # (B,H,T,Tk)
    bd = q_v @ p.transpose(0, 1, 3, 2)               # (B,H,T,P') broadcast over keys when P'==1
    scores = (ac + bd) / np.asarray(math.sqrt(dk), dtype=x.dtype)
    if attn_mask is not None and attn_mask.shape[2] > 0:
        m = (attn_mask[:, None] == 0)                # attention.py:90
        scores = np.where(m, -np.inf, scores)
        attn = softmax_lastdim(scores)
        attn = np.where(m, 0.0, attn)                # also turns the NaN rows into 0
    else:
        attn = softmax_lastdim(scores)
    attn = attn.astype(x.dtype)
    o = (attn @ v).transpose(0, 2, 1, 3).reshape(B, T, d)
    out = linear(o, sd[prefix + "linear_out.weight"], sd[prefix + "linear_out.bias"])
    return out, new_cache


def abs_mhsa(x, attn_mask, cache, sd, prefix, num_heads):
    """MultiHeadSelfAttentionModule.forward (attention.py:148-179), eval."""
    B, T, d = x.shape
    H = num_heads
    dk = d // H
    q = linear(x, sd[prefix + "linear_q.weight"], sd[prefix + "linear_q.bias"]).reshape(B, T, H, dk).transpose(0, 2, 1, 3)
    k = linear(x, sd[prefix + "linear_k.weight"], sd[prefix + "linear_k.bias"]).reshape(B, T, H, dk).transpose(0, 2, 1, 3)
    v = linear(x, sd[prefix + "linear_v.weight"], sd[prefix + "linear_v.bias"]).reshape(B, T, H, dk).transpose(0, 2, 1, 3)
    if cache is not None and cache.shape[0] > 0:
        kc, vc = cache[..., :dk], cache[..., dk:]
        k = np.concatenate([kc, k], axis=2)
        v = np.concatenate([vc, v], axis=2)
    new_cache = np.concatenate([k, v], axis=-1)
    scores = (q @ k.transpose(0, 1, 3, 2)) / np.asarray(math.sqrt(dk), dtype=x.dtype)
    if attn_mask is not None and attn_mask.shape[2] > 0:
        m = (attn_mask[:, None] == 0)
        scores = np.where(m, -np.inf, scores)
        attn = np.where(m, 0.0, softmax_lastdim(scores))
    else:
        attn = softmax_lastdim(scores)
    attn = attn.astype(x.dtype)
    o = (attn @ v).transpose(0, 2, 1, 3).reshape(B, T, d)
    return linear(o, sd[prefix + "linear_out.weight"], sd[prefix + "linear_out.bias"]), new_cache


def depthwise_conv1d(x, w, b):
    """nn.Conv1d(d, d, k, padding=(k-1)//2, groups=d) on channel-last data.
    x (B,T,d); w (d,1,k); zero padded, non-causal (convolution.py:16-23)."""
    B, T, d = x.shape
    k = w.shape[-1]
    pad = (k - 1) // 2
    xp = np.zeros((B, T + 2 * pad, d), dtype=x.dtype)
    xp[:, pad:pad + T] = x
    y = np.zeros((B, T, d), dtype=x.dtype)
    for j in range(k):
        y += xp[:, j:j + T, :] * w[:, 0, j][None, None, :]
    return y + b


def conv_module(x, pad_mask, sd, prefix, training=False, bn_state=None):
    """ConvolutionModule.forward (convolution.py:34-49), on (B,T,d) channel-last
    (the reference transposes to (B,d,T) and back).  pad_mask (B,1,T) bool True =
    valid, or None for the empty sentinel.  Eval: BatchNorm running statistics.
    training=True: batch statistics over all B*T positions, *unmasked*
    (convolution.py:44), and running-stat update written into ``bn_state``."""
    d = x.shape[-1]
    if pad_mask is not None and pad_mask.shape[2] > 0:
        valid = pad_mask.transpose(0, 2, 1)                       # (B,T,1)
        x = np.where(valid, x, 0.0).astype(x.dtype)
    w1 = sd[prefix + "pointwise_conv1.weight"][:, :, 0]           # (2d,d)
    y = linear(x, w1, sd[prefix + "pointwise_conv1.bias"])
    y = y[..., :d] * _sigmoid(y[..., d:])                         # GLU(dim=channel), convolution.py:15,42
    y = depthwise_conv1d(y, sd[prefix + "depthwise_conv.weight"], sd[prefix + "depthwise_conv.bias"])
    g, beta = sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]
    if training:
        flat = y.reshape(-1, d)
        mean = flat.mean(axis=0)
        var = flat.var(axis=0)                                    # biased, used to normalise
        if bn_state is not None:
            n = flat.shape[0]
            mom = 0.1
            bn_state["running_mean"] = (1 - mom) * bn_state["running_mean"] + mom * mean
            bn_state["running_var"] = (1 - mom) * bn_state["running_var"] + mom * var * n / max(n - 1, 1)
            bn_state["num_batches_tracked"] = bn_state["num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + "norm.running_mean"], sd[prefix + "norm.running_var"]
    y = (y - mean) / np.sqrt(var + 1e-5) * g + beta
    y = silu(y.astype(x.dtype))
    w2 = sd[prefix + "pointwise_conv2.weight"][:, :, 0]
    y = linear(y, w2, sd[prefix + "pointwise_conv2.bias"])
    if pad_mask is not None and pad_mask.shape[2] > 0:
        y = np.where(valid, y, 0.0).astype(x.dtype)
    return y


def encoder_layer(x, attn_mask, pos_embed, pad_mask, attn_cache, sd, prefix, cfg):
    """ConformerEncoderLayer.forward (encoder_layer.py:49-71), eval mode."""
    half = np.asarray(0.5, dtype=x.dtype)
    y = layer_norm(x, sd[prefix + "norm_ff_macaron.weight"], sd[prefix + "norm_ff_macaron.bias"])
    x = x + half * feed_forward(y, sd, prefix + "feed_forward_macaron.")
    y = layer_norm(x, sd[prefix + "norm_mha.weight"], sd[prefix + "norm_mha.bias"])
    if cfg.get("use_relative", False):
        y, new_cache = rel_mhsa(y, attn_mask, pos_embed, attn_cache, sd, prefix + "self_attn.", cfg["num_heads"])
    else:
        y, new_cache = abs_mhsa(y, attn_mask, attn_cache, sd, prefix + "self_attn.", cfg["num_heads"])
    x = x + y
    y = layer_norm(x, sd[prefix + "norm_conv.weight"], sd[prefix + "norm_conv.bias"])
    x = x + conv_module(y, pad_mask, sd, prefix + "conv_module.")
    y = layer_norm(x, sd[prefix + "norm_ff.weight"], sd[prefix + "norm_ff.bias"])
    x = x + half * feed_forward(y, sd, prefix + "feed_forward.")
    x = layer_norm(x, sd[prefix + "norm_final.weight"], sd[prefix + "norm_final.bias"])
    return x.astype(y.dtype), new_cache


def _conv2d_s2_relu(x, w, b):
    """nn.Conv2d(cin, cout, 3, 2) + ReLU on (B,cin,H,W) via im2col."""
    B, cin, H, W = x.shape
    Ho, Wo = (H - 3) // 2 + 1, (W - 3) // 2 + 1
    cols = np.empty((B, Ho, Wo, cin, 3, 3), dtype=x.dtype)
    for i in range(3):
        for j in range(3):
            cols[:, :, :, :, i, j] = x[:, :, i:i + 2 * Ho - 1:2, j:j + 2 * Wo - 1:2].transpose(0, 2, 3, 1)
    y = cols.reshape(B * Ho * Wo, cin * 9) @ w.reshape(w.shape[0], -1).T + b
    return np.maximum(y, 0).reshape(B, Ho, Wo, -1).transpose(0, 3, 1, 2)


def subsampling(feats, pad_mask, sd, dtype):
    """ConvolutionSubSampling.forward (convolution.py:70-76) without the
    positional-encoding call.  feats (B,Tin,idim); pad_mask (B,1,Tin)."""
    outs = []
    for b in range(feats.shape[0]):                 # per utterance keeps im2col small
        x = feats[b:b + 1, None].astype(dtype)
        x = _conv2d_s2_relu(x, sd["embed.conv.0.weight"], sd["embed.conv.0.bias"])
        x = _conv2d_s2_relu(x, sd["embed.conv.2.weight"], sd["embed.conv.2.bias"])
        _, c, t, f = x.shape
        x = x.transpose(0, 2, 1, 3).reshape(1, t, c * f)
        outs.append(linear(x, sd["embed.out.0.weight"], sd["embed.out.0.bias"]))
    return np.concatenate(outs, axis=0), pad_mask[:, :, 2::2][:, :, 2::2]


def _pe_table(cfg, dtype):
    if cfg.get("use_relative", False):
        return rel_pos_table(cfg.get("max_len", 5000), cfg["encoder_dim"], dtype)
    return abs_pos_table(cfg.get("max_len", 5000), cfg["encoder_dim"], dtype)


def _cast_sd(sd, dtype):
    return {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in sd.items()}


def encoder_layers(x, attn_mask, pos_embed, pad_mask, sd, cfg):
    """The measured path: encoder.py:72-74 (layer loop + after_norm)."""
    for i in range(cfg["encoder_num_layers"]):
        x, _ = encoder_layer(x, attn_mask, pos_embed, pad_mask, None, sd, f"encoders.{i}.", cfg)
    return layer_norm(x, sd["after_norm.weight"], sd["after_norm.bias"])


def encoder_embed(feats, lengths, sd, cfg, dtype=np.float32, draws=None,
                  decoding_chunk_size=0, num_decoding_left_chunks=-1):
    """encoder.py:59-71: cmvn -> pad mask -> subsampling -> positions -> attention mask."""
    sd = _cast_sd(sd, dtype)
    feats = np.asarray(feats, dtype=dtype)
    if "global_cmvn.mean" in sd:                                   # cmvn.py:22-33
        feats = (feats - sd["global_cmvn.mean"]) * sd["global_cmvn.istd"]
    B, Tin, _ = feats.shape
    pad_mask = ~make_pad_mask(lengths, Tin)[:, None, :]
    x, pad_mask = subsampling(feats, pad_mask, sd, dtype)
    pe = _pe_table(cfg, dtype)
    pos_embed = pe[0:B]                                            # attention.py:20 -- sliced by batch size (D2)
    if not cfg.get("use_relative", False):
        x = x + pos_embed                                          # attention.py:119-120, (B,1,d) broadcast over T
    attn_mask = make_attn_mask(x.shape[1], pad_mask,
                               cfg.get("use_dynamic_chunk_size", False),
                               cfg.get("use_dynamic_left_chunk", False),
                               decoding_chunk_size, cfg.get("static_chunk_size", -1),
                               num_decoding_left_chunks, draws)
    return x, pos_embed, pad_mask, attn_mask


def encoder_forward(feats, lengths, sd, cfg, dtype=np.float32, draws=None,
                    decoding_chunk_size=0, num_decoding_left_chunks=-1):
    """ConformerEncoder.forward (encoder.py:54-75), eval.  Returns
    (outputs (B,T,d), pad_mask (B,1,T) bool, attn_mask)."""
    sdc = _cast_sd(sd, dtype)
    x, pos_embed, pad_mask, attn_mask = encoder_embed(
        feats, lengths, sd, cfg, dtype, draws, decoding_chunk_size, num_decoding_left_chunks)
    out = encoder_layers(x, attn_mask, pos_embed, pad_mask, sdc, cfg)
    return out, pad_mask, attn_mask


def encoder_forward_chunk(feats, offset, required_cache_size, attn_cache, sd, cfg, dtype=np.float32):
    """ConformerEncoder.forward_chunk (encoder.py:78-123), B == 1, no mask.
    attn_cache: (L,H,C,2dk) or an array whose size(0) == 0.  Returns
    (out (1,chunk,d), new_attn_cache (L,H,C',2dk))."""
    sd = _cast_sd(sd, dtype)
    feats = np.asarray(feats, dtype=dtype)
    if "global_cmvn.mean" in sd:
        feats = (feats - sd["global_cmvn.mean"]) * sd["global_cmvn.istd"]
    ones = np.ones((1, 1, feats.shape[1]), dtype=bool)
    x, _ = subsampling(feats, ones, sd, dtype)
    pe = _pe_table(cfg, dtype)
    if not cfg.get("use_relative", False):
        x = x + pe[offset:offset + 1]                              # attention.py:119-120 with size(0) == 1
    L = attn_cache.shape[0]
    cache_size = attn_cache.shape[2] if attn_cache.ndim == 4 else 0
    chunk = x.shape[1]
    key_size = cache_size + chunk
    pos_embed = pe[offset - cache_size: offset - cache_size + key_size]   # encoder.py:98-100
    if required_cache_size < 0:
        nxt = 0
    elif required_cache_size == 0:
        nxt = key_size
    else:
        nxt = max(key_size - required_cache_size, 0)
    caches = []
    for i in range(cfg["encoder_num_layers"]):
        c = attn_cache[i:i + 1] if L > 0 else None
        x, nc = encoder_layer(x, None, pos_embed, None, c, sd, f"encoders.{i}.", cfg)
        caches.append(nc[:, :, nxt:, :])
    out = layer_norm(x, sd["after_norm.weight"], sd["after_norm.bias"])
    return out, np.concatenate(caches, axis=0)


def encoder_forward_chunk_by_chunk(feats, decoding_chunk_size, num_decoding_left_chunks, sd, cfg,
                                   dtype=np.float32):
    """ConformerEncoder.forward_chunk_by_chunk (encoder.py:125-153)."""
    stride = 4 * decoding_chunk_size
    window = (decoding_chunk_size - 1) * 4 + 7
    n = feats.shape[1]
    cache = np.zeros((0, 0, 0, 0), dtype=dtype)
    outs, offset = [], 0
    req = decoding_chunk_size * num_decoding_left_chunks
    for cur in range(0, n - 7 + 1, stride):
        end = min(cur + window, n)
        o, cache = encoder_forward_chunk(feats[:, cur:end], offset, req, cache, sd, cfg, dtype)
        outs.append(o)
        offset += o.shape[1]
    return np.concatenate(outs, axis=1)


def ctc_greedy_ids(encoder_out, valid_lengths, ctc_w, ctc_b):
    """Harness-defined CTC greedy decode (the reference has only the
    ctc_lo projection, decoder.py:14,19): argmax over ctc_lo(encoder_out), collapse
    repeats, drop blank 0.  Returns (frame_argmax (B,T) int64, list of id lists)."""
    logits = linear(encoder_out, ctc_w, ctc_b)
    best = logits.argmax(axis=-1)
    hyps = []
    for b, n in enumerate(valid_lengths):
        seq, prev = [], -1
        for t in best[b, :int(n)]:
            if t != prev and t != 0:
                seq.append(int(t))
            prev = t
        hyps.append(seq)
    return best, hyps


# --------------------------------------------------------------------------
# deterministic, framework-independent weights (so the GPU box can rebuild the
# exact tensors the golden fixtures were generated with, without the reference)
# --------------------------------------------------------------------------


def make_state_dict(cfg, seed=0, dtype=np.float32):
    """Random-init encoder weights in the reference's state_dict layout and
    registration order.  Distributions follow PyTorch's defaults
    (uniform +-1/sqrt(fan_in) for Linear/Conv, xavier-uniform pos_bias) but the
    normalisation layers get non-trivial affine parameters and the BatchNorm gets
    non-trivial running statistics so that folding mistakes are visible."""
    rs = np.random.RandomState(seed)
    d, F, H = cfg["encoder_dim"], cfg["hidden_dim"], cfg["num_heads"]
    k, idim, L = cfg["kernel_size"], cfg["input_dim"], cfg["encoder_num_layers"]
    dk = d // H
    sd = {}

    def uni(shape, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return rs.uniform(-bound, bound, size=shape).astype(dtype)

    def norm(prefix):
        sd[prefix + "weight"] = (1.0 + 0.1 * rs.standard_normal(d)).astype(dtype)
        sd[prefix + "bias"] = (0.1 * rs.standard_normal(d)).astype(dtype)

    def ffn(prefix):
        sd[prefix + "w_1.weight"] = uni((F, d), d)
        sd[prefix + "w_1.bias"] = uni((F,), d)
        sd[prefix + "w_2.weight"] = uni((d, F), F)
        sd[prefix + "w_2.bias"] = uni((d,), F)

    sd["embed.conv.0.weight"] = uni((d, 1, 3, 3), 9)
    sd["embed.conv.0.bias"] = uni((d,), 9)
    sd["embed.conv.2.weight"] = uni((d, d, 3, 3), 9 * d)
    sd["embed.conv.2.bias"] = uni((d,), 9 * d)
    fdim = ((idim - 1) // 2 - 1) // 2
    sd["embed.out.0.weight"] = uni((d, d * fdim), d * fdim)
    sd["embed.out.0.bias"] = uni((d,), d * fdim)
    for i in range(L):
        p = f"encoders.{i}."
        ffn(p + "feed_forward.")
        if cfg.get("use_relative", False):
            xav = math.sqrt(6.0 / (H + dk))
            sd[p + "self_attn.linear_pos.weight"] = uni((d, d), d)
        for n in ("k", "q", "v", "out"):
            sd[p + f"self_attn.linear_{n}.weight"] = uni((d, d), d)
            sd[p + f"self_attn.linear_{n}.bias"] = uni((d,), d)
        if cfg.get("use_relative", False):
            sd[p + "self_attn.pos_bias_u"] = rs.uniform(-xav, xav, size=(H, dk)).astype(dtype)
            sd[p + "self_attn.pos_bias_v"] = rs.uniform(-xav, xav, size=(H, dk)).astype(dtype)
        c = p + "conv_module."
        sd[c + "pointwise_conv1.weight"] = uni((2 * d, d, 1), d)
        sd[c + "pointwise_conv1.bias"] = uni((2 * d,), d)
        sd[c + "depthwise_conv.weight"] = uni((d, 1, k), k)
        sd[c + "depthwise_conv.bias"] = uni((d,), k)
        sd[c + "norm.weight"] = (1.0 + 0.1 * rs.standard_normal(d)).astype(dtype)
        sd[c + "norm.bias"] = (0.1 * rs.standard_normal(d)).astype(dtype)
        sd[c + "norm.running_mean"] = (0.1 * rs.standard_normal(d)).astype(dtype)
        sd[c + "norm.running_var"] = rs.uniform(0.5, 1.5, size=d).astype(dtype)
        sd[c + "norm.num_batches_tracked"] = np.asarray(3, dtype=np.int64)
        sd[c + "pointwise_conv2.weight"] = uni((d, d, 1), d)
        sd[c + "pointwise_conv2.bias"] = uni((d,), d)
        ffn(p + "feed_forward_macaron.")
        for n in ("norm_ff", "norm_ff_macaron", "norm_mha", "norm_conv", "norm_final"):
            norm(p + n + ".")
    norm("after_norm.")
    return sd


def conformer_cfg(name="M", **over):
    """Named hyper-parameter sets: M = the reference's shipped config
    (train.sh:16-37, deploy_common.py:12-33); L = BASELINE.json config 3."""
    base = dict(input_dim=80, kernel_size=15, encoder_dim=256, dropout=0.1, attention_dropout=0.1,
                pos_enc_dropout=0.1, hidden_dim=2048, num_heads=4, encoder_num_layers=12,
                max_len=5000, use_relative=True, use_dynamic_chunk_size=False,
                use_dynamic_left_chunk=False, static_chunk_size=-1)
    if name == "L":
        base.update(kernel_size=31, encoder_dim=512, num_heads=8, encoder_num_layers=17)
    elif name == "tiny":
        base.update(encoder_dim=64, hidden_dim=128, num_heads=2, encoder_num_layers=2, kernel_size=7)
    base.update(over)
    return base
