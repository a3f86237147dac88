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


def conv_module_train(x, pad_mask, sd, p, bn_state=None):
    """convolution.py:34-49 in training mode: BatchNorm batch statistics over all B*T positions, unmasked
    (convolution.py:44); ``bn_state`` (dict running_mean / running_var / num_batches_tracked) is updated in place
    like nn.BatchNorm1d does (momentum 0.1, unbiased running variance)."""
    x = x.transpose(1, 2)
    if pad_mask is not None and pad_mask.size(2) > 0:
        x = x.masked_fill(~pad_mask, 0.0)
    y = F.glu(F.conv1d(x, sd[p + "pointwise_conv1.weight"], sd[p + "pointwise_conv1.bias"]), dim=1)
    w = sd[p + "depthwise_conv.weight"]
    y = F.conv1d(y, w, sd[p + "depthwise_conv.bias"], padding=(w.size(-1) - 1) // 2, groups=w.size(0))
    rm = bn_state["running_mean"] if bn_state is not None else None
    rv = bn_state["running_var"] if bn_state is not None else None
    y = F.batch_norm(y, rm, rv, sd[p + "norm.weight"], sd[p + "norm.bias"], True, 0.1, 1e-5)
    if bn_state is not None:
        bn_state["num_batches_tracked"] += 1
    y = F.conv1d(F.silu(y), sd[p + "pointwise_conv2.weight"], sd[p + "pointwise_conv2.bias"])
    if pad_mask is not None and pad_mask.size(2) > 0:
        y = y.masked_fill(~pad_mask, 0.0)
    return y.transpose(1, 2)


def _layers(x, attn_mask, pos_embed, pad_mask, sd, cfg, training=False, bn_states=None):
    H = cfg["num_heads"]
    for i in range(cfg["encoder_num_layers"]):
        p = f"encoders.{i}."
        x = x + 0.5 * feed_forward(_ln(x, sd, p + "norm_ff_macaron."), sd, p + "feed_forward_macaron.")
        x = x + rel_mhsa(_ln(x, sd, p + "norm_mha."), attn_mask, pos_embed, sd, p + "self_attn.", H)
        y = _ln(x, sd, p + "norm_conv.")
        if training:
            x = x + conv_module_train(y, pad_mask, sd, p + "conv_module.", None if bn_states is None else bn_states[i])
        else:
            x = x + conv_module(y, pad_mask, sd, p + "conv_module.")
        x = x + 0.5 * feed_forward(_ln(x, sd, p + "norm_ff."), sd, p + "feed_forward.")
        x = _ln(x, sd, p + "norm_final.")
    return _ln(x, sd, "after_norm.")


def encoder_layers(x, attn_mask, pos_embed, pad_mask, sd, cfg):
    """encoder.py:72-74, relative-position model, eval mode."""
    with torch.no_grad():
        return _layers(x, attn_mask, pos_embed, pad_mask, sd, cfg)


def encoder_layers_train(x, attn_mask, pos_embed, pad_mask, sd, cfg, bn_states=None):
    """The same loop in training mode with every dropout at 0 (encoder_layer.py:56-70 with p = 0 is the identity),
    differentiable: ``sd`` may hold leaf tensors with requires_grad, torch autograd then yields the gradients the
    reference's modules would produce (same ATen ops in the same order)."""
    return _layers(x, attn_mask, pos_embed, pad_mask, sd, cfg, training=True, bn_states=bn_states)


# ---------------------------------------------------------------------------------------------- front-end
def subsampling(feats, pad_mask, sd):
    """ConvolutionSubSampling.forward (convolution.py:70-76) without the positional-encoding call."""
    x = feats.unsqueeze(1)
    x = F.relu(F.conv2d(x, sd["embed.conv.0.weight"], sd["embed.conv.0.bias"], stride=2))
    x = F.relu(F.conv2d(x, sd["embed.conv.2.weight"], sd["embed.conv.2.bias"], stride=2))
    b, c, t, f = x.shape
    x = F.linear(x.transpose(1, 2).contiguous().view(b, t, c * f), sd["embed.out.0.weight"], sd["embed.out.0.bias"])
    return x, pad_mask[:, :, 2::2][:, :, 2::2]


def rel_pos_table(max_len, d):
    """attention.py:10-16 (no sqrt(d) scaling, fp32)."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(1)


def chunk_mask(T, chunk, left):
    """utils.py:96-111 closed form."""
    i = torch.arange(T).unsqueeze(1)
    j = torch.arange(T).unsqueeze(0)
    vis = j < (i // chunk + 1) * chunk
    if left >= 0:
        vis = vis & (j >= (i // chunk - left) * chunk)
    return vis


def encoder_embed(feats, lengths, sd, cfg, grad=False):
    """encoder.py:59-71 for the relative-position model without dynamic chunks: pad mask -> subsampling ->
    positions (table sliced by batch size, D2) -> attention mask (pad mask, AND static chunk mask if configured)."""
    B, Tin, _ = feats.shape
    pad = (torch.arange(Tin).unsqueeze(0) < lengths.unsqueeze(1).long()).unsqueeze(1)
    with torch.set_grad_enabled(grad):
        x, pad = subsampling(feats, pad, sd)
    pos = rel_pos_table(cfg.get("max_len", 5000), cfg["encoder_dim"])[:B]
    attn = pad
    c = cfg.get("static_chunk_size", -1)
    if c > 0:
        attn = pad & chunk_mask(x.size(1), c, -1).unsqueeze(0)
    return x, pos, pad, attn


def encoder_forward(feats, lengths, sd, cfg, batch_chunk=None):
    """ConformerEncoder.forward (encoder.py:54-75), eval, relative positions.  ``batch_chunk`` bounds the (B,H,T,T)
    temporaries by running the utterances in groups (eval mode couples nothing across utterances; the position row of
    utterance b stays pe[b], D2)."""
    x, pos, pad, attn = encoder_embed(feats, lengths, sd, cfg)
    B = x.size(0)
    step = batch_chunk or B
    outs = [encoder_layers(x[i:i + step], attn[i:i + step], pos[i:i + step], pad[i:i + step], sd, cfg)
            for i in range(0, B, step)]
    return torch.cat(outs, 0), pad, attn, x, pos


# ---------------------------------------------------------------------------------------------- CTC head
def ctc_loss(encoder_out, encoder_out_lens, padded_labels, label_lengths, w, b):
    """CTCDecoder.forward (decoder.py:18-23) WITHOUT its always-on dropout: ctc_lo -> log_softmax ->
    nn.CTCLoss(reduction='sum', blank=0) / padded_labels.size(1)."""
    logits = F.linear(encoder_out, w, b)
    probs = logits.transpose(0, 1).log_softmax(2)
    loss = F.ctc_loss(probs.float(), padded_labels, encoder_out_lens, label_lengths, blank=0, reduction="sum")
    return loss / padded_labels.size(1)


def train_step_grads(feats, lengths, labels, label_lengths, sd, ctc_w, ctc_b, cfg):
    """Training-mode forward (dropout 0) + CTC loss + backward on torch's CPU autograd: the gradients the reference's
    training_step (module.py:49-69 -> model.py:115-124 -> decoder.py:18-23) produces for the CTC branch.
    Returns (loss, out, {state_dict key: grad}, {"ctc_lo.weight"/"ctc_lo.bias": grad}, bn_states)."""
    leaves = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in sd.items()}
    w = ctc_w.clone().requires_grad_()
    b = ctc_b.clone().requires_grad_()
    x, pos, pad, attn = encoder_embed(feats, lengths, leaves, cfg, grad=True)
    st = [{"running_mean": leaves[f"encoders.{i}.conv_module.norm.running_mean"],
           "running_var": leaves[f"encoders.{i}.conv_module.norm.running_var"], "num_batches_tracked": 0}
          for i in range(cfg["encoder_num_layers"])]
    out = encoder_layers_train(x, attn, pos, pad, leaves, cfg, bn_states=st)
    out_lens = pad.squeeze(1).sum(1)
    loss = ctc_loss(out, out_lens, labels, label_lengths, w, b)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.is_floating_point() and v.requires_grad}
    return loss.detach(), out.detach(), grads, {"ctc_lo.weight": w.grad, "ctc_lo.bias": b.grad}, st
