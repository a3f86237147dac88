"""Kernel chains of the Conformer encoder layer on top of the C-ABI ops.

This is the host-side schedule of the hot path (reference: encoder_layer.py:49-71 and the
modules it calls).  All tensors handled here are 2-D ``(N = B*T, channels)`` row-major
device buffers; the residual stream ``x`` is fp32 and is updated *in place* by GEMM
epilogues, GEMM operands are in the compute dtype (fp32 or bf16).  Per layer the
schedule is:

    W1+SiLU | W2+0.5*res+LN_mha | QKV | flash-attn | Wo+res+LN_conv(+mask) | PW1+GLU |
    dwconv+BN+SiLU | PW2+mask+res+LN_ff | W1+SiLU | W2+0.5*res+LN_final+LN_ffm(next layer)

(11 launches; each LayerNorm rides in the epilogue of the residual GEMM that produces its input)

Derived weights (bf16 copies, fused QKV with pos_bias_u folded into the q bias, the
BatchNorm-folded depthwise filter) are caches keyed on parameter versions; the
``state_dict`` keeps the reference layout untouched.
"""
from __future__ import annotations

import math
import os
import threading

import torch

from . import _native as N
from . import ops

_LN_EPS = 1e-5
# native kernel launches executed through CUDA-graph replays (they bypass cfm_launch_count)
GRAPH_REPLAYED_LAUNCHES = [0]


# --------------------------------------------------------------------------- compute dtype
def resolve_dtype(module):
    """fp32 unless the module was switched with ``set_compute_dtype`` or bf16 autocast is on."""
    dt = getattr(module, "compute_dtype", None)
    if dt is None and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        dt = torch.bfloat16
    return dt or torch.float32


def wants_autograd(module, *inputs):
    """True when the caller expects a differentiable result: grad mode is on and an input or a parameter requires grad."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs):
        return True
    return any(p.requires_grad for p in module.parameters())


def check_inference_only(module, dropout_p, *inputs):
    """Guard of the STANDALONE sub-module forwards (feed-forward / attention / convolution module called on their own):
    they run the inference kernels, whose outputs carry no autograd graph.  Differentiable and dropout-capable execution
    is implemented at the granularity the reference trains at -- ConformerEncoderLayer / ConformerEncoder (training.py)
    -- so a call that expects gradients or dropout here raises instead of silently returning a detached tensor."""
    for t in inputs:
        if isinstance(t, torch.Tensor) and not t.is_cuda:
            raise RuntimeError("expected a CUDA tensor (the B200 Conformer kernels have no CPU path)")
    if module.training and dropout_p > 0.0:
        raise NotImplementedError("dropout > 0 in training mode is implemented by ConformerEncoderLayer / ConformerEncoder "
                                  "(conformer_pytorch_lightning_b200.training), not by the stand-alone sub-module forward")
    if wants_autograd(module, *inputs):
        raise NotImplementedError("this stand-alone sub-module forward is inference only (its output would be detached); "
                                  "differentiate through ConformerEncoderLayer / ConformerEncoder, or call it under "
                                  "torch.no_grad()")


# --------------------------------------------------------------------------- workspace
class Workspace:
    """Named scratch buffers, re-used across calls on the same thread (no hidden allocation
    inside the native ops; the buffers live in PyTorch's caching allocator).  One flat buffer per
    (name, dtype, device) that only ever grows (geometrically); a request returns a view of its head, so shapes that
    change from call to call (streaming key lengths) do not accumulate buffers.  A buffer that is outgrown stays
    referenced (captured CUDA graphs may hold its address): total memory is bounded by twice the largest request."""

    def __init__(self):
        self._t = {}
        self._retired = []

    def get(self, name, shape, dtype, device):
        key = (name, dtype, device)
        n = 1
        for v in shape:
            n *= int(v)
        t = self._t.get(key)
        if t is None or t.numel() < n:
            if t is not None:
                self._retired.append(t)
                n_alloc = max(n, 2 * t.numel())
            else:
                n_alloc = n
            t = torch.empty(max(n_alloc, 1), dtype=dtype, device=device)
            self._t[key] = t
        return t[:n].view(shape)


_tls = threading.local()
_L2_PREFETCH = os.environ.get("CFM_B200_L2_PREFETCH", "0") == "1"     # measured: 1.688 ms with, 1.598 ms without (off)


def _side_stream(dev):
    ss = getattr(_tls, "side", None)
    if ss is None:
        ss = _tls.side = {}
    key = torch.device(dev)
    if key not in ss:
        ss[key] = torch.cuda.Stream(dev)
    return ss[key]



def thread_workspace():
    ws = getattr(_tls, "ws", None)
    if ws is None:
        ws = _tls.ws = Workspace()
    return ws


# --------------------------------------------------------------------------- derived weights
class Derived:
    """Cache of tensors derived from a module's parameters, one slot per compute dtype (bf16 and fp32 copies coexist:
    a captured CUDA graph holds raw pointers into its slot, so switching the compute dtype back and forth must not free
    the other dtype's tensors).  A slot is rebuilt -- in place when shapes allow, so captured graphs stay valid --
    whenever a parameter version, storage or the training flag changes; ``generation`` counts the rebuilds that had to
    re-allocate (captured graphs that baked in the old addresses must be dropped, see ConformerEncoder)."""

    def __init__(self):
        self.slots = {}
        self.generation = 0
        self.force = False        # one-shot: rebuild even if the key matches (used while capturing a training graph, so
                                  # that the re-derivation itself becomes part of the graph and runs at every replay)

    def get(self, module, dtype, build):
        tensors = list(module.parameters(recurse=True)) + list(module.buffers(recurse=True))
        key = (module.training,) + tuple((p.data_ptr(), p._version) for p in tensors)
        slot = self.slots.get(dtype)
        force, self.force = self.force, False
        if slot is None or slot[0] != key or force:
            with torch.no_grad():
                new = build(dtype)
            old = slot[1] if slot is not None else None
            if old is not None and old.keys() == new.keys() and all(
                    old[k].shape == new[k].shape and old[k].dtype == new[k].dtype and old[k].device == new[k].device
                    for k in new):
                # Entries that ARE module state (fp32 parameters / buffers handed through without a copy) are never written
                # to: in train mode ``dw_b`` is the depthwise bias parameter itself, in eval mode it is the BatchNorm-folded
                # bias -- copying the latter into the former would overwrite the parameter.  Such entries are replaced
                # (and the generation bumped: graphs that baked in the old address are dropped); private buffers are
                # refreshed in place so that captured graphs stay valid.
                owned = {t.untyped_storage().data_ptr() for t in tensors}
                # persistent sources: module state and the optimizer's low-precision mirrors of it (see mirror_valid): a new
                # entry that is a view of one of those is adopted as it is (it tracks its source without any copy)
                persistent = owned | {t._cfm_mirror.untyped_storage().data_ptr() for t in tensors
                                      if getattr(t, "_cfm_mirror", None) is not None}
                replaced = False
                for k in new:
                    if old[k].data_ptr() == new[k].data_ptr():
                        continue
                    if (old[k].untyped_storage().data_ptr() in owned or old[k].untyped_storage().data_ptr() in slot[2]
                            or new[k].untyped_storage().data_ptr() in persistent):
                        old[k] = new[k]
                        replaced = True
                    else:
                        old[k].copy_(new[k])
                new = old
                if replaced:
                    self.generation += 1
            else:
                self.generation += 1
            # storages of the module state this slot was built from (a re-pointed parameter leaves its old storage behind:
            # entries still aliasing it must be replaced, not refreshed, at the next rebuild)
            self.slots[dtype] = (key, new, {t.untyped_storage().data_ptr() for t in tensors})
            return new
        return slot[1]


def _act(w, dtype):
    w = w.detach()
    return w.contiguous() if w.dtype == dtype else w.to(dtype).contiguous()


def _f32(w):
    return w.detach().float().contiguous()


# ---- low-precision mirrors of the parameters (optim.FlatAdam): the optimizer kernel writes the bf16 copy of every updated
# parameter into a flat mirror buffer laid out like its flat fp32 buffer; ``p._cfm_mirror`` is the view that belongs to p and
# ``p._cfm_mirror_version`` the parameter version it was written for.  While the two agree the mirror IS the compute-dtype
# copy: no cast kernel, and a captured graph that reads it sees every optimizer step.
def mirror_valid(p):
    return getattr(p, "_cfm_mirror", None) is not None and p._version == getattr(p, "_cfm_mirror_version", -1)


def _actp(p, dtype, shape=None):
    """Compute-dtype copy of parameter ``p`` (optionally reshaped): its mirror when that is current, else a cast."""
    if mirror_valid(p) and p._cfm_mirror.dtype == dtype:
        return p._cfm_mirror if shape is None else p._cfm_mirror.view(shape)
    return _act(p if shape is None else p.detach().reshape(shape), dtype)


def _actp_cat(ps, dtype):
    """cat(ps, 0) in the compute dtype; free when the parameters' mirrors are adjacent in memory (bucket layout)."""
    if all(mirror_valid(p) and p._cfm_mirror.dtype == dtype for p in ps):
        ms = [p._cfm_mirror for p in ps]
        es = ms[0].element_size()
        if all(m.is_contiguous() and m.shape[1:] == ms[0].shape[1:] for m in ms) and all(
                ms[i + 1].untyped_storage().data_ptr() == ms[0].untyped_storage().data_ptr() and
                ms[i + 1].data_ptr() == ms[i].data_ptr() + ms[i].numel() * es for i in range(len(ms) - 1)):
            rows = sum(m.shape[0] for m in ms)
            shape = (rows,) + tuple(ms[0].shape[1:])
            return torch.as_strided(ms[0], shape, ms[0].stride())
    return _act(torch.cat([p.detach() for p in ps], 0), dtype)


def ffn_weights(m, dtype):
    return {"w1": _actp(m.w_1.weight, dtype), "b1": _f32(m.w_1.bias),
            "w2": _actp(m.w_2.weight, dtype), "b2": _f32(m.w_2.bias)}


def mhsa_weights(m, dtype):
    rel = hasattr(m, "pos_bias_u")
    bq = m.linear_q.bias.detach().float()
    if rel:
        bq = bq + m.pos_bias_u.detach().float().reshape(-1)     # (q + u) folded into the q bias
    d = {"wqkv": _actp_cat([m.linear_q.weight, m.linear_k.weight, m.linear_v.weight], dtype),
         "bqkv": torch.cat([bq, m.linear_k.bias.detach().float(), m.linear_v.bias.detach().float()]).contiguous(),
         "wo": _actp(m.linear_out.weight, dtype), "bo": _f32(m.linear_out.bias)}
    if rel:
        d["wpos"] = _actp(m.linear_pos.weight, dtype)
        d["u"] = _f32(m.pos_bias_u)
        d["vb"] = _f32(m.pos_bias_v)
    return d


def conv_weights(m, dtype):
    dd = m.pointwise_conv2.weight.shape[0]
    k = m.depthwise_conv.weight.shape[-1]
    w1 = m.pointwise_conv1.weight.detach().reshape(2 * dd, dd)
    b1 = m.pointwise_conv1.bias
    b1 = _f32(b1) if b1 is not None else torch.zeros(2 * dd, device=w1.device)
    dw = m.depthwise_conv.weight.detach().float().reshape(dd, k)
    db = m.depthwise_conv.bias
    db = db.detach().float() if db is not None else torch.zeros(dd, device=w1.device)
    if not m.training:
        # eval: fold BatchNorm running statistics + conv bias (convolution.py:43-44)
        scale = m.norm.weight.detach().float() * torch.rsqrt(m.norm.running_var.float() + m.norm.eps)
        dw = dw * scale[:, None]
        db = (db - m.norm.running_mean.float()) * scale + m.norm.bias.detach().float()
    return {"w1": _actp(m.pointwise_conv1.weight, dtype, (2 * dd, dd)), "b1": b1, "dw_w": dw.t().contiguous(), "dw_b": db.contiguous(),
            "w2": _actp(m.pointwise_conv2.weight, dtype, (dd, dd)), "b2": _f32(m.pointwise_conv2.bias),
            "gamma": _f32(m.norm.weight), "beta": _f32(m.norm.bias)}


# --------------------------------------------------------------------------- module chains
def _residual_gemm(a, w, bias, x, alpha, row_valid, ln):
    """x += alpha * rowmask(a w^T + bias), optionally with the LayerNorm(s) that follow fused in.
    ln = None or dict(y=, g1=, b1=, g2=None, b2=None, y_row_valid=None)."""
    if ln is None:
        ops.gemm(a, w, bias, x, N.EPI_RESIDUAL, residual=x, alpha=alpha, row_valid=row_valid)
    else:
        ops.gemm_ln(a, w, bias, x, ln["y"], alpha=alpha, g1=ln["g1"], b1=ln["b1"], g2=ln.get("g2"), b2=ln.get("b2"),
                    row_valid=row_valid, y_row_valid=ln.get("y_row_valid"), eps=_LN_EPS)


def ffn_into(x, y, W, alpha, ws, ln=None):
    """x += alpha * (w_2 silu(w_1 y + b1) + b2)   (feedforward.py:16-21 + encoder_layer.py:58,69)."""
    n = y.shape[0]
    # one fused kernel on the tcgen05 engine (hidden activation stays on chip); `h` is only touched by the
    # unfused fallback inside the library (fp32 path, unsupported shapes)
    h = ws.get("ffn_h", (n, W["w1"].shape[0]), y.dtype, y.device)
    ops.ffn(y, W["w1"], W["b1"], W["w2"], W["b2"], x, alpha=alpha, ln=ln, hidden_ws=h)


def _mask_u8(mask):
    """Reference semantics are ``mask.eq(0)`` on any dtype (attention.py:90)."""
    if mask is None or mask.dim() != 3 or mask.size(2) == 0:
        return None
    if mask.dtype not in (torch.bool, torch.uint8):
        mask = mask != 0
    if mask.stride(2) != 1:
        mask = mask.contiguous()
    return mask


def mhsa_into(x, y, B, T, H, W, attn_mask, pos_embed, cache, want_cache, ws, ln=None, qkv_done=False):
    """x += linear_out(attn(...))  (attention.py:54-100 / 148-179 + encoder_layer.py:60-62).
    Returns new_cache (B,H,Tk,128) fp32 when want_cache else None.  qkv_done: the Q/K/V projections of ``y`` were already
    written to the workspace by the feed-forward call that produced ``y`` (see run_layers)."""
    n, d = y.shape
    dev, dt = y.device, y.dtype
    qkv = ws.get("qkv", (n, 3 * d), dt, dev)
    if not qkv_done:
        ops.gemm(y, W["wqkv"], W["bqkv"], qkv, N.EPI_BIAS)
    q5 = qkv.view(B, T, 3, H, 64)
    q, k, v = q5[:, :, 0], q5[:, :, 1], q5[:, :, 2]
    if cache is not None and cache.dim() == 4 and cache.size(0) > 0:
        # caller-held streaming cache (B,H,C,2*dk) (attention.py:70-74): PyTorch plumbing, tiny tensors
        kc = cache[..., :64].permute(0, 2, 1, 3).to(dt)
        vc = cache[..., 64:].permute(0, 2, 1, 3).to(dt)
        k = torch.cat([kc, k], dim=1)
        v = torch.cat([vc, v], dim=1)
    Tk = k.shape[1]
    new_cache = None
    if want_cache:
        new_cache = torch.cat([k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)], dim=-1).float()
    key_bias = None
    k_use = k
    if "wpos" in W and pos_embed is not None:
        rows = pos_embed.numel() // d
        if rows % B != 0:
            raise RuntimeError(f"pos_embed with {rows} rows cannot be viewed as (B={B}, -1, H, d_k)")
        P = rows // B
        if P == Tk and Tk > 1:
            # streaming position term: fold into keys + per-key bias
            pe = pos_embed.reshape(rows, d).to(dt).contiguous()
            p = ws.get("pos_p", (rows, d), dt, dev)
            ops.gemm(pe, W["wpos"], None, p, N.EPI_BIAS)
            k_use = ws.get("k_fold", (B, Tk, H, 64), dt, dev)
            key_bias = ws.get("key_bias", (B, H, Tk), torch.float32, dev)
            ops.relpos_keys(k, p.view(B, P, d), W["u"], W["vb"], k_use, key_bias)
        elif P != 1:
            raise RuntimeError(f"pos_embed gives {P} position rows per batch element; expected 1 or Tk={Tk}")
        # P == 1 (batched forward): matrix_bd is constant along keys -> softmax-invariant
    ctx = ws.get("attn_ctx", (n, d), dt, dev)
    if x.is_contiguous() and (ln is None or ln.get("g2") is None):
        # one library call: a single fused kernel per 128 query rows on the tcgen05 engine (bf16, 4 heads, T <= 256),
        # attention kernel -> ctx -> residual GEMM(+LN) otherwise
        ops.mhsa_out(q, k_use, v, W["wo"], W["bo"], x, mask=_mask_u8(attn_mask), key_bias=key_bias,
                     scale=1.0 / math.sqrt(64.0), ln=ln, ctx_ws=ctx)
        return new_cache
    ops.attention(q, k_use, v, ctx.view(B, T, d), mask=_mask_u8(attn_mask), key_bias=key_bias,
                  scale=1.0 / math.sqrt(64.0))
    _residual_gemm(ctx, W["wo"], W["bo"], x, 1.0, None, ln)
    return new_cache


def conv_into(x, y, B, T, W, row_valid, module, ws, ln=None):
    """x += mask(pw2(silu(bn(dw(glu(pw1(y)))))))  (convolution.py:34-49 + encoder_layer.py:64-66).
    ``y`` must already be zeroed on padded rows (done by the LayerNorm kernel's row mask)."""
    n, d = y.shape
    dev, dt = y.device, y.dtype
    g = ws.get("conv_glu", (n, d), dt, dev)
    c = ws.get("conv_dw", (n, d), dt, dev)
    if not module.training and y.is_contiguous() and x.is_contiguous() and (ln is None or ln.get("g2") is None):
        # inference: one library call -- a single fused kernel on the tcgen05 engine (bf16, d=256, k=15), the
        # GLU-GEMM / depthwise / GEMM(+LN) chain through g, c otherwise
        ops.conv_module(y, W["w1"], W["b1"], W["dw_w"], W["dw_b"], W["w2"], W["b2"], x, B, T, row_valid=row_valid, ln=ln,
                        glu_ws=g, dw_ws=c)
        return
    ops.gemm(y, W["w1"], W["b1"], g, N.EPI_BIAS_GLU)
    if not module.training:
        ops.dwconv(g.view(B, T, d), W["dw_w"], W["dw_b"], c.view(B, T, d), apply_silu=True)
    else:
        # BatchNorm batch statistics over all B*T rows, unmasked (convolution.py:44) + running update
        raw = ws.get("conv_raw", (n, d), torch.float32, dev)
        ops.dwconv(g.view(B, T, d), W["dw_w"], W["dw_b"], raw.view(B, T, d), apply_silu=False)
        st = ws.get("bn_st", (2, d), torch.float32, dev)
        st.zero_()
        ops.bn_stats(raw, st[0], st[1])
        mean = st[0] / n
        var = (st[1] / n - mean * mean).clamp_min_(0.0)
        bn = module.norm
        with torch.no_grad():
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var * (n / max(n - 1, 1)), alpha=mom)
            bn.num_batches_tracked += 1
        ops.bn_apply_silu(raw, mean.contiguous(), torch.rsqrt(var + bn.eps).contiguous(), W["gamma"], W["beta"], c)
    _residual_gemm(c, W["w2"], W["b2"], x, 1.0, row_valid, ln)


def _row_valid(pad_mask, B, T):
    if pad_mask is None or pad_mask.dim() != 3 or pad_mask.size(2) == 0:
        return None
    m = pad_mask
    if m.dtype not in (torch.bool, torch.uint8):
        m = m != 0
    m = m.expand(B, 1, T).contiguous()
    return m.view(B * T)


def run_layers(inputs, layers, after_norm, attn_mask, pos_embed, pad_mask, attn_caches, want_cache, dtype, inplace=False,
               ws=None, out_buf=None):
    """Layer loop of encoder.py:72-74 / 109-118.

    inputs (B,T,d) fp32 (not modified unless ``inplace``: the CUDA-graph plans run on their own static input buffer, which
    then doubles as the residual stream); layers: list of ConformerEncoderLayer; after_norm: LayerNorm
    module or None; attn_caches: list (one per layer) of (B,H,C,128) tensors or None.
    Returns (out (B,T,d) fp32 fresh tensor, [new caches])."""
    B, T, d = inputs.shape
    n = B * T
    dev = inputs.device
    ws = ws if ws is not None else thread_workspace()      # (a private workspace per concurrent sub-batch, see encoder)
    if inplace and inputs.dtype == torch.float32 and inputs.is_contiguous():
        x = inputs.view(n, d)
    else:
        x = torch.empty((n, d), dtype=torch.float32, device=dev)
        x.copy_(inputs.reshape(n, d))
    y = ws.get("ln_y", (n, d), dtype, dev)
    y2 = ws.get("ln_y2", (n, d), dtype, dev)
    row_valid = _row_valid(pad_mask, B, T)
    attn_mask = _mask_u8(attn_mask)          # normalised once, not per layer
    new_caches = []
    out = x
    ffm_done = False                         # the first feed-forward of layer i was already applied by layer i-1's chain
    # next-layer weight prefetch into L2 from a side stream (a parallel branch when captured into a CUDA graph): the
    # prefetch CTAs run on the SMs the single-wave kernels of the current layer leave idle
    prefetch = _L2_PREFETCH and dtype == torch.bfloat16 and n >= 4096 and len(layers) > 1
    side = _side_stream(dev) if prefetch else None
    cur = torch.cuda.current_stream(dev) if prefetch else None
    for i, layer in enumerate(layers):
        Wl = layer.derived_weights(dtype)
        H = layer.self_attn.num_heads
        if prefetch and i + 1 < len(layers):
            Wn_ = layers[i + 1].derived_weights(dtype)
            side.wait_stream(cur)
            ops.l2_prefetch([Wn_["ffm"]["w1"], Wn_["ffm"]["w2"], Wn_["mha"]["wqkv"], Wn_["mha"]["wo"], Wn_["conv"]["w1"],
                             Wn_["conv"]["w2"], Wn_["ff"]["w1"], Wn_["ff"]["w2"]], side)
        if i == 0:
            ops.layernorm(x, Wl["ffm_g"], Wl["ffm_b"], y=y)
        # every residual GEMM carries the LayerNorm that feeds the next module in its epilogue
        if not ffm_done:
            # first feed-forward + attention LayerNorm + Q/K/V projections in one library call
            b = dict(Wl["ffm"], alpha=0.5, g1=Wl["mha_g"], be1=Wl["mha_b"])
            h = ws.get("ffn_h", (n, b["w1"].shape[0]), y.dtype, dev)
            qkv = ws.get("qkv", (n, 3 * d), y.dtype, dev)
            ops.ffn_chain(y, None, b, x, y, proj=(Wl["mha"]["wqkv"], Wl["mha"]["bqkv"], qkv), hidden_ws=h)
        cache = attn_caches[i] if attn_caches is not None else None
        new_caches.append(mhsa_into(x, y, B, T, H, Wl["mha"], attn_mask, pos_embed, cache, want_cache, ws,
                                    ln={"y": y, "g1": Wl["conv_g"], "b1": Wl["conv_b"], "y_row_valid": row_valid},
                                    qkv_done=True))
        # the fused convolution module reads a halo of neighbouring rows of y, so its LayerNorm output goes to y2
        conv_into(x, y, B, T, Wl["conv"], row_valid, layer.conv_module, ws,
                  ln={"y": y2, "g1": Wl["ff_g"], "b1": Wl["ff_b"]})
        if i + 1 < len(layers):
            Wn = layers[i + 1].derived_weights(dtype)
            # this layer's second feed-forward (+ norm_final chained with norm_ff_macaron of the next layer) and the next
            # layer's first feed-forward (+ its attention LayerNorm) act on the same rows: one library call, one kernel
            # on the tcgen05 engine (the residual stream and the LayerNorm output in between stay on chip)
            ffm_done = Wl["ff"]["w1"].shape == Wn["ffm"]["w1"].shape
            if ffm_done:
                a = dict(Wl["ff"], alpha=0.5, g1=Wl["fin_g"], be1=Wl["fin_b"], g2=Wn["ffm_g"], be2=Wn["ffm_b"])
                b = dict(Wn["ffm"], alpha=0.5, g1=Wn["mha_g"], be1=Wn["mha_b"])
                h = ws.get("ffn_h", (n, a["w1"].shape[0]), y.dtype, dev)
                qkv = ws.get("qkv", (n, 3 * d), y.dtype, dev)
                ops.ffn_chain(y2, a, b, x, y, proj=(Wn["mha"]["wqkv"], Wn["mha"]["bqkv"], qkv), hidden_ws=h)
            else:
                ffn_into(x, y2, Wl["ff"], 0.5, ws, ln={"y": y, "g1": Wl["fin_g"], "b1": Wl["fin_b"],
                                                       "g2": Wn["ffm_g"], "b2": Wn["ffm_b"]})
        else:
            ffn_into(x, y2, Wl["ff"], 0.5, ws)
            if after_norm is not None:
                out = out_buf.view(n, d) if out_buf is not None else torch.empty((n, d), dtype=torch.float32, device=dev)
                ops.layernorm(x, Wl["fin_g"], Wl["fin_b"], g2=_f32(after_norm.weight), b2=_f32(after_norm.bias), y=out)
            else:
                ops.layernorm(x, Wl["fin_g"], Wl["fin_b"], x_out=x)
    if not layers and after_norm is not None:
        out = torch.empty((n, d), dtype=torch.float32, device=dev)
        ops.layernorm(x, _f32(after_norm.weight), _f32(after_norm.bias), y=out)
    if prefetch:
        cur.wait_stream(side)                # join the side branch (required inside graph capture)
    return out.view(B, T, d), new_caches
