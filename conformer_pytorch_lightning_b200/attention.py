"""Drop-in for the reference's src/attention.py: positional-encoding tables and the two
multi-head self-attention modules, executed by the native QKV GEMM + flash-attention kernels."""
import math

import torch
import torch.nn as nn

from . import engine


def _sincos_table(max_len, d_model):
    # attention.py:12-16 / 111-115 of the reference: interleaved sin/cos, no sqrt(d) scaling
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, 1, d_model)
    pe[:, 0, 0::2] = torch.sin(pos * div)
    pe[:, 0, 1::2] = torch.cos(pos * div)
    return pe


class _PositionTable(nn.Module):
    """The reference keeps ``pe`` as a plain attribute and re-casts it *in place* to the dtype of
    every input, so one bf16 call permanently degrades the table.  Here the master table is immutable
    (a non-persistent buffer: it follows ``.to(device)`` and stays out of the state_dict) and per-(device, dtype)
    casts are cached; there is no per-call mutable state (callable from any thread)."""

    def __init__(self, d_model, dropout, max_len, table_dtype):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", _sincos_table(max_len, d_model).to(table_dtype), persistent=False)
        self._cast = {}

    def _table(self, device, dtype):
        if device == self.pe.device and dtype == self.pe.dtype:
            return self.pe
        key = (device, dtype)
        t = self._cast.get(key)
        if t is None or t.data_ptr() == 0:
            t = self._cast[key] = self.pe.to(device).to(dtype)
        return t

    def _apply(self, fn, *args, **kwargs):
        self._cast = {}
        return super()._apply(fn, *args, **kwargs)

    def position_encoding(self, offset, size, apply_dropout=True, like=None):
        """``like``: tensor whose device / dtype the slice should have (default: the table's own fp32 / fp16)."""
        table = self.pe if like is None else self._table(like.device, like.dtype)
        pos_embed = table[offset: offset + size]
        if apply_dropout:
            pos_embed = self.dropout(pos_embed)
        return pos_embed


class RelativePositionalEncoding(_PositionTable):
    """attention.py:6-29.  NB: the table is sliced by ``inputs.size(0)`` = batch size."""

    def __init__(self, d_model, dropout, max_len=5000):
        super().__init__(d_model, dropout, max_len, torch.float32)

    def forward(self, inputs, offset=0):
        pos_embed = self.position_encoding(offset, inputs.size(0), False, like=inputs)
        return self.dropout(inputs), self.dropout(pos_embed)


class PositionalEncoding(_PositionTable):
    """attention.py:105-127: absolute variant; the reference builds this table in float16."""

    def __init__(self, d_model, dropout, max_len=5000):
        super().__init__(d_model, dropout, max_len, torch.float16)

    def forward(self, inputs, offset=0):
        pos_embed = self.position_encoding(offset, inputs.size(0), False, like=inputs)
        x = inputs + pos_embed
        return self.dropout(x), self.dropout(pos_embed)


def _check_head_dim(encoder_dim, num_heads):
    if encoder_dim % num_heads != 0 or encoder_dim // num_heads != 64:
        raise ValueError(f"the native attention kernels are specialised for a head dimension of 64 (Conformer-M: 256/4, "
                         f"Conformer-L: 512/8); got encoder_dim={encoder_dim}, num_heads={num_heads}")


class _MHSABase(nn.Module):
    def _run(self, query, key, value, inputs_attn_mask, pos_embed, cache, out_dropout):
        engine.check_inference_only(self, self.dropout.p, query)
        dtype = engine.resolve_dtype(self)
        B, T, d = query.shape
        self_attn = (key is query and value is query) or (key.shape == query.shape and value.shape == query.shape and
                                                          torch.equal(key, query) and torch.equal(value, query))
        if not self_attn:
            return self._cross(query, key, value, inputs_attn_mask, cache, dtype)
        y = query.reshape(B * T, d).to(dtype).contiguous()
        x = torch.zeros((B * T, d), dtype=torch.float32, device=y.device)
        new_cache = engine.mhsa_into(x, y, B, T, self.num_heads, self.derived_weights(dtype), inputs_attn_mask,
                                     pos_embed, cache, True, engine.thread_workspace())
        return x.view(B, T, d).to(query.dtype), new_cache.to(query.dtype)

    def _cross(self, query, key, value, mask, cache, dtype):
        """query, key and value are different tensors (attention.py:62-64 projects each with its own Linear; the
        reference's encoder never does this, its decoder_layer.py does): three projection GEMMs, then the attention
        kernel with Tq != Tk and the output projection.  For the relative variant the position term needs one position
        row per batch element (constant along keys, softmax-invariant) -- anything else raises."""
        import math as _m
        B, Tq, d = query.shape
        Tk = key.shape[1]
        if key.shape[0] != B or value.shape != key.shape or key.shape[2] != d:
            raise RuntimeError("attention: key / value must be (B, Tk, d) with the batch and model size of query")
        if cache is not None and cache.dim() == 4 and cache.size(0) > 0:
            raise NotImplementedError("cross-attention with a streaming cache is not implemented")
        W = self.derived_weights(dtype)
        H = self.num_heads
        dev = query.device
        ws = engine.thread_workspace()
        q = ws.get("xq", (B * Tq, d), dtype, dev)
        k = ws.get("xk", (B * Tk, d), dtype, dev)
        v = ws.get("xv", (B * Tk, d), dtype, dev)
        engine.ops.gemm(query.reshape(B * Tq, d).to(dtype).contiguous(), W["wqkv"][:d], W["bqkv"][:d].contiguous(), q, engine.N.EPI_BIAS)
        engine.ops.gemm(key.reshape(B * Tk, d).to(dtype).contiguous(), W["wqkv"][d:2 * d], W["bqkv"][d:2 * d].contiguous(), k, engine.N.EPI_BIAS)
        engine.ops.gemm(value.reshape(B * Tk, d).to(dtype).contiguous(), W["wqkv"][2 * d:], W["bqkv"][2 * d:].contiguous(), v, engine.N.EPI_BIAS)
        ctx = ws.get("attn_ctx", (B * Tq, d), dtype, dev)
        engine.ops.attention(q.view(B, Tq, H, 64), k.view(B, Tk, H, 64), v.view(B, Tk, H, 64), ctx.view(B, Tq, d),
                             mask=engine._mask_u8(mask), scale=1.0 / _m.sqrt(64.0))
        x = torch.zeros((B * Tq, d), dtype=torch.float32, device=dev)
        engine.ops.gemm(ctx, W["wo"], W["bo"], x, engine.N.EPI_RESIDUAL, residual=x, alpha=1.0)
        new_cache = torch.cat([k.view(B, Tk, H, 64).permute(0, 2, 1, 3), v.view(B, Tk, H, 64).permute(0, 2, 1, 3)], dim=-1).float()
        return x.view(B, Tq, d).to(query.dtype), new_cache.to(query.dtype)

    def derived_weights(self, dtype):
        return self._derived.get(self, dtype, lambda dt: engine.mhsa_weights(self, dt))


class RelativeMultiHeadSelfAttentionModule(_MHSABase):
    """attention.py:34-100.  matrix_ac + matrix_bd without rel_shift:
    (q+u).k_j + (q+v).p_j = (q+u).(k_j+p_j) + (v-u).p_j, i.e. flash attention over shifted keys plus a
    per-key bias; with one position row per batch element (batched forward, D2) the bd term is
    constant along keys and drops out of the softmax."""

    def __init__(self, encoder_dim, num_heads, dropout):
        super().__init__()
        _check_head_dim(encoder_dim, num_heads)
        self.d_k = encoder_dim // num_heads
        self.num_heads = num_heads
        self.linear_pos = nn.Linear(encoder_dim, encoder_dim, bias=False)
        self.linear_k = nn.Linear(encoder_dim, encoder_dim)
        self.linear_q = nn.Linear(encoder_dim, encoder_dim)
        self.linear_v = nn.Linear(encoder_dim, encoder_dim)
        self.linear_out = nn.Linear(encoder_dim, encoder_dim)
        self.pos_bias_u = nn.Parameter(torch.Tensor(self.num_heads, self.d_k))
        self.pos_bias_v = nn.Parameter(torch.Tensor(self.num_heads, self.d_k))
        self.dropout = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)
        self._derived = engine.Derived()

    def forward(self, query, key, value, inputs_attn_mask, pos_embed=None, cache=torch.zeros((0, 0, 0, 0))):
        return self._run(query, key, value, inputs_attn_mask, pos_embed, cache, False)


class MultiHeadSelfAttentionModule(_MHSABase):
    """attention.py:130-179 (use_relative=False)."""

    def __init__(self, encoder_dim, num_heads, dropout):
        super().__init__()
        _check_head_dim(encoder_dim, num_heads)
        self.d_k = encoder_dim // num_heads
        self.num_heads = num_heads
        self.linear_k = nn.Linear(encoder_dim, encoder_dim)
        self.linear_q = nn.Linear(encoder_dim, encoder_dim)
        self.linear_v = nn.Linear(encoder_dim, encoder_dim)
        self.linear_out = nn.Linear(encoder_dim, encoder_dim)
        self.dropout = nn.Dropout(dropout)
        self._derived = engine.Derived()

    def forward(self, query, key, value, inputs_attn_mask, pos_embed=None, cache=torch.zeros((0, 0, 0, 0))):
        return self._run(query, key, value, inputs_attn_mask, None, cache, True)
