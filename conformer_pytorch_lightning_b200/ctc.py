"""CTC head (scope row f2): greedy decoding and the training loss.

The reference's ``CTCDecoder`` (decoder.py:7-23) owns ``ctc_lo = nn.Linear(encoder_dim, vocab_size)`` and a training
loss; greedy decoding of its logits (argmax per frame, collapse repeats, drop blank 0) is what the parity gate "CTC
greedy ids" checks.  ``CTCGreedyHead`` keeps the same parameter (state_dict keys ``ctc_lo.weight`` / ``ctc_lo.bias``) and
runs the projection + argmax in one native call: on the bf16 / tcgen05 engine the (frames x vocab) logit matrix is
never materialised.
"""
import torch
import torch.nn as nn

from . import _native as N
from . import engine, ops
from . import train_ops as TO

__all__ = ["CTCGreedyHead", "CTCDecoder"]


class CTCGreedyHead(nn.Module):
    def __init__(self, encoder_dim, vocab_size, blank=0):
        super().__init__()
        self.ctc_lo = nn.Linear(encoder_dim, vocab_size)
        self.blank = blank
        self.compute_dtype = None

    def frame_ids(self, hs):
        """hs (B,T,d) encoder output on a CUDA device -> (B,T) int64 argmax of ctc_lo(hs)."""
        if not hs.is_cuda:
            raise RuntimeError("CTCGreedyHead: expected a CUDA tensor (the B200 kernels have no CPU path)")
        B, T, d = hs.shape
        dt = engine.resolve_dtype(self)
        x = hs.reshape(B * T, d).to(dt).contiguous()
        w = self.ctc_lo.weight.detach().to(dt).contiguous()
        b = self.ctc_lo.bias.detach().float().contiguous() if self.ctc_lo.bias is not None else None
        return ops.ctc_argmax(x, w, b).view(B, T).long()

    @torch.no_grad()
    def greedy(self, hs, lengths):
        """-> (frame ids (B,T) int64, list of token-id lists): collapse repeats, drop blanks, honour valid lengths."""
        ids = self.frame_ids(hs)
        T = ids.size(1)
        prev = torch.cat([torch.full_like(ids[:, :1], -1), ids[:, :-1]], dim=1)
        lengths = torch.as_tensor(lengths, device=ids.device)
        keep = (ids != self.blank) & (ids != prev) & (torch.arange(T, device=ids.device)[None, :] < lengths[:, None])
        ids_c, keep_c = ids.cpu(), keep.cpu()
        return ids, [ids_c[b][keep_c[b]].tolist() for b in range(ids.size(0))]


class _CTCLossFunction(torch.autograd.Function):
    """loss = sum_b nll_b / Lmax of log_softmax(x W^T + b) (decoder.py:19-22): native projection (tcgen05 GEMM on the bf16
    path, vocabulary padded to a multiple of 128 with zero rows), native log-softmax statistics + alpha / beta
    recursions; backward = native gradient w.r.t. the logits (written over the logits) + dgrad / wgrad GEMMs."""

    @staticmethod
    def forward(ctx, hs, weight, bias, labels, in_len, lab_len, dtype):
        B, T, d = hs.shape
        V = weight.shape[0]
        Vp = (V + 127) // 128 * 128
        dev = hs.device
        x = hs.reshape(B * T, d).to(dtype).contiguous()
        wp = torch.zeros((Vp, d), dtype=dtype, device=dev)
        wp[:V] = weight.detach().to(dtype)
        bp = torch.zeros(Vp, dtype=torch.float32, device=dev)
        if bias is not None:
            bp[:V] = bias.detach().float()
        logits = torch.empty((B * T, Vp), dtype=dtype, device=dev)
        ops.gemm(x, wp, bp, logits, N.EPI_BIAS)
        lab32 = labels.to(device=dev, dtype=torch.int32).contiguous()
        il = in_len.to(device=dev, dtype=torch.int32).contiguous()
        ll = lab_len.to(device=dev, dtype=torch.int32).contiguous()
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        ws = TO.ctc_loss_ws(B, T, lab32.shape[1], dev)
        TO.ctc_loss_fwd(logits, B, T, V, lab32, il, ll, nll, ws)
        ctx.saved = (x, wp, logits, lab32, il, ll, nll, ws)
        ctx.dims = (B, T, d, V, Vp, bias is not None, hs.dtype)
        return nll.sum() / max(lab32.shape[1], 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss):
        x, wp, logits, lab32, il, ll, nll, ws = ctx.saved
        ctx.saved = None
        B, T, d, V, Vp, has_bias, in_dtype = ctx.dims
        dev = x.device
        # d loss / d nll_b = dloss / Lmax.  1 / Lmax is folded into the logit gradient; dloss stays on the device (reading it
        # on the host would stall the host until the whole forward has finished, and with it every launch of the backward)
        TO.ctc_loss_bwd(logits, B, T, V, lab32, il, ll, nll, ws, 1.0 / max(lab32.shape[1], 1), logits)   # in place: logits -> dlogits
        dl = logits
        dl.mul_(dloss.to(device=dev, dtype=dl.dtype))
        dx = torch.empty((B * T, d), dtype=x.dtype, device=dev)
        ops.gemm_ex(dl, wp.t(), dx)                                                      # dgrad
        gw = torch.zeros((Vp, d), dtype=torch.float32, device=dev)
        ops.gemm_ex(dl.t(), x.t(), gw, accumulate=True)                                  # wgrad
        gb = None
        if has_bias:
            gbp = torch.zeros(Vp, dtype=torch.float32, device=dev)
            TO.colsum(dl, gbp)
            gb = gbp[:V]
        return dx.view(B, T, d).to(in_dtype), gw[:V], gb, None, None, None, None


class CTCDecoder(nn.Module):
    """Drop-in for the reference's CTCDecoder (decoder.py:7-23): same constructor, parameter (``ctc_lo``) and
    ``forward(encoder_out, encoder_out_lens, padded_labels, label_lengths) -> loss`` = CTCLoss(reduction='sum') of
    log_softmax(ctc_lo(dropout(encoder_out))) divided by padded_labels.size(1).  Like the reference the input dropout is
    applied in eval mode too (F.dropout's default training=True); set ``dropout=0`` for a deterministic loss."""

    def __init__(self, vocab_size, encoder_dim, dropout):
        super().__init__()
        self.ctc_lo = nn.Linear(encoder_dim, vocab_size)
        self.dropout = dropout
        self.compute_dtype = None

    def forward(self, encoder_out, encoder_out_lens, padded_labels, label_lengths):
        if not encoder_out.is_cuda:
            raise RuntimeError("CTCDecoder: expected a CUDA tensor (the B200 kernels have no CPU path)")
        dt = engine.resolve_dtype(self)
        x = nn.functional.dropout(encoder_out, self.dropout) if self.dropout > 0 else encoder_out
        return _CTCLossFunction.apply(x, self.ctc_lo.weight, self.ctc_lo.bias, padded_labels, encoder_out_lens,
                                      label_lengths, dt)
