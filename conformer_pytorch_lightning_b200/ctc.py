"""CTC head, greedy part (scope row f2).

The reference's ``CTCDecoder`` (decoder.py:7-23) owns ``ctc_lo = nn.Linear(encoder_dim, vocab_size)`` and a training
loss; greedy decoding of its logits (argmax per frame, collapse repeats, drop blank 0) is what the parity gate "CTC
greedy ids" checks.  ``CTCGreedyHead`` keeps the same parameter (state_dict keys ``ctc_lo.weight`` / ``ctc_lo.bias``) and
runs the projection + argmax in one native call: on the bf16 / tcgen05 engine the (frames x vocab) logit matrix is
never materialised.
"""
import torch
import torch.nn as nn

from . import engine, ops

__all__ = ["CTCGreedyHead"]


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
