"""Throughput driver around ``ConformerEncoder.forward`` for offline / batched inference.

The reference runs ``encoder(feats, lengths)`` batch after batch from a DataLoader (model.py:35-50 -> encoder.py:50-75);
on a B200 the layer stack of one 64 x 10 s batch takes about as long as moving its features in and its encoder states
out over PCIe, so this driver keeps three CUDA streams busy at once: while batch i is being encoded, batch i+1 is
copied host -> device and the result of batch i-1 device -> host.  Every batch goes through the unmodified public
``forward``; nothing is cached between batches.
"""
import torch

__all__ = ["EncoderPipeline"]


class _Slot:
    __slots__ = ("dev_in", "in_ready", "in_free")

    def __init__(self):
        self.dev_in = None
        self.in_ready = torch.cuda.Event()
        self.in_free = None


class EncoderPipeline:
    """``stream(batches)`` yields ``(out_host, pad_mask)`` per batch, in order, ``depth`` batches behind the one being
    submitted; ``run(batches)`` collects them into a list.

    batches: iterable of ``(feats, lengths)``; ``feats`` (B, T, idim) fp32 in pinned host memory (pageable memory
    works but serialises the copy), ``lengths`` (B,) int tensor on the host or on the device.
    ``out_bufs``: optional list of >= ``depth`` pinned host tensors that receive the encoder outputs round-robin (a
    yielded buffer is overwritten ``len(out_bufs)`` batches later); by default a fresh pinned tensor per batch.
    ``depth`` = number of batches in flight.
    """

    def __init__(self, encoder, depth=2):
        p = next(encoder.parameters())
        if not p.is_cuda:
            raise RuntimeError("EncoderPipeline: the encoder must live on a CUDA device (there is no CPU path)")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.encoder = encoder
        self.device = p.device
        self.depth = depth
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.slots = [_Slot() for _ in range(depth)]

    def run(self, batches, out_bufs=None, **forward_kwargs):
        return list(self.stream(batches, out_bufs, **forward_kwargs))

    @torch.no_grad()
    def stream(self, batches, out_bufs=None, **forward_kwargs):
        if out_bufs is not None and len(out_bufs) < self.depth:
            raise ValueError("out_bufs must hold at least `depth` tensors")
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        pending = []
        for i, (feats, lengths) in enumerate(batches):
            slot = self.slots[i % self.depth]
            with torch.cuda.stream(self.h2d):
                if slot.in_free is not None:
                    self.h2d.wait_event(slot.in_free)           # the batch that used this slot has been encoded
                if slot.dev_in is None or slot.dev_in.shape != feats.shape or slot.dev_in.dtype != feats.dtype:
                    slot.dev_in = torch.empty(feats.shape, dtype=feats.dtype, device=dev)
                slot.dev_in.copy_(feats, non_blocking=True)
                lens_d = lengths.to(dev, non_blocking=True)
                slot.in_ready.record(self.h2d)
            lens_d.record_stream(cur)
            cur.wait_event(slot.in_ready)
            out, mask = self.encoder(slot.dev_in, lens_d, **forward_kwargs)
            slot.in_free = torch.cuda.Event()
            slot.in_free.record(cur)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(slot.in_free)
                host = out_bufs[i % len(out_bufs)] if out_bufs is not None else torch.empty(out.shape, dtype=out.dtype).pin_memory()
                host.copy_(out, non_blocking=True)
                out.record_stream(self.d2h)
                done = torch.cuda.Event()
                done.record(self.d2h)
            pending.append((host, mask, done))
            if len(pending) >= self.depth:
                h, m, ev = pending.pop(0)
                ev.synchronize()
                yield h, m
        for h, m, ev in pending:
            ev.synchronize()
            yield h, m
