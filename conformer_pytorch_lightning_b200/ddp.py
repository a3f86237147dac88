"""Data-parallel gradient synchronisation for the native training path (one process per GPU, NCCL over NVLink).

The reference trains with Lightning's DDP strategy (executor.py:137-154): every rank holds a replica, gradients are
averaged across ranks.  Here the encoder's native backward (training.py) produces all gradients of a layer in one flat
fp32 bucket and hands it to ``GradSync.bucket_ready`` the moment that layer's kernels are enqueued; the all-reduce of
layer i runs on NCCL's stream while the compute stream continues with layers i-1, ...  ``finish`` (called at the end of
the backward) makes the compute stream wait for all of them, so the gradients autograd then accumulates into
``param.grad`` are already averaged.  Parameters outside the layer stack (sub-sampling front-end, CTC head) are
synchronised by ``sync_grads`` after ``loss.backward()``.  BatchNorm statistics stay per rank like the reference's
(non-Sync) BatchNorm1d.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group=None, average=True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        self._pending = []
        self.buckets_sent = 0
        self.bytes_sent = 0

    def bucket_ready(self, flat):
        """Launch the all-reduce of one flat gradient buffer (asynchronous with respect to the compute stream)."""
        if self.world == 1:
            return
        if self.average:
            flat.mul_(1.0 / self.world)
        self._pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.buckets_sent += 1
        self.bytes_sent += flat.numel() * flat.element_size()

    def finish(self):
        """The current stream waits for every outstanding all-reduce (device-side dependency, no host sync on NCCL)."""
        for h in self._pending:
            h.wait()
        self._pending = []


def attach(encoder, process_group=None, average=True):
    """Give ``encoder`` (ConformerEncoder) a GradSync: its backward then all-reduces each layer's gradients as soon as
    they are complete."""
    encoder.grad_sync = GradSync(process_group, average)
    return encoder.grad_sync


def sync_grads(params, process_group=None, average=True):
    """All-reduce the gradients of ``params`` (those NOT covered by an attached GradSync: front-end, heads) in one
    flat bucket."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    if average:
        flat.mul_(1.0 / dist.get_world_size(process_group))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=process_group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def broadcast_parameters(module, src=0, process_group=None):
    """Rank ``src``'s parameters and buffers to every rank (what DDP does at construction)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=process_group)
