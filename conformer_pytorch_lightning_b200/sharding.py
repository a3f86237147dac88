"""Utterance sharding for multi-GPU runs: the encoder path has no cross-utterance coupling (eval BatchNorm uses
running statistics; the reference's training BatchNorm is per-rank, never SyncBN), so a batch is split by utterance
with no data-path collective.  The reference expresses data parallelism the same way, as a rank-strided split of the
sample list (src/dataset.py:19-21,57: ``data[rank::world_size]``)."""
import torch
import torch.distributed as dist


def shard_indices(n_utts, rank, world):
    """Stride-assign a length-sorted batch: rank r gets utterances r, r+world, ...  (similar token count per rank)."""
    if rank >= n_utts:
        return torch.zeros(0, dtype=torch.int64)
    return torch.arange(rank, n_utts, world)


def all_gather_outputs(local_out, local_idx, n_utts):
    """The 'final output gather' of the north_star: every rank ends up with the (n_utts, T, d) tensor in the original
    utterance order.  NCCL on GPUs, gloo on CPU; pads ragged shards to a common size."""
    world = dist.get_world_size()
    per = (n_utts + world - 1) // world
    pad = per - local_out.size(0)
    buf = torch.cat([local_out, local_out.new_zeros((pad,) + tuple(local_out.shape[1:]))]) if pad else local_out
    idx = torch.cat([local_idx, local_idx.new_full((pad,), -1)]) if pad else local_idx
    outs = [torch.empty_like(buf) for _ in range(world)]
    idxs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(outs, buf.contiguous())
    dist.all_gather(idxs, idx.contiguous())
    full = local_out.new_zeros((n_utts,) + tuple(local_out.shape[1:]))
    for o, i in zip(outs, idxs):
        keep = i >= 0
        full[i[keep]] = o[keep]
    return full
