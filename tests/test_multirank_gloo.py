"""world_size-2 gloo test (CPU) of the utterance sharding used for multi-GPU runs: stride-assign a
length-sorted batch to ranks, every rank handles only its shard, shards re-assembled in the original order equal
the single-rank result.  The per-shard compute here is the oracle port (no GPU in this container)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from conformer_pytorch_lightning_b200 import sharding  # noqa: E402


def _worker(rank, world, port, feats, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import conformer_oracle as O
    cfg = O.conformer_cfg("M", encoder_num_layers=1)
    sd = O.make_state_dict(cfg, 3)
    idx = sharding.shard_indices(len(lens), rank, world)
    # parity mode: keep the global T_max so padded rows and shapes match the single-process result
    out, mask, _ = O.encoder_forward(feats[idx], lens[idx], sd, cfg)
    gathered = sharding.all_gather_outputs(torch.from_numpy(out), idx, len(lens))
    if rank == 0:
        q.put(gathered.numpy())
    dist.destroy_process_group()


def test_shard_indices_balance_and_cover():
    for n, w in [(64, 8), (16, 2), (7, 2), (5, 8), (128, 8)]:
        seen = []
        for r in range(w):
            idx = sharding.shard_indices(n, r, w)
            assert list(idx) == list(range(r, n, w))
            seen += list(idx)
        assert sorted(seen) == list(range(n))


def test_two_rank_sharded_forward_equals_single_rank():
    from oracle import conformer_oracle as O
    rs = np.random.RandomState(0)
    feats = rs.standard_normal((4, 120, 80)).astype(np.float32)
    lens = np.asarray([120, 111, 90, 64], dtype=np.int32)          # sorted descending like processor.py:292-297
    cfg = O.conformer_cfg("M", encoder_num_layers=1)
    ref, _, _ = O.encoder_forward(feats, lens, O.make_state_dict(cfg, 3), cfg)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, feats, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5


def _grad_sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conformer_pytorch_lightning_b200 import ddp
    sync = ddp.GradSync(average=True)
    buckets = [torch.full((5,), float(rank + 1) * (i + 1)) for i in range(3)]      # "layer" buckets, last layer first
    for b in reversed(buckets):
        sync.bucket_ready(b)
    sync.finish()
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank))
    ddp.broadcast_parameters(lin, src=0)
    lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
    lin.bias.grad = None
    ddp.sync_grads(lin.parameters())
    q.put((rank, [b.tolist() for b in buckets], lin.weight.detach().clone().tolist(), lin.weight.grad.tolist(), sync.buckets_sent))
    dist.destroy_process_group()


def test_grad_sync_averages_buckets_world2():
    """ddp.GradSync / sync_grads / broadcast_parameters over gloo, world size 2: every bucket ends as the mean over ranks
    on every rank (the all-reduce the native backward issues per layer)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7
    procs = [ctx.Process(target=_grad_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, buckets, w, g, sent in res:
        for i, b in enumerate(buckets):
            assert b == [1.5 * (i + 1)] * 5                      # mean of (1, 2) * (i + 1)
        assert all(v == 0.0 for row in w for v in row)           # rank 0's weights everywhere
        assert all(v == 1.5 for row in g for v in row)
        assert sent == 3
