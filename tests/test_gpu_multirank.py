"""Sharded-batch multi-GPU parity on hardware (SURVEY 8e "Parity"): a length-sorted batch stride-sharded over N GPUs,
each rank encoding its shard with the global T_max kept, outputs gathered over NCCL, equals the single-GPU output BIT FOR
BIT (same kernels, per-utterance data, tile shapes independent of the batch size).  Needs >= 2 visible GPUs (skipped on
a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _inputs():
    rs = np.random.RandomState(5)
    B, tin = 12, 998
    feats = rs.standard_normal((B, tin, 80)).astype(np.float32)
    lens = np.sort(rs.randint(tin // 2, tin + 1, size=B))[::-1].copy()
    lens[0] = tin
    return feats, lens.astype(np.int32)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import conformer_oracle as O
    from _util import build_encoder
    from conformer_pytorch_lightning_b200 import sharding
    cfg = O.conformer_cfg("M", encoder_num_layers=3)
    enc = build_encoder(cfg, 2, device=f"cuda:{rank}", compute_dtype=torch.bfloat16)
    feats, lens = _inputs()
    idx = sharding.shard_indices(len(lens), rank, world)
    with torch.no_grad():
        # parity mode: every shard keeps the global T_max (all shards are padded to feats.shape[1])
        out, mask = enc(torch.from_numpy(feats[idx.numpy()]).cuda(), torch.from_numpy(lens[idx.numpy()]).cuda())
        full = sharding.all_gather_outputs(out, idx.cuda(), len(lens))
    if rank == 0:
        q.put(full.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_forward_equals_single_gpu_bitwise():
    from oracle import conformer_oracle as O
    from _util import build_encoder
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = O.conformer_cfg("M", encoder_num_layers=3)
    enc = build_encoder(cfg, 2, compute_dtype=torch.bfloat16)
    feats, lens = _inputs()
    with torch.no_grad():
        ref, _ = enc(torch.from_numpy(feats).cuda(), torch.from_numpy(lens).cuda())
    assert got.shape == tuple(ref.shape)
    assert np.array_equal(got, ref.cpu().numpy())          # bitwise
