"""Size-independent properties of the CUDA path at BASELINE.json's full sizes: utterance independence, batch-permutation
equivariance, chunk-mask causality, padding invariance, determinism.  (Oracle parity at the same full sizes is in
tests/test_gpu_fullsize_parity.py.)  pytest -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import conformer_oracle as O
from _util import build_encoder


def _inputs(B, Tin, seed=0, ragged=False):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, Tin, 80, generator=g).cuda()
    if ragged:
        lens = torch.randint(Tin // 2, Tin + 1, (B,), generator=g).sort(descending=True).values
        lens[0] = Tin
    else:
        lens = torch.full((B,), Tin)
    return feats, lens.to(torch.int32).cuda()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_c2_utterance_independence_and_permutation(dtype):
    """C2 shape (64 x 10 s, T = 248): an utterance's output depends on that utterance only (eval BatchNorm, per-utterance
    attention / depthwise conv), so (a) a sub-batch reproduces the same rows bit for bit although it is tiled
    differently, and (b) permuting the batch permutes the output."""
    cfg = O.conformer_cfg("M", encoder_num_layers=4 if dtype == torch.float32 else 12)
    enc = build_encoder(cfg, 0, compute_dtype=dtype)
    B = 64 if dtype == torch.bfloat16 else 8
    feats, lens = _inputs(B, 998, ragged=True)
    with torch.no_grad():
        full, mask = enc(feats, lens)
        sub, _ = enc(feats[5:9].contiguous(), lens[5:9].contiguous())
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
        pout, pmask = enc(feats[perm].contiguous(), lens[perm].contiguous())
        again, _ = enc(feats, lens)
    assert torch.isfinite(full).all()
    assert torch.equal(full, again)                                   # deterministic (no atomics on the path)
    assert torch.equal(pmask, mask[perm]) and torch.equal(pout, full[perm])
    Ts = sub.shape[1]
    if dtype == torch.float32:
        # the fp32 positional term of row b is pe[b]: a per-row constant that softmax cancels exactly
        assert torch.equal(sub, full[5:9, :Ts])
    else:
        assert torch.equal(sub, full[5:9, :Ts])


def test_c4_padding_invariance_long_form():
    """C4 shape (60 s utterances, T = 1498, attention dominated): valid frames of a padded utterance equal the frames
    obtained when the same utterance is run alone at its own length (key-padding mask + conv masking)."""
    cfg = O.conformer_cfg("M", encoder_num_layers=3)
    enc = build_encoder(cfg, 3, compute_dtype=torch.bfloat16)
    feats, lens = _inputs(4, 5998, seed=2, ragged=True)
    with torch.no_grad():
        full, mask = enc(feats, lens)
        n = int(lens[2])
        alone, m1 = enc(feats[2:3, :n].contiguous(), lens[2:3].contiguous())
    valid = int(m1.sum())
    assert valid == int(mask[2].sum())
    err = (alone[0, :valid].float() - full[2, :valid].float()).abs().max() / full[2, :valid].abs().max()
    # the depthwise conv sees GLU(bias) instead of zero padding right after the last valid frame when the utterance is
    # padded (reference semantics, convolution.py:36-43), so the last (k-1)/2 * layers frames may differ slightly
    k_reach = 7 * 3
    err_inner = (alone[0, :valid - k_reach].float() - full[2, :valid - k_reach].float()).abs().max() / full.abs().max()
    assert float(err_inner) < 2e-2, float(err_inner)
    assert float(err) < 0.5


def test_c5_chunk_mask_causality():
    """C5-style static chunk-16 attention mask at 16 x 10 s: perturbing frames of chunk c must not change the output of
    earlier chunks beyond the reach of the (non-causal) depthwise conv, and later chunks must change."""
    cfg = O.conformer_cfg("M", encoder_num_layers=2, static_chunk_size=16)
    enc = build_encoder(cfg, 5, compute_dtype=torch.bfloat16)
    feats, lens = _inputs(16, 998, seed=4)
    with torch.no_grad():
        a, _ = enc(feats, lens)
        f2 = feats.clone()
        f2[:, 700:] += 1.0                                  # encoder frames >= (700-3)/4 ~ 174 and their conv halo
        b, _ = enc(f2, lens)
    # first perturbed encoder frame: 174.  Layer 1: its chunk (160-175) sees it through attention, the k=15 depthwise
    # conv then reaches back to 153; layer 2: chunk 144-159 contains 153.. -> attention touches 144.., conv reaches 137.
    assert torch.equal(a[:, :128], b[:, :128])
    assert not torch.equal(a[:, 176:], b[:, 176:])


def test_conformer_l_shape_runs():
    """C3 geometry (d=512, 8 heads, k=31): finite outputs, deterministic, mask bit-exact."""
    cfg = O.conformer_cfg("L", encoder_num_layers=2)
    enc = build_encoder(cfg, 7, compute_dtype=torch.bfloat16)
    feats, lens = _inputs(8, 1998, seed=6, ragged=True)
    with torch.no_grad():
        a, m = enc(feats, lens)
        b, _ = enc(feats, lens)
    assert tuple(a.shape) == (8, 498, 512) and torch.isfinite(a).all() and torch.equal(a, b)
    ref_mask = (torch.arange(1998, device="cuda")[None] < lens[:, None])[:, None, :][:, :, 2::2][:, :, 2::2]
    assert torch.equal(m, ref_mask)
