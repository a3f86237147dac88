"""Feature front-end on device (scope row f1 remainder): native Kaldi fbank + global CMVN against the reference's own
feature call frozen in tests/golden/fbank_wav01.npz (torchaudio.compliance.kaldi.fbank on samples/0-1.wav,
processor.py:185-191) and against oracle/fbank_oracle.py.  pytest -m gpu."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import conformer_pytorch_lightning_b200 as C
from conformer_pytorch_lightning_b200 import _native
from oracle import fbank_oracle as FB
from _util import GOLDEN, max_rel

FEAT_TOL = 1e-4          # fp32 max-rel, like the encoder's fp32 gate


def _batch():
    z = np.load(os.path.join(GOLDEN, "fbank_wav01.npz"))
    waves = [z["wav0"].astype(np.float32), z["wav1"].astype(np.float32)]
    n = np.asarray([len(w) for w in waves], dtype=np.int32)
    pad = np.zeros((2, n.max()), np.float32)
    for i, w in enumerate(waves):
        pad[i, :len(w)] = w
    return z, torch.from_numpy(pad).cuda(), torch.from_numpy(n).cuda()


def test_fbank_matches_reference_features_and_oracle():
    z, wave, n = _batch()
    fb = C.Fbank().cuda()
    before = _native.kernel_launches("fbank_frames")
    feats, frames = fb(wave, n)
    assert _native.kernel_launches("fbank_frames") == before + 1
    assert frames.tolist() == [z["fbank0"].shape[0], z["fbank1"].shape[0]]
    assert tuple(feats.shape) == (2, z["fbank0"].shape[0], 80)
    for i in range(2):
        m = z[f"fbank{i}"].shape[0]
        got = feats[i, :m].cpu().numpy()
        assert max_rel(got, z[f"fbank{i}"]) < FEAT_TOL                       # the reference's own feature call
        assert max_rel(got, FB.fbank(z[f"wav{i}"].astype(np.float32))) < FEAT_TOL
        assert float(feats[i, m:].abs().max()) == 0.0 if m < feats.shape[1] else True      # pad_sequence zeros


def test_cmvn_module_and_fused_cmvn(tmp_path):
    z, wave, n = _batch()
    rs = np.random.RandomState(0)
    stats = {"mean_stat": (rs.uniform(8, 12, 80) * 1000).tolist(), "var_stat": (rs.uniform(110, 150, 80) * 1000).tolist(),
             "frame_num": 1000}
    path = tmp_path / "cmvn.json"
    path.write_text(json.dumps(stats))
    cmvn = C.GlobalCMVN(str(path)).cuda()
    mean, istd = FB.load_cmvn_stats(stats)
    assert np.allclose(cmvn.mean.cpu().numpy(), mean, rtol=1e-6) and np.allclose(cmvn.istd.cpu().numpy(), istd, rtol=1e-6)
    assert list(cmvn.state_dict().keys()) == ["mean", "istd"]                # cmvn.py:18-19 buffers
    plain, _ = C.Fbank().cuda()(wave, n)
    fused, _ = C.Fbank(cmvn=cmvn).cuda()(wave, n)
    sep = cmvn(plain)
    assert max_rel(sep.cpu().numpy(), FB.cmvn(plain.cpu().numpy(), mean, istd)) < 1e-6
    assert max_rel(fused.cpu().numpy(), sep.cpu().numpy()) < 1e-6            # padding frames: (0 - mean) * istd in both
    novar = C.GlobalCMVN.from_stats(mean, istd, norm_var=False).cuda()
    assert max_rel(novar(plain).cpu().numpy(), plain.cpu().numpy() - mean) < 1e-6
    with pytest.raises(RuntimeError, match="CUDA"):
        cmvn(plain.cpu())


def test_fbank_feeds_the_encoder():
    """wave -> native fbank (+CMVN) -> ConformerEncoder equals the encoder on the oracle's features of the same audio."""
    from oracle import conformer_oracle as O
    from _util import build_encoder
    z, wave, n = _batch()
    mean = np.full(80, 10.0, np.float32)
    istd = np.full(80, 0.25, np.float32)
    cmvn = C.GlobalCMVN.from_stats(mean, istd).cuda()
    feats, frames = C.Fbank(cmvn=cmvn).cuda()(wave, n)
    cfg = O.conformer_cfg("M", encoder_num_layers=2)
    enc = build_encoder(cfg, 1, compute_dtype=torch.float32)
    with torch.no_grad():
        out, mask = enc(feats, frames.to(torch.int32))
    ref_feats = np.zeros(tuple(feats.shape), np.float32)
    for i in range(2):
        f = FB.fbank(z[f"wav{i}"].astype(np.float32))
        ref_feats[i, :len(f)] = f
    ref_feats = FB.cmvn(ref_feats, mean, istd).astype(np.float32)
    ref, ref_mask, _ = O.encoder_forward(ref_feats, frames.cpu().numpy().astype(np.int32), O.make_state_dict(cfg, 1), cfg)
    assert np.array_equal(mask.cpu().numpy(), ref_mask)
    assert max_rel(out.cpu().numpy(), ref) < 2e-4
