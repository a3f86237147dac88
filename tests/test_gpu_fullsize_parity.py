"""Encoder-level parity of the CUDA path against the oracle AT BASELINE.json's FULL SIZES (configs[1..3]):
C2 = Conformer-M 64 x 10 s (T = 248), C3 = Conformer-L 17 layers 32 x 20 s (T = 498), C4 = Conformer-M 16 x 60 s ragged
(T = 1498, padding masks).  The checker is oracle/conformer_oracle_torch.py (the ATen-CPU restatement that reproduces the
reference's golden outputs exactly, tests/test_oracle_golden.py); it finishes these sizes in seconds on the host cores.
Compared over ALL positions (padded rows included); masks bit exact; the kernel families that must have
served the call are asserted through cfm_kernel_launches.  Mirrors the loop of src/encoder.py:72-74.  pytest -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import conformer_oracle as O
from oracle import conformer_oracle_torch as OT
from _util import build_encoder, max_rel
from conformer_pytorch_lightning_b200 import _native

FP32_TOL = 1e-4      # north_star: fp32 max-rel 1e-4
BF16_TOL = 2e-2      # north_star: bf16 max-rel 2e-2

# name: (cfg, B, Tin, ragged, utterances per oracle group)
SIZES = {
    "C2": ("M", 64, 998, False, 16),
    "C3": ("L", 32, 1998, False, 8),
    "C4": ("M", 16, 5998, True, 2),
}
_cache = {}


def _case(name):
    """Inputs + oracle outputs of one full-size case (computed once per session)."""
    if name in _cache:
        return _cache[name]
    cfg_name, B, Tin, ragged, group = SIZES[name]
    cfg = O.conformer_cfg(cfg_name)
    rs = np.random.RandomState(1234)
    feats = rs.standard_normal((B, Tin, 80)).astype(np.float32)
    if ragged:
        lens = np.sort(rs.randint(Tin // 2, Tin + 1, size=B))[::-1].copy()
        lens[0] = Tin
    else:
        lens = np.full((B,), Tin)
    lens = lens.astype(np.int32)
    sd = OT.to_torch_sd(O.make_state_dict(cfg, 0))
    torch.set_num_threads(max(1, torch.get_num_threads()))
    out, pad, attn, x, pos = OT.encoder_forward(torch.from_numpy(feats), torch.from_numpy(lens), sd, cfg, batch_chunk=group)
    _cache[name] = dict(cfg=cfg, feats=feats, lens=lens, out=out.numpy(), pad=pad.numpy(), attn=attn.numpy(), x=x.numpy(),
                        pos=pos.numpy())
    return _cache[name]


def _counts(names):
    return {n: _native.kernel_launches(n) for n in names}


FUSED = ("ffn_fused", "mhsa_fused", "conv_fused", "attention_tc", "attention_pp", "gemm_tc", "gemm_simt", "dwconv")


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_fullsize_bf16_measured_path_vs_oracle(name):
    """encode_layers (the measured path) in bf16 on the oracle's own fp32 boundary tensors."""
    c = _case(name)
    enc = build_encoder(c["cfg"], 0, compute_dtype=torch.bfloat16)
    t = lambda a: torch.from_numpy(a).cuda()
    enc.use_cuda_graphs = False
    before = _counts(FUSED)
    with torch.no_grad():
        out = enc.encode_layers(t(c["x"]), t(c["attn"]), t(c["pos"]), t(c["pad"]))      # one eager pass, counted
    torch.cuda.synchronize()
    after = _counts(FUSED)
    enc.use_cuda_graphs = True
    with torch.no_grad():
        outs = [enc.encode_layers(t(c["x"]), t(c["attn"]), t(c["pos"]), t(c["pad"])) for _ in range(3)]  # eager, capture, replay
    ran = {k: after[k] - before[k] for k in FUSED}
    err = max_rel(out.cpu().numpy(), c["out"])
    print(f"{name}: bf16 measured path max-rel {err:.4f}; kernels per pass {ran}")
    assert err < BF16_TOL
    assert all(torch.equal(out, o) for o in outs)     # the benchmarked graph replay is the same computation
    L = c["cfg"]["encoder_num_layers"]
    assert ran["gemm_simt"] == 0                      # nothing fell back to the CUDA-core GEMM
    if name in ("C2", "C4"):
        # per pass: 13 chained feed-forward launches (12 layers) and one fused convolution module per layer
        assert ran["ffn_fused"] == L + 1 and ran["conv_fused"] == L and ran["dwconv"] == 0
    if name == "C2":
        assert ran["mhsa_fused"] == L and ran["attention_tc"] + ran["attention_pp"] == 0
    if name == "C4" and ran["mhsa_fused"] == 0:
        assert ran["attention_pp"] == L               # long sequences: ping-pong flash kernel + residual GEMM


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_fullsize_bf16_forward_vs_oracle(name):
    """ConformerEncoder.forward from fbank features (native front-end on the bf16 path) vs the oracle's full forward."""
    c = _case(name)
    enc = build_encoder(c["cfg"], 0, compute_dtype=torch.bfloat16)
    with torch.no_grad():
        out, mask = enc(torch.from_numpy(c["feats"]).cuda(), torch.from_numpy(c["lens"]).cuda())
    assert np.array_equal(mask.cpu().numpy(), c["pad"])                    # bit exact
    err = max_rel(out.cpu().numpy(), c["out"])
    print(f"{name}: bf16 forward max-rel {err:.4f}")
    assert err < BF16_TOL


@pytest.mark.parametrize("name,B", [("C2", 64), ("C3", 8), ("C4", 4)])
def test_fullsize_fp32_forward_vs_oracle(name, B):
    """fp32 path (CUDA-core engines, exact-order accumulation) at the full sequence lengths; C3 / C4 at a reduced batch
    (the fp32 engines are the parity engines, not the fast ones)."""
    c = _case(name)
    enc = build_encoder(c["cfg"], 0, compute_dtype=torch.float32)
    with torch.no_grad():
        out, mask = enc(torch.from_numpy(c["feats"][:B]).cuda(), torch.from_numpy(c["lens"][:B]).cuda())
    assert np.array_equal(mask.cpu().numpy(), c["pad"][:B])
    # the relative-position row of utterance b is pe[b] in both runs (D2) and T is the global maximum (lens[0] = Tin)
    err = max_rel(out.cpu().numpy(), c["out"][:B])
    print(f"{name}: fp32 forward max-rel {err:.2e}")
    assert err < FP32_TOL
