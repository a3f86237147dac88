"""CPU-only checks of the host side: mask helpers (bit exact vs the reference's goldens), state_dict
layout, the C-ABI library (loads, exports every declared symbol), loud failure without a GPU."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import conformer_pytorch_lightning_b200 as C
from conformer_pytorch_lightning_b200 import _native
from oracle import conformer_oracle as O
from _util import GOLDEN, FWD_CASES, encoder_kwargs, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()
    declared = set(re.findall(r"\b(cfm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _native.lib().cfm_abi_version() == 1


def test_state_dict_layout_matches_reference():
    layouts = json.load(open(os.path.join(GOLDEN, "state_dict_layout.json")))
    for name, rec in layouts.items():
        enc = C.ConformerEncoder(cmvn=None, **encoder_kwargs(rec["cfg"]))
        ours = [[k, list(v.shape), str(v.dtype)] for k, v in enc.state_dict().items()]
        assert ours == rec["keys"], name


def test_cmvn_buffers_in_state_dict():
    class CMVN(torch.nn.Module):          # same buffers as the reference's src/cmvn.py:5-33
        def __init__(self):
            super().__init__()
            self.register_buffer("mean", torch.zeros(80))
            self.register_buffer("istd", torch.ones(80))

        def forward(self, x):
            return (x - self.mean) * self.istd
    cfg = O.conformer_cfg("M", encoder_num_layers=1)
    kw = encoder_kwargs(cfg)
    enc = C.ConformerEncoder(cmvn=CMVN(), **kw)
    assert "global_cmvn.mean" in enc.state_dict() and "global_cmvn.istd" in enc.state_dict()


@pytest.mark.parametrize("name", FWD_CASES)
def test_masks_bit_exact(name):
    g = load_golden(name)
    cfg, fw = g["cfg"], g["fw"]
    lens = torch.from_numpy(g["lens"])
    Tin = g["feats"].shape[1]
    pad = ~C.make_pad_mask(lens, Tin).unsqueeze(1)
    pad = pad[:, :, 2::2][:, :, 2::2]
    assert np.array_equal(pad.numpy(), g["pad_mask"])
    if int(g["torch_seed"]) >= 0:
        torch.manual_seed(int(g["torch_seed"]))
    x = torch.zeros(pad.size(0), pad.size(2), 1)
    attn = C.make_attn_mask(x, pad, cfg["use_dynamic_chunk_size"], cfg["use_dynamic_left_chunk"],
                            fw.get("decoding_chunk_size", 0), cfg["static_chunk_size"],
                            fw.get("num_decoding_chunk_size", -1))
    assert attn.dtype == torch.bool
    assert np.array_equal(attn.numpy(), g["attn_mask"])


def test_chunk_mask_equals_reference_loop():
    for size, c, nl in [(74, 16, -1), (74, 16, 1), (49, 7, 0), (33, 40, 2), (5, 1, 3), (1, 1, -1)]:
        ours = C.subsequent_chunk_mask(size, c, nl, torch.device("cpu")).numpy()
        assert np.array_equal(ours, O.subsequent_chunk_mask(size, c, nl))


def test_positional_tables():
    rel = C.RelativePositionalEncoding(256, 0.0, 5000)
    # numpy's and torch's fp32 exp differ by an ulp in div_term, which position 4999 amplifies to ~5e-4;
    # the low positions every golden case uses agree to fp32 rounding
    assert np.abs(rel.pe.numpy() - O.rel_pos_table(5000, 256)).max() < 1e-3
    assert np.abs(rel.pe.numpy()[:64] - O.rel_pos_table(5000, 256)[:64]).max() < 1e-5
    ab = C.PositionalEncoding(256, 0.0, 5000)
    assert ab.pe.dtype == torch.float16
    x = torch.zeros(3, 7, 256)
    out, pos = rel(x, offset=5)
    assert pos.shape == (3, 1, 256)                 # sliced by batch size
    assert torch.equal(pos, rel.pe[5:8])
    # the reference mutates its table dtype in place; ours must not
    rel(x.to(torch.bfloat16))
    assert rel.pe.dtype == torch.float32


def test_cpu_tensors_raise_no_fallback():
    cfg = O.conformer_cfg("M", encoder_num_layers=1)
    enc = C.ConformerEncoder(cmvn=None, **encoder_kwargs(cfg)).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.randn(1, 100, 80), torch.tensor([100]))
    ffn = C.PositionwiseFeedForwardModule(256, 0.0, 2048).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        ffn(torch.randn(2, 5, 256))


def test_training_on_cpu_raises_too():
    """The training path (autograd + dropout) is native as well: CPU tensors raise, nothing falls back to PyTorch."""
    cfg = O.conformer_cfg("M", encoder_num_layers=1)
    enc = C.ConformerEncoder(cmvn=None, **encoder_kwargs(cfg)).train()
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.randn(1, 100, 80), torch.tensor([100]))
    dec = C.CTCDecoder(50, 256, 0.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        dec(torch.randn(1, 10, 256), torch.tensor([10]), torch.ones(1, 3, dtype=torch.long), torch.tensor([3]))


def test_flat_adam_refuses_cpu_parameters():
    """The optimizer step is native too (cfm_adam_step): CPU / non-fp32 parameters raise instead of falling back."""
    with pytest.raises(NotImplementedError, match="CUDA fp32"):
        C.FlatAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    with pytest.raises(ValueError):
        C.FlatAdam([torch.nn.Parameter(torch.zeros(4))], lr=-1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "conformer_pytorch_lightning_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, fn
