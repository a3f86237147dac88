"""Shared helpers for the test-suite (fixture loading, error metric)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "cfg" in d:
        d["cfg"] = json.loads(str(d["cfg"]))
    if "fw" in d:
        d["fw"] = json.loads(str(d["fw"]))
    if "weight_seed" in d:
        d["weight_seed"] = int(d["weight_seed"])
    return d


def max_rel(a, b):
    """SURVEY 8c tolerance metric: max |a-b| / max |b| over the whole tensor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


FWD_CASES = ["m12_pad", "m3_static16", "m3_left1", "m3_dyn_seed7", "m3_dyn_seed11", "m3_abs", "l2_pad",
             "m12_c1_wav"]


def encoder_kwargs(cfg):
    keys = ["input_dim", "kernel_size", "encoder_dim", "dropout", "attention_dropout", "pos_enc_dropout",
            "hidden_dim", "num_heads", "encoder_num_layers", "max_len", "use_relative", "use_dynamic_chunk_size",
            "use_dynamic_left_chunk", "static_chunk_size"]
    return {k: cfg[k] for k in keys}


def build_encoder(cfg, seed, device="cuda", compute_dtype=None):
    """Our drop-in encoder with the oracle's seeded weights loaded through load_state_dict."""
    import torch
    import conformer_pytorch_lightning_b200 as C
    from oracle import conformer_oracle as O
    enc = C.ConformerEncoder(cmvn=None, **encoder_kwargs(cfg))
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in O.make_state_dict(cfg, seed).items()}
    enc.load_state_dict(sd)
    enc = enc.to(device).eval()
    if compute_dtype is not None:
        enc.set_compute_dtype(compute_dtype)
    return enc
