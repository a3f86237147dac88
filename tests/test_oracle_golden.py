"""Pin the numpy oracle against outputs of the real reference (tests/golden/*.npz,
made by tests/golden/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest

from oracle import conformer_oracle as O
from _util import FWD_CASES, load_golden, max_rel

TOL = 2e-5          # fp32 numpy vs fp32 torch/MKL: different summation orders only


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_matches_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    if name == "m12_c1_wav":            # keep the CPU suite short: 12 layers x 1244 tokens is still < 10 s
        pass
    sd = O.make_state_dict(cfg, g["weight_seed"])
    fw = g["fw"]
    out, pad_mask, attn_mask = O.encoder_forward(
        g["feats"], g["lens"], sd, cfg, draws=g["draws"],
        decoding_chunk_size=fw.get("decoding_chunk_size", 0),
        num_decoding_left_chunks=fw.get("num_decoding_chunk_size", -1))
    # integer / bool work is bit exact
    assert np.array_equal(pad_mask, g["out_mask"])
    assert np.array_equal(pad_mask, g["pad_mask"])
    assert np.array_equal(attn_mask, g["attn_mask"])
    assert out.shape == g["out"].shape
    assert max_rel(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["m12_pad", "m3_static16", "l2_pad"])
def test_embed_boundary(name):
    """The tensors that enter the measured path (encoder.py:72) match."""
    g = load_golden(name)
    sd = O.make_state_dict(g["cfg"], g["weight_seed"])
    x, pos, pad, attn = O.encoder_embed(g["feats"], g["lens"], sd, g["cfg"])
    assert max_rel(x, g["embed_out"]) < TOL
    assert max_rel(pos, g["pos_embed"]) < 1e-6
    # and the layer stack alone, fed with the reference's own boundary tensors
    out = O.encoder_layers(g["embed_out"], g["attn_mask"], g["pos_embed"], g["pad_mask"], sd, g["cfg"])
    assert max_rel(out, g["out"]) < TOL


@pytest.mark.parametrize("tag", ["all", "none", "16"])
def test_streaming_chunks(tag):
    g = load_golden("m3_stream_" + tag)
    cfg = g["cfg"]
    sd = O.make_state_dict(cfg, g["weight_seed"])
    req = int(g["required_cache_size"])
    cache = np.zeros((0, 0, 0, 0), np.float32)
    offset = 0
    for i in range(3):
        o, cache = O.encoder_forward_chunk(g["feats"][:, i * 64:i * 64 + 67], offset, req, cache, sd, cfg)
        offset += o.shape[1]
        assert o.shape == g[f"out{i}"].shape
        assert cache.shape == g[f"cache{i}"].shape
        assert max_rel(o, g[f"out{i}"]) < TOL
        if cache.size:
            assert max_rel(cache, g[f"cache{i}"]) < TOL


def test_chunk_by_chunk():
    g = load_golden("m3_chunk_by_chunk")
    sd = O.make_state_dict(g["cfg"], g["weight_seed"])
    o = O.encoder_forward_chunk_by_chunk(g["feats"], 16, -1, sd, g["cfg"])
    assert max_rel(o, g["out_c16"]) < TOL
    o = O.encoder_forward_chunk_by_chunk(g["feats"], 8, 2, sd, g["cfg"])
    assert max_rel(o, g["out_c8_l2"]) < TOL
    g = load_golden("m3_abs_chunk_by_chunk")
    sd = O.make_state_dict(g["cfg"], g["weight_seed"])
    o = O.encoder_forward_chunk_by_chunk(g["feats"], 16, -1, sd, g["cfg"])
    assert max_rel(o, g["out_c16"]) < TOL


def test_ctc_greedy_ids_bit_exact():
    """fp32 CTC greedy frame ids equal the reference's (north_star gate)."""
    g = load_golden("m12_c1_wav")
    c = load_golden("m12_c1_ctc")
    crs = np.random.RandomState(int(c["ctc_seed"]))
    w = crs.uniform(-1 / 16, 1 / 16, size=(5002, 256)).astype(np.float32)
    b = crs.uniform(-1 / 16, 1 / 16, size=(5002,)).astype(np.float32)
    best, hyps = O.ctc_greedy_ids(g["out"], [311, 246, 145, 119], w, b)
    assert np.array_equal(best.astype(np.int16), c["best"])
    assert len(hyps) == 4


def test_mask_closed_form_equals_loop():
    """visible(i,j) closed form == utils.py:96-111 loop."""
    for size, c, nl in [(74, 16, -1), (74, 16, 1), (49, 7, 0), (33, 40, 2), (5, 1, 3)]:
        ref = O.subsequent_chunk_mask(size, c, nl)
        i = np.arange(size)[:, None]
        j = np.arange(size)[None, :]
        vis = (j < (i // c + 1) * c) & ((nl < 0) | (j >= (i // c - nl) * c))
        assert np.array_equal(ref, vis)


def test_training_forward_batchnorm():
    """Batch statistics over all B*T positions unmasked + running-stat update."""
    g = load_golden("m3_train_fwd")
    cfg = g["cfg"]
    sd = O.make_state_dict(cfg, g["weight_seed"])
    x, pos, pad, attn = O.encoder_embed(g["feats"], g["lens"], sd, cfg)
    for i in range(cfg["encoder_num_layers"]):
        p = f"encoders.{i}."
        # re-implement the layer with training-mode conv module
        half = np.float32(0.5)
        y = O.layer_norm(x, sd[p + "norm_ff_macaron.weight"], sd[p + "norm_ff_macaron.bias"])
        x = x + half * O.feed_forward(y, sd, p + "feed_forward_macaron.")
        y = O.layer_norm(x, sd[p + "norm_mha.weight"], sd[p + "norm_mha.bias"])
        y, _ = O.rel_mhsa(y, attn, pos, None, sd, p + "self_attn.", cfg["num_heads"])
        x = x + y
        y = O.layer_norm(x, sd[p + "norm_conv.weight"], sd[p + "norm_conv.bias"])
        st = {"running_mean": sd[p + "conv_module.norm.running_mean"].copy(),
              "running_var": sd[p + "conv_module.norm.running_var"].copy(),
              "num_batches_tracked": 3}
        x = x + O.conv_module(y, pad, sd, p + "conv_module.", training=True, bn_state=st)
        y = O.layer_norm(x, sd[p + "norm_ff.weight"], sd[p + "norm_ff.bias"])
        x = x + half * O.feed_forward(y, sd, p + "feed_forward.")
        x = O.layer_norm(x, sd[p + "norm_final.weight"], sd[p + "norm_final.bias"])
        key = f"encoders__{i}__conv_module__norm__"
        assert max_rel(st["running_mean"], g[key + "running_mean"]) < 1e-5
        assert max_rel(st["running_var"], g[key + "running_var"]) < 1e-5
        assert int(g[key + "num_batches_tracked"]) == st["num_batches_tracked"] == 4
    out = O.layer_norm(x, sd["after_norm.weight"], sd["after_norm.bias"])
    assert max_rel(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["m12_pad", "m3_static16", "m3_left1", "l2_pad"])
def test_torch_cpu_port_matches_reference(name):
    """The multi-threaded CPU port used as bench.py's cpu_baseline is pinned to the same goldens."""
    import torch
    from oracle import conformer_oracle_torch as OT
    g = load_golden(name)
    sd = OT.to_torch_sd(O.make_state_dict(g["cfg"], g["weight_seed"]))
    t = torch.from_numpy
    out = OT.encoder_layers(t(g["embed_out"]), t(g["attn_mask"]), t(g["pos_embed"]), t(g["pad_mask"]), sd, g["cfg"])
    assert max_rel(out.numpy(), g["out"]) < TOL


@pytest.mark.parametrize("name", ["m12_pad", "m3_static16", "l2_pad", "m12_c1_wav"])
def test_torch_cpu_port_full_forward_matches_reference(name):
    """The port's full forward (front-end + masks + layers), which is the checker of the full-size GPU parity tests
    (tests/test_gpu_fullsize_parity.py), reproduces the reference's outputs -- also when run in utterance groups."""
    import torch
    from oracle import conformer_oracle_torch as OT
    g = load_golden(name)
    sd = OT.to_torch_sd(O.make_state_dict(g["cfg"], g["weight_seed"]))
    out, pad, attn, x, pos = OT.encoder_forward(torch.from_numpy(g["feats"]), torch.from_numpy(g["lens"]), sd, g["cfg"],
                                                batch_chunk=2)
    assert np.array_equal(pad.numpy(), g["out_mask"]) and np.array_equal(attn.numpy(), g["attn_mask"])
    assert max_rel(x.numpy(), g["embed_out"]) < TOL and max_rel(pos.numpy(), g["pos_embed"]) < 1e-6
    assert max_rel(out.numpy(), g["out"]) < TOL


def test_torch_cpu_port_training_forward():
    """Training-mode port (BatchNorm batch statistics + running-stat updates) against the reference's golden."""
    import torch
    from oracle import conformer_oracle_torch as OT
    g = load_golden("m3_train_fwd")
    cfg = g["cfg"]
    sd = OT.to_torch_sd(O.make_state_dict(cfg, g["weight_seed"]))
    x, pos, pad, attn = OT.encoder_embed(torch.from_numpy(g["feats"]), torch.from_numpy(g["lens"]), sd, cfg)
    st = [{"running_mean": sd[f"encoders.{i}.conv_module.norm.running_mean"].clone(),
           "running_var": sd[f"encoders.{i}.conv_module.norm.running_var"].clone(), "num_batches_tracked": 3}
          for i in range(cfg["encoder_num_layers"])]
    with torch.no_grad():
        out = OT.encoder_layers_train(x, attn, pos, pad, sd, cfg, bn_states=st)
    assert max_rel(out.numpy(), g["out"]) < TOL
    for i, s in enumerate(st):
        key = f"encoders__{i}__conv_module__norm__"
        assert max_rel(s["running_mean"].numpy(), g[key + "running_mean"]) < 1e-5
        assert max_rel(s["running_var"].numpy(), g[key + "running_var"]) < 1e-5
        assert s["num_batches_tracked"] == 4


def grad_sample(g, cap=2048):
    flat = np.asarray(g, dtype=np.float32).reshape(-1)
    return flat[::max(1, flat.size // cap)]


def test_torch_cpu_port_training_gradients_match_reference():
    """CTC loss + every parameter gradient of the training-mode port equal the REAL reference's loss.backward()
    (tests/golden/m2_train_grad.npz: norms and strided subsamples of all gradients)."""
    import torch
    from oracle import conformer_oracle_torch as OT
    g = load_golden("m2_train_grad")
    cfg = g["cfg"]
    sd = OT.to_torch_sd(O.make_state_dict(cfg, g["weight_seed"]))
    t = torch.from_numpy
    loss, out, grads, cgrads, st = OT.train_step_grads(t(g["feats"]), t(g["lens"]), t(g["labels"]), t(g["lab_len"]), sd,
                                                       t(g["ctc_w"]), t(g["ctc_b"]), cfg)
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert max_rel(out.numpy(), g["out"]) < TOL
    n = 0
    for k, gr in list(grads.items()) + [("ctc." + k, v) for k, v in cgrads.items()]:
        key = k.replace(".", "__")
        ref_s, ref_n = g["gs__" + key], float(g["gn__" + key])
        if ref_n < 1e-5:                      # pos_bias_v / linear_pos: zero up to rounding
            assert float(gr.norm()) < 1e-4
            continue
        assert abs(float(gr.double().norm()) - ref_n) < 1e-3 * ref_n, k
        assert max_rel(grad_sample(gr.numpy()), ref_s) < 2e-3, k
        n += 1
    assert n >= 70
    for i, s in enumerate(st):
        assert max_rel(s["running_mean"].numpy(), g[f"encoders__{i}__conv_module__norm__running_mean"]) < 1e-5


def test_fbank_oracle_matches_reference_feature_call():
    """oracle/fbank_oracle.py vs torchaudio.compliance.kaldi.fbank as the reference calls it (processor.py:185-191) on the
    reference's samples/0.wav and 1.wav (tests/golden/fbank_wav01.npz); load_cmvn / cmvn vs their definitions."""
    from oracle import fbank_oracle as FB
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fbank_wav01.npz"))
    for i in range(2):
        got = FB.fbank(z[f"wav{i}"].astype(np.float32))
        assert got.shape == z[f"fbank{i}"].shape
        assert max_rel(got, z[f"fbank{i}"]) < 1e-4
    mean, istd = FB.load_cmvn_stats({"mean_stat": [10.0, 20.0], "var_stat": [60.0, 250.0], "frame_num": 2})
    assert np.allclose(mean, [5.0, 10.0]) and np.allclose(istd, [1 / np.sqrt(5.0), 1 / np.sqrt(25.0)])
    assert np.allclose(FB.cmvn(np.asarray([[6.0, 15.0]]), mean, istd), [[1 / np.sqrt(5.0), 1.0]])
