"""Generate the golden fixtures in tests/golden/*.npz by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (Lingeng56/conformer-pytorch-lightning) is pure Python/PyTorch and
cannot travel to the GPU box, so its outputs are frozen here.  Weights are NOT
stored: they come from ``oracle.conformer_oracle.make_state_dict(cfg, seed)``
(numpy RandomState, reproducible anywhere) and are pushed into the reference via
``load_state_dict``.  Inputs come from numpy RandomState too, except the C1 case
which is the kaldi fbank of the reference's own samples/0-3.wav.

Each fixture stores: cfg (json), weight seed, inputs, every output the reference
returns, and the intermediate tensors at the boundary of the measured path
(embed output, pos_embed, pad/attn masks) so the layer stack can be checked
without the PyTorch sub-sampling front-end.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import conformer_oracle as O  # noqa: E402
from encoder import ConformerEncoder  # noqa: E402  (reference)

torch.set_num_threads(8)


def build_ref(cfg, seed):
    enc = ConformerEncoder(input_dim=cfg["input_dim"], kernel_size=cfg["kernel_size"],
                           encoder_dim=cfg["encoder_dim"], dropout=cfg["dropout"],
                           attention_dropout=cfg["attention_dropout"],
                           pos_enc_dropout=cfg["pos_enc_dropout"], hidden_dim=cfg["hidden_dim"],
                           num_heads=cfg["num_heads"], encoder_num_layers=cfg["encoder_num_layers"],
                           cmvn=None, max_len=cfg["max_len"], use_relative=cfg["use_relative"],
                           use_dynamic_chunk_size=cfg["use_dynamic_chunk_size"],
                           use_dynamic_left_chunk=cfg["use_dynamic_left_chunk"],
                           static_chunk_size=cfg["static_chunk_size"])
    sd = O.make_state_dict(cfg, seed)
    ref_keys = list(enc.state_dict().keys())
    assert set(ref_keys) == set(sd.keys()), (set(ref_keys) ^ set(sd.keys()))
    enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return enc.eval(), ref_keys


def hooks(enc):
    """Capture the tensors at the boundary of the measured path (encoder.py:72)."""
    cap = {}

    def embed_hook(_m, _inp, out):
        cap["embed_out"] = out[0].detach().numpy().copy()
        cap["pos_embed"] = out[1].detach().numpy().copy()
        cap["pad_mask"] = out[2].detach().numpy().copy()

    def layer0_hook(_m, inp, _out):
        cap["attn_mask"] = inp[1].detach().numpy().copy()

    h1 = enc.embed.register_forward_hook(embed_hook)
    h2 = enc.encoders[0].register_forward_hook(layer0_hook)
    return cap, (h1, h2)


def save(name, cfg, seed, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, cfg=json.dumps(cfg), weight_seed=seed, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} kB  " +
          " ".join(f"{k}{tuple(np.asarray(v).shape)}" for k, v in arrays.items()))


def case_forward(name, cfg, seed, feats, lens, torch_seed=None, **fw):
    enc, _ = build_ref(cfg, seed)
    cap, hs = hooks(enc)
    draws = [-1, -1]
    if torch_seed is not None:
        torch.manual_seed(torch_seed)
        # replay the generator to record what utils.make_attn_mask will draw (utils.py:131,139)
        g = torch.get_rng_state()
        T = ((feats.shape[1] - 1) // 2 - 1) // 2
        draws[0] = torch.randint(1, T, (1,)).item()
        c = draws[0]
        if not c > T // 2 and cfg["use_dynamic_left_chunk"]:
            draws[1] = torch.randint(0, T - 1, (1,)).item()
        torch.set_rng_state(g)
    with torch.no_grad():
        out, mask = enc(torch.from_numpy(feats), torch.from_numpy(lens), **fw)
    for h in hs:
        h.remove()
    save(name, cfg, seed, feats=feats, lens=lens, out=out.numpy(), out_mask=mask.numpy(),
         draws=np.asarray(draws), torch_seed=-1 if torch_seed is None else torch_seed,
         fw=json.dumps(fw), **cap)
    return enc, out


def main():
    rs = np.random.RandomState(1234)
    feats2 = rs.standard_normal((2, 200, 80)).astype(np.float32)
    lens2 = np.asarray([200, 150], dtype=np.int32)

    M = O.conformer_cfg("M")
    # FP1 of SURVEY appendix A: pad mask only, full 12 layers
    case_forward("m12_pad", M, 0, feats2, lens2)
    # FP2: static chunk 16
    M3 = O.conformer_cfg("M", encoder_num_layers=3)
    case_forward("m3_static16", dict(M3, static_chunk_size=16), 1, feats2, lens2)
    # limited left context + padding -> fully masked rows
    feats3 = rs.standard_normal((3, 300, 80)).astype(np.float32)
    lens3 = np.asarray([300, 222, 131], dtype=np.int32)
    case_forward("m3_left1", dict(M3, use_dynamic_chunk_size=True), 2, feats3, lens3,
                 decoding_chunk_size=16, num_decoding_chunk_size=1)
    # dynamic chunk + dynamic left chunks under a torch seed
    for ts in (7, 11):
        case_forward(f"m3_dyn_seed{ts}", dict(M3, use_dynamic_chunk_size=True, use_dynamic_left_chunk=True),
                     3, feats3, lens3, torch_seed=ts)
    # absolute-position variant (attention.py:105-179)
    case_forward("m3_abs", dict(M3, use_relative=False), 4, feats2, lens2)
    # Conformer-L geometry (d=512, 8 heads, k=31), 2 layers
    L2 = O.conformer_cfg("L", encoder_num_layers=2)
    case_forward("l2_pad", L2, 5, feats3, lens3)

    # streaming: three consecutive forward_chunk calls, for each cache policy (encoder.py:78-123)
    enc, _ = build_ref(M3, 6)
    x1 = rs.standard_normal((1, 67 + 64 * 2, 80)).astype(np.float32)
    for req in (-1, 0, 16):
        arrays = {}
        cache = torch.zeros((0, 0, 0, 0))
        cnn = torch.zeros((0, 0, 0, 0))
        offset = 0
        with torch.no_grad():
            for i in range(3):
                chunk = torch.from_numpy(x1[:, i * 64: i * 64 + 67])
                o, cache, cnn = enc.forward_chunk(chunk, offset, req, cache, cnn)
                offset += o.size(1)
                arrays[f"out{i}"] = o.numpy()
                arrays[f"cache{i}"] = cache.numpy()
                assert tuple(cnn.shape) == (3, 0, 0, 0)
        tag = {-1: "all", 0: "none", 16: "16"}[req]
        save(f"m3_stream_{tag}", M3, 6, feats=x1, required_cache_size=req, **arrays)
    with torch.no_grad():
        o, m = enc.forward_chunk_by_chunk(torch.from_numpy(x1), 16, -1)
        o2, _ = enc.forward_chunk_by_chunk(torch.from_numpy(x1), 8, 2)
    save("m3_chunk_by_chunk", M3, 6, feats=x1, out_c16=o.numpy(), out_c8_l2=o2.numpy(), mask=m.numpy())
    # same for the absolute-position model
    enc, _ = build_ref(dict(M3, use_relative=False), 4)
    with torch.no_grad():
        o, _ = enc.forward_chunk_by_chunk(torch.from_numpy(x1), 16, -1)
    save("m3_abs_chunk_by_chunk", dict(M3, use_relative=False), 4, feats=x1, out_c16=o.numpy())

    # C1: BASELINE.json configs[0] -- fbank of the reference's own sample wavs
    import scipy.io.wavfile as wavfile
    import torchaudio.compliance.kaldi as kaldi
    mats = []
    for i in range(4):
        sr, wav = wavfile.read(f"/root/reference/samples/{i}.wav")
        assert sr == 16000 and wav.dtype == np.int16
        w = torch.from_numpy(wav.astype(np.float32))[None]         # processor.py:183: waveform * (1 << 15)
        mats.append(kaldi.fbank(w, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
                                energy_floor=0.0, sample_frequency=16000))
    mats.sort(key=lambda m: -m.size(0))                            # processor.py:292-297
    lens = np.asarray([m.size(0) for m in mats], dtype=np.int32)
    feats = torch.nn.utils.rnn.pad_sequence(mats, batch_first=True).numpy()
    # un-normalised fbank has mean ~10; the reference would apply CMVN from a stats file that is not in
    # the repo -> use a fixed per-utterance-independent affine so activations stay in a sane range
    feats = ((feats - 10.0) / 4.0).astype(np.float32) * (np.arange(feats.shape[1])[None, :, None] < lens[:, None, None])
    feats = feats.astype(np.float32)
    enc, out = case_forward("m12_c1_wav", M, 0, feats, lens)
    crs = np.random.RandomState(99)
    ctc_w = crs.uniform(-1 / 16, 1 / 16, size=(5002, 256)).astype(np.float32)
    ctc_b = crs.uniform(-1 / 16, 1 / 16, size=(5002,)).astype(np.float32)
    logits = torch.nn.functional.linear(out, torch.from_numpy(ctc_w), torch.from_numpy(ctc_b))
    best = logits.argmax(-1).numpy()
    top2 = logits.topk(2, dim=-1).values
    print("C1 ctc min top1-top2 margin", float((top2[..., 0] - top2[..., 1]).min()))
    np.savez_compressed(os.path.join(HERE, "m12_c1_ctc.npz"), ctc_seed=99, best=best.astype(np.int16))

    # training-mode forward: BatchNorm batch statistics + running-stat update (convolution.py:44)
    cfg_t = dict(M3, dropout=0.0, attention_dropout=0.0, pos_enc_dropout=0.0, static_chunk_size=16)
    enc, _ = build_ref(cfg_t, 8)
    enc.train()
    out, mask = enc(torch.from_numpy(feats3), torch.from_numpy(lens3))
    after = {k: v.detach().numpy() for k, v in enc.state_dict().items() if "norm.running" in k or "num_batches" in k}
    save("m3_train_fwd", cfg_t, 8, feats=feats3, lens=lens3, out=out.detach().numpy(),
         **{k.replace(".", "__"): v for k, v in after.items()})


def grad_sample(g, cap=2048):
    """Strided subsample of a gradient tensor (keeps the fixture small) -- the test re-applies the same stride."""
    flat = np.asarray(g, dtype=np.float32).reshape(-1)
    stride = max(1, flat.size // cap)
    return flat[::stride].copy()


def case_train_grad():
    """BASELINE configs[4] in miniature: training-mode forward + CTC loss + backward through the REAL reference encoder
    and its CTCDecoder (decoder.py:7-23), static chunk-16 attention mask, every dropout 0 (module.py:49-69 is
    loss.backward() on exactly this graph).  Stores the loss, the encoder output, and for EVERY parameter the gradient's
    L2 norm and a strided subsample."""
    from decoder import CTCDecoder  # reference
    cfg = O.conformer_cfg("M", encoder_num_layers=2, hidden_dim=512, dropout=0.0, attention_dropout=0.0, pos_enc_dropout=0.0,
                          static_chunk_size=16)
    seed, V, Lmax = 21, 300, 9
    rs = np.random.RandomState(77)
    feats = rs.standard_normal((3, 299, 80)).astype(np.float32)
    lens = np.asarray([299, 250, 173], dtype=np.int32)
    labels = rs.randint(1, V, size=(3, Lmax)).astype(np.int64)
    labels[1, 4] = labels[1, 3]
    lab_len = np.asarray([9, 7, 4], dtype=np.int64)
    ctc_w = rs.uniform(-1 / 16, 1 / 16, size=(V, 256)).astype(np.float32)
    ctc_b = rs.uniform(-1 / 16, 1 / 16, size=(V,)).astype(np.float32)
    enc, _ = build_ref(cfg, seed)
    enc.train()
    dec = CTCDecoder(V, 256, 0.0)
    dec.load_state_dict({"ctc_lo.weight": torch.from_numpy(ctc_w), "ctc_lo.bias": torch.from_numpy(ctc_b)})
    out, mask = enc(torch.from_numpy(feats), torch.from_numpy(lens))
    out_lens = mask.squeeze(1).sum(1)
    loss = dec(out, out_lens, torch.from_numpy(labels), torch.from_numpy(lab_len))
    loss.backward()
    arrays = {}
    for k, p in list(enc.named_parameters()) + [("ctc." + k, p) for k, p in dec.named_parameters()]:
        key = k.replace(".", "__")
        g = p.grad.detach().numpy()
        arrays["gn__" + key] = np.float64(np.linalg.norm(g.astype(np.float64)))
        arrays["gs__" + key] = grad_sample(g)
    after = {k.replace(".", "__"): v.detach().numpy() for k, v in enc.state_dict().items() if "norm.running" in k}
    save("m2_train_grad", cfg, seed, feats=feats, lens=lens, labels=labels, lab_len=lab_len, ctc_w=ctc_w, ctc_b=ctc_b,
         loss=np.float64(loss.item()), out=out.detach().numpy(), out_lens=out_lens.numpy(), **arrays, **after)


def case_fbank():
    """The reference's feature call (processor.py:185-191) on its own samples/0.wav and 1.wav, dither 0."""
    import scipy.io.wavfile as wavfile
    import torchaudio.compliance.kaldi as kaldi
    arrays = {}
    for i in range(2):
        sr, wav = wavfile.read(f"/root/reference/samples/{i}.wav")
        assert sr == 16000 and wav.dtype == np.int16
        w = torch.from_numpy(wav.astype(np.float32))[None]         # == torchaudio waveform * (1 << 15)
        feat = kaldi.fbank(w, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0,
                           sample_frequency=16000)
        arrays[f"wav{i}"] = wav
        arrays[f"fbank{i}"] = feat.numpy()
    path = os.path.join(HERE, "fbank_wav01.npz")
    np.savez_compressed(path, **arrays)
    print("fbank_wav01:", os.path.getsize(path) // 1000, "kB", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fbank":
        case_fbank()
    elif len(sys.argv) > 1 and sys.argv[1] == "train_grad":
        case_train_grad()
    else:
        main()
        case_train_grad()
        case_fbank()
