"""Module/encoder-level parity of the CUDA path against (a) the committed golden outputs of the real
reference and (b) the numpy oracle on the same seeded inputs.  pytest -m gpu (B200 box)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import conformer_oracle as O
from _util import FWD_CASES, build_encoder, load_golden, max_rel

FP32_TOL = 1e-4      # north_star: fp32 max-rel 1e-4
BF16_TOL = 2e-2      # north_star: bf16 max-rel 2e-2


def _run_forward(g, dtype):
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=dtype)
    if int(g["torch_seed"]) >= 0:
        torch.manual_seed(int(g["torch_seed"]))
    fw = g["fw"]
    with torch.no_grad():
        out, mask = enc(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["lens"]).cuda(), **fw)
    return enc, out, mask


@pytest.mark.parametrize("name", FWD_CASES)
def test_encoder_forward_fp32_vs_reference_golden(name):
    g = load_golden(name)
    _, out, mask = _run_forward(g, torch.float32)
    assert out.dtype == torch.float32 and tuple(out.shape) == g["out"].shape
    assert np.array_equal(mask.cpu().numpy(), g["out_mask"])           # bit exact
    assert max_rel(out.cpu().numpy(), g["out"]) < FP32_TOL              # all positions, padded rows too (D11)


@pytest.mark.parametrize("name", FWD_CASES)
def test_encoder_forward_bf16_vs_reference_golden(name):
    g = load_golden(name)
    _, out, mask = _run_forward(g, torch.bfloat16)
    assert np.array_equal(mask.cpu().numpy(), g["out_mask"])
    err = max_rel(out.cpu().numpy(), g["out"])
    print(f"{name}: bf16 max-rel {err:.4f}")
    assert err < BF16_TOL


@pytest.mark.parametrize("name", ["m12_pad", "m3_static16", "m3_left1", "l2_pad"])
def test_measured_path_only(name):
    """encode_layers (= encoder.py:72-74) fed with the reference's own boundary tensors."""
    g = load_golden(name)
    enc = build_encoder(g["cfg"], g["weight_seed"])
    t = lambda a: torch.from_numpy(a).cuda()
    with torch.no_grad():
        out = enc.encode_layers(t(g["embed_out"]), t(g["attn_mask"]), t(g["pos_embed"]), t(g["pad_mask"]))
    assert max_rel(out.cpu().numpy(), g["out"]) < FP32_TOL


def test_ctc_greedy_ids_bit_exact_fp32():
    g = load_golden("m12_c1_wav")
    c = load_golden("m12_c1_ctc")
    _, out, mask = _run_forward(g, torch.float32)
    crs = np.random.RandomState(int(c["ctc_seed"]))
    w = torch.from_numpy(crs.uniform(-1 / 16, 1 / 16, size=(5002, 256)).astype(np.float32)).cuda()
    b = torch.from_numpy(crs.uniform(-1 / 16, 1 / 16, size=(5002,)).astype(np.float32)).cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    best = torch.nn.functional.linear(out, w, b).argmax(-1).cpu().numpy().astype(np.int16)
    assert np.array_equal(best, c["best"])
    # the native CTC head (fp32 path: CUDA-core GEMM through the chunked logit workspace) must give the same frame ids
    from conformer_pytorch_lightning_b200 import CTCGreedyHead
    head = CTCGreedyHead(256, 5002).cuda()
    with torch.no_grad():
        head.ctc_lo.weight.copy_(w)
        head.ctc_lo.bias.copy_(b)
        ids, hyps = head.greedy(out, mask[:, 0, :].sum(-1))
    assert np.array_equal(ids.cpu().numpy().astype(np.int16), c["best"])
    assert len(hyps) == out.shape[0]
    # bf16: report the agreement rate (reference's own bf16 agrees with its fp32 on ~94 % of frames, D9)
    _, outb, _ = _run_forward(g, torch.bfloat16)
    bestb = torch.nn.functional.linear(outb, w, b).argmax(-1).cpu().numpy().astype(np.int16)
    valid = mask.cpu().numpy()[:, 0, :]
    rate = float((bestb == c["best"])[valid].mean())
    print(f"bf16 CTC frame-id agreement with the fp32 reference: {rate:.4f}")
    assert rate > 0.85
    head.compute_dtype = torch.bfloat16                   # tcgen05 GEMM with the argmax epilogue
    with torch.no_grad():
        idsb = head.frame_ids(outb).cpu().numpy().astype(np.int16)
    rate_native = float((idsb == c["best"])[valid].mean())
    print(f"bf16 encoder + native bf16 CTC head agreement with the fp32 reference: {rate_native:.4f}")
    assert rate_native > 0.85


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("tag", ["all", "none", "16"])
def test_streaming_forward_chunk(tag, dtype, tol):
    g = load_golden("m3_stream_" + tag)
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=dtype)
    req = int(g["required_cache_size"])
    cache = torch.zeros((0, 0, 0, 0))
    cnn = torch.zeros((0, 0, 0, 0))
    offset = 0
    feats = torch.from_numpy(g["feats"]).cuda()
    for i in range(3):
        with torch.no_grad():
            o, cache, cnn = enc.forward_chunk(feats[:, i * 64:i * 64 + 67], offset, req, cache, cnn)
        offset += o.size(1)
        assert tuple(o.shape) == g[f"out{i}"].shape
        assert tuple(cache.shape) == g[f"cache{i}"].shape
        assert tuple(cnn.shape) == (3, 0, 0, 0)
        assert max_rel(o.cpu().numpy(), g[f"out{i}"]) < tol
        if cache.numel():
            assert max_rel(cache.cpu().numpy(), g[f"cache{i}"]) < tol


def test_forward_chunk_by_chunk():
    g = load_golden("m3_chunk_by_chunk")
    enc = build_encoder(g["cfg"], g["weight_seed"])
    feats = torch.from_numpy(g["feats"]).cuda()
    with torch.no_grad():
        o, m = enc.forward_chunk_by_chunk(feats, 16, -1)
        o2, _ = enc.forward_chunk_by_chunk(feats, 8, 2)
        # the reference's greedy_search passes a 1-element length tensor as the chunk size (model.py:206-209)
        o3, _ = enc.forward_chunk_by_chunk(feats, torch.tensor([16]))
    assert tuple(m.shape) == g["mask"].shape and m.dtype == torch.float32
    assert max_rel(o.cpu().numpy(), g["out_c16"]) < FP32_TOL
    assert max_rel(o2.cpu().numpy(), g["out_c8_l2"]) < FP32_TOL
    assert max_rel(o3.cpu().numpy(), g["out_c16"]) < FP32_TOL
    g = load_golden("m3_abs_chunk_by_chunk")
    enc = build_encoder(g["cfg"], g["weight_seed"])
    with torch.no_grad():
        o, _ = enc.forward_chunk_by_chunk(torch.from_numpy(g["feats"]).cuda(), 16, -1)
    assert max_rel(o.cpu().numpy(), g["out_c16"]) < FP32_TOL


def test_training_mode_forward_batchnorm():
    g = load_golden("m3_train_fwd")
    enc = build_encoder(g["cfg"], g["weight_seed"]).train()
    with torch.no_grad():
        out, _ = enc(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["lens"]).cuda())
    assert max_rel(out.cpu().numpy(), g["out"]) < FP32_TOL
    sd = enc.state_dict()
    for i in range(3):
        key = f"encoders.{i}.conv_module.norm."
        gk = key.replace(".", "__")
        assert max_rel(sd[key + "running_mean"].cpu().numpy(), g[gk + "running_mean"]) < 1e-4
        assert max_rel(sd[key + "running_var"].cpu().numpy(), g[gk + "running_var"]) < 1e-4
        assert int(sd[key + "num_batches_tracked"]) == int(g[gk + "num_batches_tracked"])


def test_submodules_vs_oracle():
    """Layer / attention / conv / ffn modules called directly (module-level API), vs the numpy oracle."""
    cfg = O.conformer_cfg("M", encoder_num_layers=1)
    sd = O.make_state_dict(cfg, 21)
    enc = build_encoder(cfg, 21)
    layer = enc.encoders[0]
    rs = np.random.RandomState(5)
    x = rs.standard_normal((2, 40, 256)).astype(np.float32)
    lens = np.array([40, 29])
    pad = (np.arange(40)[None, :] < lens[:, None])[:, None, :]
    pos = O.rel_pos_table(5000, 256)[:2]
    t = lambda a: torch.from_numpy(a).cuda()
    p = "encoders.0."
    with torch.no_grad():
        y = layer.feed_forward(t(x))
        assert max_rel(y.cpu().numpy(), O.feed_forward(x, sd, p + "feed_forward.")) < FP32_TOL
        y, c = layer.conv_module(t(x), t(pad))
        assert max_rel(y.cpu().numpy(), O.conv_module(x, pad, sd, p + "conv_module.")) < FP32_TOL
        assert tuple(c.shape) == (0, 0, 0)
        y, cache = layer.self_attn(t(x), t(x), t(x), t(pad), t(pos))
        ry, rc = O.rel_mhsa(x, pad, pos, None, sd, p + "self_attn.", 4)
        assert max_rel(y.cpu().numpy(), ry) < FP32_TOL and max_rel(cache.cpu().numpy(), rc) < FP32_TOL
        # float masks follow `.eq(0)` semantics; empty sentinel = no mask
        y2, _ = layer.self_attn(t(x), t(x), t(x), t(pad.astype(np.float32)), t(pos))
        assert torch.equal(y, y2)
        y3, _ = layer.self_attn(t(x), t(x), t(x), torch.ones((0, 0, 0), device="cuda"), t(pos))
        ry3, _ = O.rel_mhsa(x, None, pos, None, sd, p + "self_attn.", 4)
        assert max_rel(y3.cpu().numpy(), ry3) < FP32_TOL
        out, am, nc, cc = layer(t(x), t(pad), t(pos), t(pad))
        ro, rcache = O.encoder_layer(x, pad, pos, pad, None, sd, p, cfg)
        assert max_rel(out.cpu().numpy(), ro) < FP32_TOL and max_rel(nc.cpu().numpy(), rcache) < FP32_TOL
        assert tuple(cc.shape) == (0, 0, 0)


def test_weights_refresh_after_load_state_dict():
    """Derived (bf16 / folded) weights are caches: loading new parameters must invalidate them."""
    g = load_golden("m3_static16")
    enc = build_encoder(g["cfg"], 123, compute_dtype=torch.bfloat16)       # wrong weights first
    feats, lens = torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["lens"]).cuda()
    with torch.no_grad():
        bad, _ = enc(feats, lens)
    assert max_rel(bad.cpu().numpy(), g["out"]) > 0.1
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in O.make_state_dict(g["cfg"], g["weight_seed"]).items()}
    enc.load_state_dict(sd)
    with torch.no_grad():
        good, _ = enc(feats, lens)
    assert max_rel(good.cpu().numpy(), g["out"]) < BF16_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cuda_graph_replay_equals_eager(dtype):
    """encode_layers captures a CUDA graph on the 2nd call of a shape; replays must equal eager bit for bit,
    follow new inputs / masks, and pick up new weights."""
    g = load_golden("m3_left1")
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=dtype)
    eager = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=dtype)
    eager.use_cuda_graphs = False
    feats = torch.from_numpy(g["feats"]).cuda()
    lens = torch.from_numpy(g["lens"]).cuda()
    fw = g["fw"]
    for it in range(4):
        f = feats if it % 2 == 0 else feats.flip(0) * 0.5
        l = lens if it < 2 else torch.tensor([300, 150, 299], dtype=torch.int32, device="cuda")
        with torch.no_grad():
            a, ma = enc(f, l, **fw)
            b, mb = eager(f, l, **fw)
        assert torch.equal(ma, mb)
        assert torch.equal(a, b), f"iteration {it}"
    assert any(p.get("graph") is not None for p in enc._plans.values())
    # new weights must be honoured by the captured plan
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in O.make_state_dict(g["cfg"], 777).items()}
    enc.load_state_dict(sd)
    eager.load_state_dict(sd)
    with torch.no_grad():
        a, _ = enc(feats, lens, **fw)
        b, _ = eager(feats, lens, **fw)
    assert torch.equal(a, b)


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_encoder_pipeline_matches_forward(depth):
    """EncoderPipeline overlaps the copies of neighbouring batches with compute; every batch must come back exactly as
    a plain forward call returns it, in order, also when host output buffers are recycled."""
    from conformer_pytorch_lightning_b200 import EncoderPipeline
    g = load_golden("m3_static16")
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=torch.bfloat16)
    feats, lens = torch.from_numpy(g["feats"]), torch.from_numpy(g["lens"])
    batches = [((feats * (1.0 + 0.1 * i)).pin_memory(), lens) for i in range(7)]
    want = []
    with torch.no_grad():
        for f, l in batches:
            o, m = enc(f.cuda(), l.cuda(), **g["fw"])
            want.append((o.cpu(), m.cpu()))
    pipe = EncoderPipeline(enc, depth=depth)
    got = pipe.run(batches, **g["fw"])
    assert len(got) == len(want)
    for (o, m), (wo, wm) in zip(got, want):
        assert o.is_pinned() and torch.equal(o, wo) and torch.equal(m.cpu(), wm)
    bufs = [torch.empty_like(want[0][0]).pin_memory() for _ in range(depth)]
    n = 0
    for (o, m), (wo, wm) in zip(pipe.stream(batches, bufs, **g["fw"]), want):
        assert torch.equal(o, wo), n                  # checked before the buffer is recycled
        n += 1
    assert n == len(want)
    assert max_rel(want[0][0].numpy(), g["out"]) < BF16_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_streaming_graph_replay_equals_eager(dtype):
    """forward_chunk captures a CUDA graph per (chunk, cache size, trim point); with 2 left chunks the steady-state
    shape repeats, so a 3-pass run exercises eager, capture and replay.  Replays must equal eager bit for bit."""
    g = load_golden("m3_chunk_by_chunk")
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=dtype)
    eager = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=dtype)
    eager.use_cuda_graphs = False
    feats = torch.from_numpy(g["feats"]).cuda()[:1]
    with torch.no_grad():
        want, _ = eager.forward_chunk_by_chunk(feats, 16, 2)
        for it in range(3):
            got, _ = enc.forward_chunk_by_chunk(feats, 16, 2)
            assert torch.equal(got, want), f"pass {it}"
    assert any(k[0] == "chunk" and p.get("graph") is not None for k, p in enc._plans.items() if isinstance(k, tuple))


def test_compute_dtype_alternation_on_one_encoder():
    """a captured bf16 graph must survive a switch to fp32 and back -- the derived bf16 weight copies
    are kept per dtype, so the graph's baked-in addresses stay valid."""
    g = load_golden("m3_static16")
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=torch.bfloat16)
    feats, lens = torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["lens"]).cuda()
    with torch.no_grad():
        bf = [enc(feats, lens)[0] for _ in range(3)]                 # eager, capture, replay
        enc.set_compute_dtype(torch.float32)
        f32 = [enc(feats, lens)[0] for _ in range(3)]
        junk = [torch.randn(1 << 20, device="cuda") for _ in range(64)]   # churn the allocator
        enc.set_compute_dtype(torch.bfloat16)
        bf2 = [enc(feats, lens)[0] for _ in range(3)]
        enc.set_compute_dtype(torch.float32)
        f32b = enc(feats, lens)[0]
    del junk
    assert all(torch.equal(bf[0], o) for o in bf + bf2)
    assert all(torch.equal(f32[0], o) for o in f32 + [f32b])
    assert max_rel(f32[0].cpu().numpy(), g["out"]) < FP32_TOL
    assert max_rel(bf2[2].cpu().numpy(), g["out"]) < BF16_TOL


def test_graph_plans_are_lru_bounded():
    """ADVICE r1 (medium) / VERDICT weak #9: plans are evicted least-recently-used, and unbounded-left-context streaming
    does not create per-chunk plans."""
    cfg = O.conformer_cfg("M", encoder_num_layers=2)
    enc = build_encoder(cfg, 1, compute_dtype=torch.bfloat16)
    enc.max_plans = 4
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for tin in (300, 340, 380, 420, 460, 500, 300, 300):
            feats = torch.randn(2, tin, 80, generator=g).cuda()
            enc(feats, torch.full((2,), tin, dtype=torch.int32).cuda())
        assert len(enc._plans) <= 4
        n_before = len(enc._plans)
        out, _ = enc.forward_chunk_by_chunk(torch.randn(1, 67 + 64 * 5, 80, generator=g).cuda(), 16, -1)
        assert len(enc._plans) == n_before                       # growing cache: no plan per chunk index
        out2, _ = enc.forward_chunk_by_chunk(torch.randn(1, 67 + 64 * 5, 80, generator=g).cuda(), 16, 1)
        assert len(enc._plans) <= 4
    assert torch.isfinite(out).all() and torch.isfinite(out2).all()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_feedforward_relu_and_cross_attention_surface(dtype, tol):
    """The two corners of the reference's module surface outside the encoder's own use: activation='relu'
    (feedforward.py:10-11) and attention with key / value different from query (attention.py:62-64), vs the same
    arithmetic in PyTorch fp32."""
    import conformer_pytorch_lightning_b200 as C
    torch.manual_seed(3)
    ffn = C.PositionwiseFeedForwardModule(256, 0.0, 1024, activation='relu').cuda().eval()
    ffn.compute_dtype = dtype
    x = torch.randn(3, 70, 256, device="cuda")
    with torch.no_grad():
        got = ffn(x)
        ref = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(x, ffn.w_1.weight, ffn.w_1.bias)),
                                         ffn.w_2.weight, ffn.w_2.bias)
    assert max_rel(got.cpu().numpy(), ref.cpu().numpy()) < tol
    att = C.MultiHeadSelfAttentionModule(256, 4, 0.0).cuda().eval()
    att.compute_dtype = dtype
    q, kv = torch.randn(2, 70, 256, device="cuda"), torch.randn(2, 90, 256, device="cuda")
    mask = torch.ones(2, 1, 90, dtype=torch.bool, device="cuda")
    mask[1, :, 77:] = False
    with torch.no_grad():
        got, cache = att(q, kv, kv, mask)
        lin = torch.nn.functional.linear
        qq = lin(q, att.linear_q.weight, att.linear_q.bias).view(2, 70, 4, 64).transpose(1, 2)
        kk = lin(kv, att.linear_k.weight, att.linear_k.bias).view(2, 90, 4, 64).transpose(1, 2)
        vv = lin(kv, att.linear_v.weight, att.linear_v.bias).view(2, 90, 4, 64).transpose(1, 2)
        s = (qq @ kk.transpose(-1, -2)) / 8.0
        m = ~mask[:, None]
        p = torch.softmax(s.masked_fill(m, float("-inf")), -1).masked_fill(m, 0.0)
        ref = lin((p @ vv).transpose(1, 2).reshape(2, 70, 256), att.linear_out.weight, att.linear_out.bias)
    assert tuple(cache.shape) == (2, 4, 90, 128)
    assert max_rel(got.cpu().numpy(), ref.cpu().numpy()) < tol


def test_output_dtype_knob():
    """encoder.output_dtype = bf16: the returned states are the fp32 states rounded to bf16 (eager and graph replay)."""
    g = load_golden("m3_static16")
    enc = build_encoder(g["cfg"], g["weight_seed"], compute_dtype=torch.bfloat16)
    feats, lens = torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["lens"]).cuda()
    with torch.no_grad():
        ref = [enc(feats, lens)[0] for _ in range(3)]
        enc.output_dtype = torch.bfloat16
        got = [enc(feats, lens)[0] for _ in range(3)]
    assert got[0].dtype == torch.bfloat16 and ref[0].dtype == torch.float32
    assert all(torch.equal(o, ref[0].to(torch.bfloat16)) for o in got)
