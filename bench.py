#!/usr/bin/env python
"""Headline benchmark: Conformer encoder audio-seconds per wall-second (RTFx) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C4]

Workload (BASELINE.json configs[1], "C2"): Conformer-M (12 layers, d=256, 4 heads, FFN 2048, conv k=15),
bf16 compute, batch 64 x 10 s of synthetic 80-dim fbank (Tin = 998 -> T = 248 after 4x sub-sampling),
random-init weights.  One *step* = one pass of the measured path (layer loop + after_norm,
reference src/encoder.py:72-74) over one batch.  With N GPUs every rank processes its own batch of 64
utterances (weak scaling, no data-path collective); the JSON line is printed by rank 0.

value : RTFx of the measured path with its inputs (sub-sampled features, masks) resident in HBM,
        CUDA events per step on the launch stream, L2 flushed between steps, max over ranks.
e2e   : the same metric through the public API with pinned HOST fbank features: ``EncoderPipeline`` drives
        ``ConformerEncoder.forward(feats, lengths)`` batch after batch (H2D copy + sub-sampling front-end + layers + D2H of
        the output every step), overlapping the copies of neighbouring batches with compute; `serial_latency_ms` is one
        synchronous forward call with both copies.
roofline : dominant kernel = the fused feed-forward kernel (w_1 + SiLU + w_2 + residual + LayerNorm, ~60 % of
        the step), timed in-step with CUDA events around each of its launches; the peak is the measured BURST bf16
        figure when the clocks sampled during the timed region sat at max with no power cap, else the sustained one
        (both fractions are printed); `roofline_kernels` gives the same for the fused attention and convolution-module
        kernels; `roofline_hbm` for the HBM-bound depthwise-conv kernel when it runs (unfused convolution path only).
cpu_baseline / --impl reference : the ATen-CPU oracle port of the reference's CPU path (the reference is
        Python and cannot travel to the GPU box) on a bounded sample of the same workload.
extra.eager_gpu_baseline : the same ATen port as eager PyTorch on this GPU (fp32 TF32-off and bf16 autocast): the
        cuBLAS / cuDNN bar on the same box (report only).
extra.strong_scaling (N > 1) : ONE global batch sharded by utterance + NCCL all-gather of the outputs, host to host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# diagnostic only (never used for a reported number): CFM_BENCH_NOFLUSH=1 keeps L2 warm between steps
NOFLUSH = os.environ.get("CFM_BENCH_NOFLUSH", "0") == "1"

WORKLOADS = {
    # name: (cfg name, batch, seconds, padded lengths?)
    "C2": ("M", 64, 10.0, False),
    "C3": ("L", 32, 20.0, False),
    "C4": ("M", 16, 60.0, True),
}


def frames(seconds):
    tin = int((16000 * seconds - 400) // 160 + 1)
    return tin, ((tin - 1) // 2 - 1) // 2


def make_inputs(workload, seed=1234):
    cfg_name, B, sec, padded = WORKLOADS[workload]
    tin, T = frames(sec)
    rs = np.random.RandomState(seed)
    feats = rs.standard_normal((B, tin, 80)).astype(np.float32)
    if padded:
        lens = np.sort(rs.randint(tin // 2, tin + 1, size=B))[::-1].copy()
        lens[0] = tin
    else:
        lens = np.full((B,), tin)
    return cfg_name, feats, lens.astype(np.int32), T, float(np.sum(((lens.astype(np.float64) - 1) * 160 + 400) / 16000.0))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rtfx(workload, sample_b, repeats, min_seconds=0.0):
    """The CPU port of the reference's algorithm (oracle/conformer_oracle_torch.py: the same ATen CPU
    kernels the reference's modules run, all host threads), measured path only, fp32."""
    import torch
    from oracle import conformer_oracle as O
    from oracle import conformer_oracle_torch as OT
    torch.set_num_threads(os.cpu_count())
    cfg_name, feats, lens, T, _ = make_inputs(workload)
    cfg = O.conformer_cfg(cfg_name)
    sd = O.make_state_dict(cfg, 0)
    sample_b = min(sample_b, feats.shape[0])
    f, l = feats[:sample_b], lens[:sample_b]
    audio = float(np.sum(((l.astype(np.float64) - 1) * 160 + 400) / 16000.0))
    # boundary tensors of the measured path; the sub-sampling front-end is outside it (random stand-in
    # with the right shape/statistics: timing does not depend on the values)
    Ts = ((f.shape[1] - 1) // 2 - 1) // 2
    x = torch.from_numpy(np.random.RandomState(7).standard_normal((sample_b, Ts, cfg["encoder_dim"])).astype(np.float32))
    pad = torch.from_numpy(~O.make_pad_mask(l, f.shape[1])[:, None, :][:, :, 2::2][:, :, 2::2])
    pos = torch.from_numpy(O.rel_pos_table(cfg["max_len"], cfg["encoder_dim"])[:sample_b])
    attn = pad
    sdc = OT.to_torch_sd(sd)
    best = float("inf")
    times = []
    while len(times) < repeats or (sum(times) < min_seconds and len(times) < 40):
        t0 = time.perf_counter()
        OT.encoder_layers(x, attn, pos, pad, sdc, cfg)
        dt = time.perf_counter() - t0
        times.append(dt)
        best = min(best, dt)
    return audio / best, best, sample_b, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    steps = max(1, args.steps)
    # size the per-step sample so that the whole (warmup + steps) run stays within ~2 minutes
    _, probe, _, _ = cpu_port_rtfx(args.workload, 1, 1)
    budget = 120.0 / (args.warmup + steps)
    sb = int(max(1, min(args.cpu_sample, budget // max(probe, 1e-3))))
    rtfx, best, sb, times = cpu_port_rtfx(args.workload, sb, args.warmup + steps)
    timed = times[args.warmup:] if len(times) > args.warmup else times
    ms = 1e3 * float(np.mean(timed))
    cfg_name, feats, lens, T, _ = make_inputs(args.workload)
    audio = float(np.sum(((lens[:sb].astype(np.float64) - 1) * 160 + 400) / 16000.0))
    val = audio / (ms / 1e3)
    sample = f"measured path (12 layers + after_norm) on {sb} of the {feats.shape[0]} utterances per step"
    line = {"impl": "reference", "metric": "encoder audio-sec/sec (RTFx), measured path", "value": val,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: Conformer-{cfg_name} fp32 ATen-CPU port of the reference on the host cores, sample B={sb} x 10 s (T={T})"},
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def pin_rank_cores(local, world):
    """One process per GPU: give every rank its own slice of the host cores (NUMA-node local when the GPU's node is
    known) so that the N Python launch loops and pinned-memory copies do not migrate over each other.  Opt-in
    (CFM_BENCH_PIN=1): measured on the 8-GPU box it did not help (serial latency 3.93 ms pinned vs 3.89 ms unpinned, e2e
    2.03e6 vs 2.19e6 audio-s/s) -- the N = 8 end-to-end limiter is the aggregate PCIe / host-memory traffic, not the
    scheduler."""
    if os.environ.get("CFM_BENCH_PIN", "0") != "1":
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < 2 * world:
            return None
        node_cores = None
        try:
            import torch
            bus = torch.cuda.get_device_properties(local).pci_bus_id
            dom = torch.cuda.get_device_properties(local).pci_domain_id
            dev_id = torch.cuda.get_device_properties(local).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
            node = int(open(path).read().strip())
            if node >= 0:
                txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
                nc = []
                for part in txt.split(","):
                    a, _, b = part.partition("-")
                    nc += list(range(int(a), int(b or a) + 1))
                node_cores = [c for c in nc if c in cores]
        except Exception:
            node_cores = None
        per = len(cores) // world
        mine = cores[local * per:(local + 1) * per]
        if node_cores and len(node_cores) >= per:
            # the ranks that share this NUMA node split its cores among themselves
            k = local % max(1, len(node_cores) // per)
            mine = node_cores[k * per:(k + 1) * per] or mine
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import build_encoder
    from oracle import conformer_oracle as O  # only for cfg + seeded weights + the cpu_baseline leg
    from conformer_pytorch_lightning_b200 import _native, ops
    from conformer_pytorch_lightning_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pinned_cores = pin_rank_cores(local, world)
    real_stdout = None
    if world > 1:
        # NCCL writes its banner / INFO log to stdout whenever it likes; stdout must carry exactly one JSON line, so for
        # the whole multi-rank run file descriptor 1 points at stderr (the NCCL log stays visible there, at whatever
        # NCCL_DEBUG level the caller chose) and the JSON line is written to the saved descriptor at the end.
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        warm = torch.zeros(1, device=torch.device("cuda", local))
        dist.all_reduce(warm)
        torch.cuda.synchronize()
    dev = torch.device("cuda", local)

    cfg_name, feats_np, lens_np, T, audio_s = make_inputs(args.workload, seed=1234 + rank)
    cfg = O.conformer_cfg(cfg_name)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    enc = build_encoder(cfg, 0, device=dev, compute_dtype=dtype)
    B = feats_np.shape[0]

    feats_host = torch.from_numpy(feats_np).pin_memory()
    lens_dev = torch.from_numpy(lens_np).to(dev)
    with torch.no_grad():
        feats_dev = feats_host.to(dev)
        pad = ~enc_make_pad(lens_dev, feats_dev.size(1))
        x_emb, pos, pad = enc.embed(feats_dev, pad)
        from conformer_pytorch_lightning_b200.utils import make_attn_mask
        attn = make_attn_mask(x_emb, pad, False, False, 0, -1, -1)
    x_emb = x_emb.contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        with torch.no_grad():
            return enc.encode_layers(x_emb, attn, pos, pad)

    # the end-to-end legs of the bf16 pipeline return bf16 encoder states (half the device->host bytes)
    out_dt = torch.bfloat16 if dtype == torch.bfloat16 else torch.float32
    out_host = torch.empty((B, T, cfg["encoder_dim"]), dtype=out_dt).pin_memory()

    def e2e_step():
        with torch.no_grad():
            out, mask = enc(feats_host.to(dev, non_blocking=True), lens_dev)
            out_host.copy_(out, non_blocking=True)          # pinned destination; the timed region syncs after it
            torch.cuda.current_stream().synchronize()
            return out_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, wall=False):
        for _ in range(warmup):
            fn()
        barrier()
        total_ms, n0 = 0.0, _native.launch_count() + engine.GRAPH_REPLAYED_LAUNCHES[0]
        t_wall = time.perf_counter()
        for _ in range(steps):
            if not NOFLUSH:
                flush.fill_(1)                  # evict L2 (126 MB) between steps; not inside the event bracket
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if wall:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                total_ms += 1e3 * (time.perf_counter() - t0)
            else:
                s.record()
                fn()
                e.record()
                e.synchronize()
                total_ms += s.elapsed_time(e)
        barrier()
        launches = _native.launch_count() + engine.GRAPH_REPLAYED_LAUNCHES[0] - n0
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, time.perf_counter() - t_wall

    with ClockSampler(local) as clk:
        total_ms, launches, _ = timed(step, args.steps, args.warmup)
    ms_per_step = total_ms / args.steps
    value = world * audio_s / (ms_per_step / 1e3)

    if args.profile:
        if rank == 0:
            out = real_stdout if real_stdout is not None else sys.stdout
            out.write(json.dumps({"profile_only": True, "ms_per_step": ms_per_step, "value": value, "gpu_launches": launches}) + "\n")
            out.flush()
        return
    # serial latency of one forward call (H2D -> front-end -> layers -> D2H -> sync), L2 flushed between calls
    enc.output_dtype = out_dt if out_dt != torch.float32 else None
    lat_ms, _, _ = timed(e2e_step, max(3, args.steps // 4), 3, wall=True)
    lat_ms /= max(3, args.steps // 4)
    # e2e throughput: the same forward calls driven by EncoderPipeline (copies of neighbouring batches overlap the
    # compute of the current one on separate streams).  Every step copies its fbank batch host->device and its
    # encoder output device->host inside the timed region; wall clock from the first submit to the last delivery.
    from conformer_pytorch_lightning_b200 import EncoderPipeline
    pipe = EncoderPipeline(enc, depth=int(os.environ.get("CFM_B200_PIPE_DEPTH", "2")))
    outs = [torch.empty((B, T, cfg["encoder_dim"]), dtype=out_dt).pin_memory() for _ in range(4)]
    lens_host = torch.from_numpy(lens_np)
    e2e_steps = max(8, args.steps)
    for _ in pipe.stream(((feats_host, lens_host) for _ in range(3)), outs):
        pass
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for h, _ in pipe.stream(((feats_host, lens_host) for _ in range(e2e_steps)), outs):
        checksum += float(h[0, 0, 0])                   # the delivered host buffer is read every step
    torch.cuda.synchronize()
    e2e_t = torch.tensor([1e3 * (time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.item())
    e2e_val = world * audio_s / (e2e_ms / 1e3)
    out_bytes = B * T * cfg["encoder_dim"] * (2 if out_dt == torch.bfloat16 else 4)
    enc.output_dtype = None

    # ---- roofline of the dominant kernels, timed in-step with CUDA events around each library call of the step
    #      (eager launches: graph replays cannot be bracketed per kernel).  One entry per call family; the kernel that
    #      actually served the calls is read back from the library's per-kernel launch counters.
    pk = peaks()
    clocks = clk.summary()
    burst_ok = ("sw_power_cap" not in clocks["reasons"] and clocks["sm_mhz"] is not None and clocks["sm_max_mhz"]
                and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    # a kernel timed inside a ~30 ms timed region at full boost clock with no power cap runs against the BURST tensor
    # peak; under a power cap / reduced clocks the sustained figure applies.  Both fractions are printed.
    tf_peak, tf_kind = (pk["tf_burst"], "burst") if burst_ok else (pk["tf_sustained"], "sustained")
    n_tok, d, F, kc = B * T, cfg["encoder_dim"], cfg["hidden_dim"], cfg["kernel_size"]
    traffic = {}
    for tname in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and args.workload == "C2":      # ncu captures were taken on the C2 shapes
            traffic = json.load(open(tpath))
            break
    FAMILIES = ("ffn_fused", "mhsa_fused", "conv_fused", "gemm_tc", "attention_tc", "attention_pp", "dwconv", "gemm_simt", "attention_simt",
                "layernorm")
    probes = {}                                       # call family -> list of (start, end, flop, bytes)

    def flops_of(name, a, kw):
        if name == "ffn":
            return 4.0 * n_tok * d * F
        if name == "ffn_chain":                                    # (y, a, b, x, y_out, proj=...)
            pj = kw.get("proj")
            return (2 if a[1] is not None else 1) * 4.0 * n_tok * d * F + (2.0 * n_tok * d * pj[0].shape[0] if pj is not None else 0.0)
        if name == "mhsa_out":                                     # scores + PV over the full T x T square + linear_out
            return 4.0 * B * T * T * d + 2.0 * n_tok * d * d
        if name == "conv_module":                                  # pw1 (d -> 2d) + pw2 (d -> d) + depthwise taps
            return 6.0 * n_tok * d * d + 2.0 * n_tok * d * kc
        if name == "gemm":
            return 2.0 * a[0].shape[0] * a[1].shape[0] * a[0].shape[1]
        if name == "attention":
            return 4.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[1] * a[0].shape[2] * a[0].shape[3]
        return 0.0

    def wrap(name, key):
        orig = getattr(ops, name)

        def probe(*a, **kw):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            r = orig(*a, **kw)
            e_.record()
            probes.setdefault(key, []).append((s_, e_, flops_of(name, a, kw)))
            return r
        setattr(ops, name, probe)
        return name, orig

    graphs = enc.use_cuda_graphs
    enc.use_cuda_graphs = False
    saved = [wrap("ffn", "ffn"), wrap("ffn_chain", "ffn"), wrap("mhsa_out", "mhsa"), wrap("conv_module", "conv"),
             wrap("dwconv", "dw"), wrap("gemm", "gemm"), wrap("attention", "attn")]
    fam0 = {f: _native.kernel_launches(f) for f in FAMILIES}
    PROBE_PASSES = 3
    try:
        for _ in range(PROBE_PASSES):
            flush.fill_(1)
            # head start for the CPU: the GPU spins ~2.5 ms while all launches of the step are enqueued, so the
            # bracketed intervals are device execution time, not host launch gaps
            torch.cuda._sleep(5_000_000)
            step()
        torch.cuda.synchronize()
    finally:
        for name, orig in saved:
            setattr(ops, name, orig)
        enc.use_cuda_graphs = graphs
    fam = {f: (_native.kernel_launches(f) - fam0[f]) // PROBE_PASSES for f in FAMILIES}

    def tensor_roofline(key, fused_family, fused_label, unfused_label):
        if key not in probes:
            return None
        ts = [s_.elapsed_time(e_) for s_, e_, _ in probes[key]]
        fl = [f for _, _, f in probes[key]]
        ach = sum(fl) / (sum(ts) * 1e-3) / 1e12
        per_step = len(ts) // PROBE_PASSES
        fused = fam.get(fused_family, 0) > 0
        tr = traffic.get(fused_family + "_kernel") if fused else None
        return {"bound": "tensor", "kernel": fused_label if fused else unfused_label,
                "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "frac_burst": ach / pk["tf_burst"], "frac_sustained": ach / pk["tf_sustained"],
                "traffic": (tr["dram_read_bytes"] + tr["dram_write_bytes"]) if tr else None,
                "launch_us": float(np.mean(ts)) * 1e3, "launches_per_step": per_step,
                "gflop_per_launch": float(np.mean(fl)) / 1e9, "share_of_step": per_step * float(np.mean(ts)) / ms_per_step,
                "peak_source": f"{pk['source']}, {tf_kind} bf16 (clocks {clocks['sm_mhz']} / {clocks['sm_max_mhz']} MHz, "
                               f"reasons {clocks['reasons']})"}

    # one launch of ffn_fused_kernel = one or two feed-forward modules (4*M*d*F FLOP each: SURVEY 8d gives 8*N*d*F per layer
    # for its two modules) + optionally the Q/K/V projections (2*M*d*3d) of the layer that follows
    roofline = tensor_roofline(
        "ffn", "ffn_fused",
        f"ffn_fused_kernel (feed-forward module(s) w_1+SiLU+w_2+residual+LayerNorm, chained across the layer boundary, "
        f"+ Q/K/V projections, in one kernel) M={n_tok} d={d} F={F}",
        f"gemm_tc_kernel x2 per feed-forward module (w_1+SiLU, w_2+residual+LayerNorm; the fused kernel does not cover "
        f"d={d}) M={n_tok} d={d} F={F}")
    roofline_kernels = [r for r in (
        tensor_roofline("mhsa", "mhsa_fused",
                        f"mhsa_fused_kernel (scores+mask+softmax+PV for all heads + linear_out + residual + LayerNorm) B={B} T={T} d={d}",
                        f"attention_pp_kernel (two-query-tile ping-pong flash attention) + gemm_tc_kernel (linear_out + residual + LayerNorm) B={B} T={T} d={d}"),
        tensor_roofline("conv", "conv_fused",
                        f"conv_fused_kernel (pointwise_conv1+GLU, depthwise k={kc}+BatchNorm+SiLU, pointwise_conv2+mask+residual+LayerNorm) M={n_tok} d={d}",
                        f"gemm_tc_kernel (pw1+GLU) + dwconv_kernel + gemm_tc_kernel (pw2+residual+LayerNorm) M={n_tok} d={d} k={kc}"),
    ) if r is not None]
    # ---- HBM roofline of the depthwise-conv kernel (north_star: "depthwise-conv kernels at >= 70 % of HBM bandwidth").
    #      In the fused C2 path the GLU / depthwise tensors never reach HBM (conv_fused_kernel); dwconv_kernel is the live
    #      kernel of the unfused path (d != 256: C3; fp32; training forward / backward).  It is timed here stand-alone on
    #      the workload's (B, T, d) bf16 tensor, L2 flushed before every launch, CUDA events around the launch.
    xdw = torch.randn(B, T, d, device=dev).to(torch.bfloat16)
    ydw = torch.empty_like(xdw)
    wdw = torch.randn(kc, d, device=dev) * 0.2
    bdw = torch.zeros(d, device=dev)
    tdw = []
    for _ in range(7):
        flush.fill_(1)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        ops.dwconv(xdw, wdw, bdw, ydw)
        e_.record()
        e_.synchronize()
        tdw.append(s_.elapsed_time(e_))
    t_dw = float(np.median(tdw[2:]))
    dw_bytes = 2.0 * n_tok * d * 2                    # read + write one bf16 (N,d) tensor
    ach_dw = dw_bytes / (t_dw * 1e-3) / 1e9
    in_situ = fam.get("dwconv", 0)
    roofline_hbm = {
        "bound": "hbm", "kernel": f"dwconv_kernel k={kc} + folded BatchNorm + SiLU, (N={n_tok}, d={d}) bf16",
        "achieved": ach_dw, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_dw / pk["hbm_gbs"],
        "traffic": (traffic["dwconv_kernel"]["dram_read_bytes"] + traffic["dwconv_kernel"]["dram_write_bytes"])
        if "dwconv_kernel" in traffic else None,
        "launch_us": t_dw * 1e3, "launches_per_step": in_situ, "peak_source": pk["source"],
        "how": "stand-alone launch on the workload's tensor, cold L2, CUDA events" +
               ("" if in_situ else "; not on this workload's measured path (fused into conv_fused_kernel)")}

    # ---- strong scaling (BASELINE configs[1]: ONE batch sharded over the GPUs) with the final output gather
    strong = None
    if world > 1:
        strong = strong_scaling_leg(args, enc, cfg, dev, rank, world, dtype, flush, barrier)

    # ---- the same-box GPU bar: the reference's algorithm as eager PyTorch (cuBLAS / cuDNN / ATen kernels) on this B200
    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        eager = eager_gpu_baseline(cfg, dev, x_emb, attn, pos, pad, audio_s)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            # bounded sample: whole batches of the same workload, repeated until ~12 s of CPU work have been timed
            v, best, sb, tms = cpu_port_rtfx(args.workload, args.cpu_sample, 3, min_seconds=12.0)
            cpu = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"ATen-CPU port (oracle/conformer_oracle_torch.py), measured path on {sb} of {B} utterances per pass, "
                             f"{len(tms)} passes = {sum(tms):.1f} s of CPU work, best pass {best:.2f} s (mean {sum(tms) / len(tms):.2f} s)"}
        algo_tf = {"C2": 1.024, "C3": 3.552, "C4": 1.914}.get(args.workload)      # SURVEY 8d totals (TFLOP per pass)
        line = {"metric": "encoder audio-sec/sec (RTFx), measured path", "value": value, "unit": "audio-s/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: Conformer-{cfg_name} {cfg['encoder_num_layers']}L d={d} "
                                       f"encoder layers+after_norm, batch {B} x {WORKLOADS[args.workload][2]:.0f} s per GPU (T={T})",
                           "per_gpu_batch": B, "timing": "CUDA events per step on the launch stream; 256 MiB write flushes L2 between steps",
                           "host_cores_per_rank": None if pinned_cores is None else len(pinned_cores)},
                "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": int(feats_host.numel() * 4 + lens_np.nbytes),
                        "d2h_bytes_per_step": int(out_bytes), "ms_per_step": e2e_ms, "steps": e2e_steps,
                        "serial_latency_ms": lat_ms,
                        "output_dtype": str(out_dt).replace("torch.", ""),
                        "api": "EncoderPipeline(encoder, depth=2).stream((feats_pinned_host, lengths_host), pinned_out_bufs): "
                               "ConformerEncoder.forward per batch; H2D / compute / D2H of consecutive batches on three streams; "
                               "wall clock over all steps (per-step working set ~1.4 GB >> L2); serial_latency_ms = one "
                               "synchronous forward call incl. both copies"},
                "gpu_launches": launches, "roofline": roofline, "roofline_kernels": roofline_kernels,
                "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "clocks": clocks,
                "kernels_per_step": {k: v for k, v in fam.items() if v},
                "extra": {"eager_gpu_baseline": eager, "strong_scaling": strong}}
        if algo_tf:
            line["model_tflops"] = algo_tf / (ms_per_step / 1e3) * world
        out = real_stdout if real_stdout is not None else sys.stdout
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.destroy_process_group()


def eager_gpu_baseline(cfg, dev, x_emb, attn, pos, pad, audio_s):
    """Report-only: the measured path of the reference as EAGER PyTorch on the same GPU -- the ATen restatement in
    oracle/conformer_oracle_torch.py issues exactly the torch ops the reference's modules issue (F.linear, F.conv1d,
    softmax, layer_norm, ...; pinned bit-exactly to the reference's goldens on CPU), here dispatched to cuBLAS / cuDNN /
    ATen CUDA kernels.  fp32 with TF32 off (the parity setting) and bf16 autocast (the fast setting)."""
    import contextlib
    import torch
    from oracle import conformer_oracle as O
    from oracle import conformer_oracle_torch as OT
    sd = {k: v.to(dev) for k, v in OT.to_torch_sd(O.make_state_dict(cfg, 0)).items()}
    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for mode in ("fp32_tf32_off", "bf16_autocast"):
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else contextlib.nullcontext()
            with ctx:
                for _ in range(2):
                    OT.encoder_layers(x_emb, attn, pos, pad, sd, cfg)
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s_.record()
                    OT.encoder_layers(x_emb, attn, pos, pad, sd, cfg)
                    e_.record()
                    e_.synchronize()
                    ts.append(s_.elapsed_time(e_))
            ms = float(np.median(ts))
            res[mode] = {"ms_per_step": ms, "value": audio_s / (ms / 1e3), "unit": "audio-s/s"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    res["what"] = ("eager PyTorch (ATen port of the reference modules, oracle/conformer_oracle_torch.py) on the same B200, "
                   "measured path, same batch, median of 5 device-timed passes")
    return res


def strong_scaling_leg(args, enc, cfg, dev, rank, world, dtype, flush, barrier):
    """ONE length-sorted global batch (the N=1 workload, identical on every rank) is stride-sharded over the ranks
    (sharding.shard_indices), each rank encodes its shard from pinned HOST features through the public forward, the
    outputs are all-gathered over NCCL (the north_star's 'final output gather') and rank 0 copies the assembled batch to
    the host -- all inside the timed region.  Wall clock per step, max over ranks."""
    import torch
    import torch.distributed as dist
    from conformer_pytorch_lightning_b200 import sharding
    cfg_name, feats_np, lens_np, T, audio_s = make_inputs(args.workload, seed=1234)
    Bg = feats_np.shape[0]
    idx = sharding.shard_indices(Bg, rank, world)
    per = (Bg + world - 1) // world
    feats_host = torch.from_numpy(feats_np[idx.numpy()]).pin_memory()
    lens_dev = torch.from_numpy(lens_np[idx.numpy()]).to(dev)
    gathered = torch.empty((world * per, T, cfg["encoder_dim"]), dtype=torch.float32, device=dev)
    host_out = torch.empty((world * per, T, cfg["encoder_dim"]), dtype=torch.float32).pin_memory() if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    gather_ms = []

    def one():
        with torch.no_grad():
            out, _ = enc(feats_host.to(dev, non_blocking=True), lens_dev)
            if out.size(0) < per:                       # ragged last shard: pad to the common shard size
                out = torch.cat([out, out.new_zeros((per - out.size(0),) + tuple(out.shape[1:]))])
            ev[0].record()
            dist.all_gather_into_tensor(gathered, out.contiguous())
            ev[1].record()
            if rank == 0:
                host_out.copy_(gathered, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            gather_ms.append(ev[0].elapsed_time(ev[1]))

    for _ in range(max(3, args.warmup)):
        one()
    gather_ms.clear()
    steps = max(8, args.steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    t = torch.tensor([1e3 * (time.perf_counter() - t0) / steps, float(np.mean(gather_ms))], device=dev, dtype=torch.float64)
    barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0].item())
    return {"scaling": "strong", "value": audio_s / (ms / 1e3), "unit": "audio-s/s", "ms_per_step": ms, "steps": steps,
            "global_batch": Bg, "per_gpu_batch": per, "all_gather_ms": float(t[1].item()),
            "all_gather_bytes": int(gathered.numel() * 4),
            "what": "one global batch stride-sharded by utterance; per step: H2D of the shard's fbank, forward, NCCL all-gather "
                    "of the (B/N, T, d) outputs, D2H of the assembled batch on rank 0; wall clock, max over ranks"}


def run_train(args):
    """Workload C5 (BASELINE.json configs[4]): Conformer-M training step -- forward + CTC loss + backward + gradient
    all-reduce + optimizer step -- with a static chunk-16 attention mask, bf16 compute, 16 utterances x 10 s per GPU
    (batch 128 on 8 GPUs; weak scaling), dropout 0.1 as shipped (train.sh).  The layer stack's forward/backward and the
    CTC head run on the native kernels (training.py, ctc.py); the sub-sampling front-end's autograd stays in PyTorch
    (outside the measured path by north_star); gradients of layer i are all-reduced over NCCL while layers i-1.. are
    still in backward (ddp.GradSync).  value = device-timed steps with the batch resident in HBM; e2e = the same step fed
    from pinned host memory with the loss read back every step."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import build_encoder
    from oracle import conformer_oracle as O       # cfg + seeded weights only
    import conformer_pytorch_lightning_b200 as C
    from conformer_pytorch_lightning_b200 import _native, ddp, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pin_rank_cores(local, world)
    real_stdout = None
    if world > 1:
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    B, sec, V, Lmax = 16, 10.0, 5002, 40
    tin, T = frames(sec)
    cfg = O.conformer_cfg("M", static_chunk_size=16, dropout=0.1, attention_dropout=0.1, pos_enc_dropout=0.1)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    enc = build_encoder(cfg, 0, device=dev, compute_dtype=dtype).train()
    dec = C.CTCDecoder(V, cfg["encoder_dim"], 0.0).to(dev)
    dec.compute_dtype = dtype
    ddp.broadcast_parameters(enc)
    ddp.broadcast_parameters(dec)
    sync = ddp.attach(enc) if world > 1 else None
    other = [p for n, p in enc.named_parameters() if n.startswith("embed.")] + list(dec.parameters())
    # module.py:140-143: torch.optim.Adam(self.parameters(), lr).  FlatAdam = the same update on flat buffers that follow the
    # gradient buckets (one launch per layer instead of PyTorch's multi-tensor walk); CFM_BENCH_TORCH_ADAM=1 measures with
    # torch.optim.Adam(fused=True) instead.
    torch_adam = os.environ.get("CFM_BENCH_TORCH_ADAM", "0") == "1"
    all_params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.Adam(all_params, lr=1e-4, fused=True) if torch_adam else C.FlatAdam(all_params, lr=1e-4)
    opt_name = "torch.optim.Adam(fused)" if torch_adam else "FlatAdam (cfm_adam_step)"
    rs = np.random.RandomState(1234 + rank)
    feats_host = torch.from_numpy(rs.standard_normal((B, tin, 80)).astype(np.float32)).pin_memory()
    lens = torch.full((B,), tin, dtype=torch.int32, device=dev)
    labels = torch.from_numpy(rs.randint(1, V - 1, size=(B, Lmax)).astype(np.int64)).to(dev)
    lab_len = torch.full((B,), Lmax, dtype=torch.int64, device=dev)
    feats_dev = feats_host.to(dev)
    audio_s = B * ((tin - 1) * 160 + 400) / 16000.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def step(feats):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
            out, mask = enc(feats, lens)
        loss = dec(out.float(), mask.squeeze(1).sum(1), labels, lab_len)
        loss.backward()
        if world > 1:
            ddp.sync_grads(other)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(feats_dev)
    barrier()
    n0 = _native.launch_count() + engine.GRAPH_REPLAYED_LAUNCHES[0]
    total_ms = 0.0
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.fill_(1)
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            step(feats_dev)
            e_.record()
            e_.synchronize()
            total_ms += s_.elapsed_time(e_)
    barrier()
    launches = _native.launch_count() + engine.GRAPH_REPLAYED_LAUNCHES[0] - n0
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    value = world * audio_s / (ms / 1e3)
    # e2e: batch from pinned host memory, loss read back on the host every step
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(5, args.steps // 2)
    for _ in range(e2e_steps):
        loss = step(feats_host.to(dev, non_blocking=True))
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _ = float(loss_host)
    e2e_t = torch.tensor([1e3 * (time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.item())
    # per-phase device time of one step (events; eager): forward / backward / optimizer
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    opt.zero_grad(set_to_none=True)
    ev[0].record()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
        out, mask = enc(feats_dev, lens)
    loss = dec(out.float(), mask.squeeze(1).sum(1), labels, lab_len)
    ev[1].record()
    loss.backward()
    if world > 1:
        ddp.sync_grads(other)
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    phases = {"forward_ms": ev[0].elapsed_time(ev[1]), "backward_ms": ev[1].elapsed_time(ev[2]),
              "optimizer_ms": ev[2].elapsed_time(ev[3])}
    # same-box GPU bar: the reference's training arithmetic as eager PyTorch autograd on this GPU (ATen port of the
    # reference modules: F.linear / conv1d / layer_norm / batch_norm / softmax / F.ctc_loss, bf16 autocast, dropout 0 --
    # dropout would only add kernels), forward + backward without the optimizer step
    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        from oracle import conformer_oracle_torch as OT
        cfg0 = dict(cfg, dropout=0.0, attention_dropout=0.0, pos_enc_dropout=0.0)
        sd_g = {k: (v.to(dev).requires_grad_() if v.is_floating_point() and "running" not in k else v.to(dev))
                for k, v in OT.to_torch_sd(O.make_state_dict(cfg0, 0)).items()}
        w_g = dec.ctc_lo.weight.detach().clone().requires_grad_()
        b_g = dec.ctc_lo.bias.detach().clone().requires_grad_()
        leaves = [v for v in sd_g.values() if v.requires_grad] + [w_g, b_g]
        pos_g = OT.rel_pos_table(cfg0["max_len"], cfg0["encoder_dim"])[:B].to(dev)
        Tin = feats_dev.size(1)
        pad_g = (torch.arange(Tin, device=dev).unsqueeze(0) < lens.unsqueeze(1).long()).unsqueeze(1)
        chunk_g = OT.chunk_mask(T, 16, -1).to(dev).unsqueeze(0)

        def eager_step():
            for v in leaves:
                v.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                x, pad2 = OT.subsampling(feats_dev, pad_g, sd_g)
                out_e = OT.encoder_layers_train(x, pad2 & chunk_g, pos_g, pad2, sd_g, cfg0)
            loss_e = OT.ctc_loss(out_e.float(), pad2.squeeze(1).sum(1), labels, lab_len, w_g, b_g)
            loss_e.backward()
        for _ in range(2):
            eager_step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            eager_step()
            e_.record()
            e_.synchronize()
            ts.append(s_.elapsed_time(e_))
        eager = {"bf16_autocast_fwd_bwd_ms": float(np.median(ts)), "ours_fwd_bwd_ms": phases["forward_ms"] + phases["backward_ms"],
                 "what": "eager PyTorch autograd (ATen port of the reference modules + F.ctc_loss) on the same B200, bf16 "
                         "autocast, dropout 0, forward + backward of one 16-utterance batch, no optimizer step"}
    fam = {f: _native.kernel_launches(f) for f in ("gemm_tc", "gemm_gen", "gemm_gen_simt", "gemm_simt", "ln_bwd", "softmax_bwd",
                                                   "ctc_grad", "dwconv_wgrad")}
    pk = peaks()
    clocks = clk.summary()
    burst_ok = ("sw_power_cap" not in clocks["reasons"] and clocks["sm_mhz"] is not None and clocks["sm_max_mhz"]
                and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    tf_peak = pk["tf_burst"] if burst_ok else pk["tf_sustained"]
    algo_tf = 6.142 / 8.0                           # C5 fwd+bwd of the layer stack per 16-utterance shard
    ctc_tf = 3 * 2.0 * B * T * cfg["encoder_dim"] * V / 1e12
    if rank == 0:
        line = {"metric": "training audio-sec/sec (fwd + CTC loss + bwd + grad all-reduce + optimizer step)", "value": value,
                "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"C5: Conformer-M training, CTC loss (V={V}, {Lmax} labels/utt), static chunk-16 mask, "
                                       f"dropout 0.1, batch {B} x {sec:.0f} s per GPU (T={T}), Adam",
                           "per_gpu_batch": B, "global_batch": B * world,
                           "timing": "CUDA events per step on the launch stream; 256 MiB write flushes L2 between steps"},
                "e2e": {"value": world * audio_s / (e2e_ms / 1e3), "unit": "audio-s/s",
                        "h2d_bytes_per_step": int(feats_host.numel() * 4), "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
                        "steps": e2e_steps,
                        "api": "ConformerEncoder.forward (train mode, autograd) + CTCDecoder.forward + loss.backward() + "
                               "ddp.GradSync per-layer all-reduce + " + opt_name + ".step(); features from pinned host "
                               "memory and the loss read back on the host every step"},
                "gpu_launches": launches, "phases": phases, "kernels_total": fam,
                "roofline": {"bound": "tensor", "kernel": "whole step (layer stack fwd+bwd + CTC head GEMMs), algorithmic FLOP of "
                             "SURVEY 8d (3 x forward) over the device-timed step",
                             "achieved": (algo_tf + ctc_tf) / (ms / 1e3), "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": (algo_tf + ctc_tf) / (ms / 1e3) / tf_peak, "traffic": None},
                "grad_allreduce": None if sync is None else {"buckets_per_step": sync.buckets_sent // max(1, (args.steps + max(3, args.warmup) + e2e_steps + 1)),
                                                            "bytes_per_step": sync.bytes_sent // max(1, (args.steps + max(3, args.warmup) + e2e_steps + 1))},
                "cpu_baseline": None, "clocks": clocks, "extra": {"eager_gpu_baseline": eager}}
        out_f = real_stdout if real_stdout is not None else sys.stdout
        out_f.write(json.dumps(line) + "\n")
        out_f.flush()
    if world > 1:
        dist.destroy_process_group()


def enc_make_pad(lens, max_len):
    from conformer_pytorch_lightning_b200.utils import make_pad_mask
    return make_pad_mask(lens, max_len).unsqueeze(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(WORKLOADS) + ["C5"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=64, help="utterances per CPU-baseline pass (capped at the batch)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-GPU bar")
    ap.add_argument("--profile", action="store_true", help="measured path only (for ncu): no e2e / probe / cpu legs")
    args = ap.parse_args()
    if args.workload == "C5":
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm covers the inference workloads C2-C4"}))
            return
        run_train(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
