#!/usr/bin/env python
"""Headline benchmark: strain-seconds searched per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (N=1): BASELINE.json configs[1] -- "whisper-base encoder log-mel binary classifier, batch
1024 windows, 1xB200": the Signal_vs_Noise two-detector classifier (whisper-base geometry, random-init
weights, seeded DoRA adapters on q/k/v merged at load) over 1024 sliding windows x 2 detectors of
synthetic Gaussian strain per step.  One window advances the search by 204 samples = 0.0996 s of
strain (MLGWSC-1/inference.py:198-199), so  value = windows * (204/2048) / seconds.

One JSON line on stdout (rank 0).  Under torchrun each rank searches its own time shard (weak
scaling); the only collective is the all-gather of the small per-rank trigger lists.

The same line carries two sub-records (skip with --no-extra):
  "mlgwsc"        BASELINE.json configs[3]/[4]: the MLGWSC-1 search (QScan + Q-Adapter + whisper-tiny) over a
                  one-hour, four-segment ragged stream, partitioned over the N ranks by sharding.plan_shards and
                  merged by the trigger all-gather (strong scaling of a fixed job), with its own roofline /
                  cpu_baseline and an in-bench check against a single-rank prefix run;
  "glitch_small"  configs[2] (N=1 only): whisper-small, 512 glitch-shaped windows, argmax.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HOP = 204
FS = 2048
SEC_PER_WINDOW = HOP / FS


def flops_per_detwin(d, L, ffn, T=1500, pruned=False):
    """FLOPs (2*MAC) per det-window by GEMM class.  pruned=False: the ALGORITHMIC work of the full
    reference computation (tiny 36.938 G, base 87.368 G, small 344.162 G).
    pruned=True: the work EXECUTED when the final layer is evaluated for the last token only
     -- used for per-kernel and roofline rates so they are not inflated."""
    conv1 = 2 * 3000 * (80 * 3) * d
    conv2 = 2 * T * (3 * d) * d
    qkv = 2 * T * d * 3 * d
    attn = 4 * T * T * d
    oproj = 2 * T * d * d
    fc1 = 2 * T * d * ffn
    fc2 = 2 * T * ffn * d
    Lf = (L - 1 + 1.0 / T) if pruned else L        # final layer: one token's row instead of T
    per = {"gemm_conv1": conv1, "gemm_conv2": conv2, "gemm_qkv": L * qkv, "attention": Lf * attn,
           "gemm_out_proj": Lf * oproj, "gemm_fc1": Lf * fc1, "gemm_fc2": Lf * fc2}
    per["total"] = sum(per.values())
    return per


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(size: str, chunk: int):
    """Random-init encoder of the named Whisper geometry + seeded DoRA + seeded 2-detector head.
    (gw_whisper_b200.synthetic: the same seeded weights the reference arm builds)."""
    from gw_whisper_b200 import B200WhisperEncoder, two_channel_ligo_binary_classifier
    from gw_whisper_b200 import synthetic as S

    base = S.make_encoder(size, 0, spread=True)
    dora = S.synthetic_dora(size, targets=("q_proj", "k_proj", "v_proj"))
    enc = B200WhisperEncoder.from_hf(base, dora=dora, chunk=chunk)
    model = two_channel_ligo_binary_classifier(enc, num_classes=1)
    S.seeded_head(model.classifier, seed=3, gain=3.0)
    model.refresh()
    return model, base


def build_mlgwsc_model(size: str, batch: int):
    """BASELINE.json configs[3]: MLGWSC-1 model = Q-transform + Q-Adapter + whisper-tiny encoder + DoRA
    + 5-layer softmax head, two detectors, 256-window batches (MLGWSC-1/inference.py:354-392,465)."""
    from gw_whisper_b200 import B200WhisperEncoder, GWWhisperClassifier, QTransformAdapter
    from gw_whisper_b200 import synthetic as S
    import torch

    base = S.make_encoder(size, 0, spread=True)
    dora = S.synthetic_dora(size, targets=("q_proj", "k_proj", "v_proj", "out_proj"))
    enc = B200WhisperEncoder.from_hf(base, dora=dora, chunk=2 * batch)
    torch.manual_seed(11)
    adapter = QTransformAdapter(n_detectors=2)
    model = GWWhisperClassifier(enc, 2, num_classes=2, q_adapter=adapter)
    S.seeded_head(model.classifier, seed=3, gain=3.0)
    model.refresh()
    return model, base


# MLGWSC-1 stream of the `mlgwsc` sub-record: one hour of two-detector strain in four ragged segments
# (BASELINE.json configs[3]; configs[4] is the same search over a month).  Lengths in samples (even, not
# multiples of the 204-sample hop or of the 256-window batch): 1500.25 s, 1100.5 s, 700.125 s, 299.125 s.
MLGWSC_SEGMENTS = (3072512, 2253824, 1433856, 612608)


def mlgwsc_segments(dev, scale: float = 1.0):
    """Synthetic whitened strain generated ON THE DEVICE, identical on every rank (seed 1234 + segment id), with a
    loud sine-Gaussian every ~97 s so that trigger counts differ from shard to shard."""
    import math
    import torch
    segs = []
    for i, n in enumerate(MLGWSC_SEGMENTS):
        n = int(n * scale) & ~1
        g = torch.Generator(device=dev).manual_seed(1234 + i)
        x = torch.randn(2, n, generator=g, device=dev)
        t = torch.arange(4096, device=dev) / 2048.0
        for j, t0 in enumerate(range(20 * 2048, n - 8192, 97 * 2048 + 333)):
            f0, tau, amp = 60.0 + 37.0 * (j % 9), 0.01 + 0.004 * (j % 5), 6.0 + 3.0 * (j % 4)
            sg = amp * torch.exp(-(t - 1.0) ** 2 / (2 * tau ** 2)) * torch.sin(2 * math.pi * f0 * t)
            x[0, t0:t0 + 4096] += sg
            x[1, t0 + 12:t0 + 12 + 4096] += 0.8 * sg
        segs.append(x)
    return segs


def mlgwsc_cpu_baseline(size: str, n_windows: int, threads: int):
    """The reference's MLGWSC-1 forward on the host cores: restated ml4gw QScan + the reference's Q-Adapter CNN
    structure + HF WhisperEncoder fp32 with unmerged DoRA + softmax-less head (oracle/)."""
    import torch
    from oracle import encoder as E, qscan as OQ
    torch.set_num_threads(threads)
    base = E.make_encoder(size, 0, spread=True)
    enc = E.attach_dora(base, E.synthetic_dora(size, targets=("q_proj", "k_proj", "v_proj", "out_proj")))
    torch.manual_seed(11)
    adapter = OQ.QTransformAdapter(n_detectors=2).eval()
    head = E.seeded_head(E.head_mlgwsc(base.config.d_model, 2, 2, softmax=False), seed=3, gain=3.0)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(n_windows, 2, 2048, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        feats = adapter(x)
        reps = torch.cat([enc(feats[:, i]).last_hidden_state[:, -1, :] for i in range(2)], dim=1)
        out = head(reps)
    dt = time.perf_counter() - t0
    return n_windows / dt, dt, float(out.mean())


def mlgwsc_record(args, dev, rank, world, lib):
    """`mlgwsc` sub-record of the JSON line: the MLGWSC-1 search (QScan + Q-Adapter +
    whisper-tiny + DoRA + USR head) over the one-hour ragged stream, partitioned over the ranks by
    `sharding.plan_shards` (whole 256-window batches, 1844-sample halo) and merged with the trigger all-gather
    `sharding.gather_triggers` (NCCL): strong scaling of a fixed job.  Timed on the device, max over ranks."""
    import torch
    import torch.distributed as dist
    from gw_whisper_b200 import _lib, sharding
    from gw_whisper_b200 import remove_softmax_from_classifier

    model, base = build_mlgwsc_model("tiny", 256)
    remove_softmax_from_classifier(model)
    segs = mlgwsc_segments(dev, args.mlgwsc_scale)
    nws = [sharding.n_windows(int(s.shape[1]), HOP) for s in segs]
    total_windows = sum(nws)
    # threshold: 97th percentile of the scores of a 512-window prefix of segment 0 (same on every rank)
    n_prefix = min(512, nws[0])
    pre_scores, _, _ = model.stream_search(segs[0][:, :(n_prefix - 1) * HOP + 2048].contiguous(), HOP, n_prefix, 1e30)
    thr = float(torch.quantile(pre_scores, 0.97))
    _, pre_idx, pre_sc = model.stream_search(segs[0][:, :(n_prefix - 1) * HOP + 2048].contiguous(), HOP, n_prefix, thr)
    group = None

    def one_pass():
        return sharding.sharded_search(model, segs, HOP, thr, rank, world, group, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one_pass()                                             # warm-up (workspaces, tensor maps, NCCL buffers)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.gww_launch_count()
    e0.record()
    (t_seg, t_idx, t_sc), all_scores = one_pass()
    e1.record()
    torch.cuda.synchronize()
    launches = lib.gww_launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # correctness inside the bench: the merged list restricted to the prefix == the single-rank prefix run
    sel = (t_seg == 0) & (t_idx < n_prefix)
    ok = bool(torch.equal(t_idx[sel], pre_idx.to(torch.int64)) and torch.equal(t_sc[sel], pre_sc))
    ok = ok and all(int(a.numel()) == n for a, n in zip(all_scores, nws))
    if not ok:
        raise RuntimeError("mlgwsc: merged sharded triggers differ from the single-rank prefix run")
    plan = sharding.plan_shards(nws, world)
    trig_per_rank = []
    key = t_seg * (1 << 40) + t_idx
    for r in range(world):
        c = 0
        for p in plan[r]:
            lo, hi = p.segment * (1 << 40) + p.first_window, p.segment * (1 << 40) + p.first_window + p.n_windows
            c += int(((key >= lo) & (key < hi)).sum())
        trig_per_rank.append(c)
    rec = None
    del all_scores
    # per-kernel-class timing of rank 0's shard (one more pass)
    nk = lib.gww_profile_num_kinds()
    ms_k = (C.c_double * nk)()
    cnt_k = (C.c_long * nk)()
    lib.gww_profile_begin()
    one_pass()
    torch.cuda.synchronize()
    _lib.check(lib.gww_profile_end(ms_k, cnt_k))
    if world > 1:
        dist.barrier()
    if rank == 0:
        names = [lib.gww_profile_kind_name(i).decode() for i in range(nk)]
        kernels = {nm: {"ms": ms_k[i], "launches": int(cnt_k[i])} for i, nm in enumerate(names) if cnt_k[i]}
        cfg = base.config
        fl = flops_per_detwin(cfg.d_model, cfg.encoder_layers, cfg.encoder_ffn_dim)
        hbm_peak = 6650.0
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            hbm_peak = float(json.load(open(pk_path)).get("hbm_gbs", hbm_peak))
        n_dw0 = 2 * sum(p.n_windows for p in plan[0])      # det-windows rank 0 processed in the profiled pass
        # front-end kernels of rank 0's shard against the roofline that bounds each (algorithmic work per det-window,
        # QScan 8 KB in + 1 MB spectrogram out; conv1 1 MB in + 4 MB of bf16 hi/lo planes out; conv2 /
        # conv3 604 MFLOP each -- counted once, although the split-precision path executes three partial products;
        # pool 64 KB in + 480 KB of time-major features out)
        peak_tf = 1590.0
        if os.path.exists(pk_path):
            pk = json.load(open(pk_path))
            peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", peak_tf)))
        spec_fe = {"qscan": ("hbm", 2048 * 4 + 512 * 512 * 4), "qadapter_conv1": ("hbm", 512 * 512 * 4 + 256 * 256 * 64),
                   "qadapter_conv2": ("tensor", 2.0 * 256 * 256 * 9 * 16 * 32), "qadapter_conv3": ("tensor", 2.0 * 128 * 128 * 9 * 32 * 64),
                   "qadapter_pool": ("hbm", 128 * 128 * 4 + 3002 * 80 * 2), "qadapter": ("hbm", 512 * 512 * 4 + 3002 * 80 * 2)}
        fe = {}
        ncu_tr = {}
        tr_path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
        if os.path.exists(tr_path):
            ncu_tr = json.load(open(tr_path))

        def fe_traffic(nm):
            """dram read+write bytes per launch (256 det-windows) from the committed ncu --set full capture"""
            keys = ["qscan_tiles", "qscan_interp"] if nm == "qscan" else [nm]
            if not all(k in ncu_tr for k in keys):
                return None
            return sum(ncu_tr[k]["dram_bytes_per_launch"] for k in keys)

        for nm, (bound, work) in spec_fe.items():
            if nm not in kernels:
                continue
            rate = work * n_dw0 / (kernels[nm]["ms"] * 1e-3)
            if bound == "hbm":
                fe[nm] = {"bound": "hbm", "achieved": rate / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": rate / 1e9 / hbm_peak,
                          "algorithmic_bytes_per_det_window": work, "ms": kernels[nm]["ms"], "traffic": fe_traffic(nm),
                          "traffic_source": ncu_tr.get("source")}
            else:
                fe[nm] = {"bound": "tensor", "achieved": rate / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                          "frac": rate / 1e12 / peak_tf, "algorithmic_flops_per_det_window": work, "ms": kernels[nm]["ms"],
                          "traffic": fe_traffic(nm), "traffic_source": ncu_tr.get("source"),
                          "note": "bf16 hi/lo split precision: two tcgen05.mma per (tap, 16-channel step) -- A_hi x [W_hi | W_lo] "
                                  "and A_lo x W_hi -- for one algorithmic product"}
        dom = max(fe, key=lambda k: fe[k]["ms"]) if fe else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            wps, dt, _ = mlgwsc_cpu_baseline("tiny", args.mlgwsc_ref_windows, threads)
            cpu = {"value": wps * SEC_PER_WINDOW, "unit": "strain-s/s", "cores": threads, "kind": "port",
                   "sample": f"{args.mlgwsc_ref_windows} windows x 2 detectors in {dt:.1f} s: restated ml4gw QScan + Q-Adapter "
                             "CNN + HF WhisperEncoder(tiny) fp32 + unmerged DoRA + head"}
        rec = {"metric": "strain-seconds searched/sec", "value": total_windows * SEC_PER_WINDOW / (ms * 1e-3),
               "unit": "strain-s/s", "n_gpus": world, "ms": ms, "scaling": "strong",
               "config": {"workload": "MLGWSC-1 search: QScan + Q-Adapter + whisper-tiny + DoRA(q,k,v,out) + USR head over "
                                      f"{sum(int(s.shape[1]) for s in segs) / FS:.1f} s of 2-detector strain in "
                                      f"{len(segs)} ragged segments, 256-window batches",
                          "windows": total_windows, "windows_per_segment": nws,
                          "parallelism": f"sharding.plan_shards: whole batches, 1844-sample halo, dp{world}",
                          "collective": "all-gather of per-rank trigger counts + padded (segment, window, score) lists, "
                                        "all-reduce-as-gather of the per-window scores (sharding.gather_triggers / gather_scores)"},
               "windows_per_s": total_windows / (ms * 1e-3), "threshold": thr,
               "triggers": int(t_idx.numel()), "triggers_per_rank": trig_per_rank,
               "prefix_check": "merged triggers == single-rank run on the first %d windows" % n_prefix,
               "model_tflops": (fl["total"] + 1.286e9) * 2 * total_windows / (ms * 1e-3) / 1e12,
               "gpu_launches_rank0": int(launches), "kernels_rank0": kernels,
               "roofline": (dict(fe[dom], kernel=dom) if dom else None), "roofline_frontend": fe, "cpu_baseline": cpu}
    return rec


def glitch_record(args, dev, lib):
    """`glitch_small` sub-record: BASELINE.json configs[2], whisper-small + 11-class head on 512 glitch-shaped
    windows (one detector) per step; strain resident in HBM."""
    import math
    import torch
    from gw_whisper_b200 import B200WhisperEncoder, glitch_one_channel_classifier
    from gw_whisper_b200 import synthetic as S
    B = 512
    base = S.make_encoder("small", 0, spread=True)
    enc = B200WhisperEncoder.from_hf(base, chunk=args.chunk)
    model = glitch_one_channel_classifier(enc, num_classes=11)
    S.seeded_head(model.classifier, seed=5, gain=3.0)
    model.refresh()
    g = torch.Generator().manual_seed(4321)
    t = torch.arange(2048) / 2048.0
    strain = torch.randn(B, 2048, generator=g)
    A = 5 + 15 * torch.rand(B, 1, generator=g)
    f0 = 30 + 470 * torch.rand(B, 1, generator=g)
    tau = 0.002 + 0.048 * torch.rand(B, 1, generator=g)
    t0 = 0.3 + 0.4 * torch.rand(B, 1, generator=g)
    strain += A * torch.exp(-(t[None] - t0) ** 2 / (2 * tau ** 2)) * torch.sin(2 * math.pi * f0 * t[None])
    strain = strain[:, None, :].to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    steps = max(1, min(args.steps, 2))

    def step():
        flush.zero_()
        return model.forward_strain(strain).argmax(1)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cfg = base.config
    fl = flops_per_detwin(cfg.d_model, cfg.encoder_layers, cfg.encoder_ffn_dim)
    return {"metric": "strain-seconds searched/sec", "value": B * SEC_PER_WINDOW / (ms * 1e-3), "unit": "strain-s/s",
            "ms_per_step": ms, "steps": steps, "windows_per_s": B / (ms * 1e-3),
            "model_tflops": fl["total"] * B / (ms * 1e-3) / 1e12,
            "config": {"workload": "Glitch_classification: whisper-small + 11-class head, log-mel, 512 glitch-shaped "
                                   "windows x 1 detector per step (argmax)", "det_windows_per_chunk": args.chunk}}


def reference_windows_per_s(size: str, n_windows: int, threads: int):
    """The reference's CPU path on `n_windows` two-detector windows: scipy.signal.resample ->
    WhisperFeatureExtractor (as installed) -> HF WhisperEncoder fp32 with unmerged DoRA -> head."""
    import numpy as np
    import torch
    from oracle import encoder as E, logmel as L

    torch.set_num_threads(threads)
    base = E.make_encoder(size, 0, spread=True)
    dora = E.synthetic_dora(size, targets=("q_proj", "k_proj", "v_proj"))
    model = E.TwoChannelOracle(E.attach_dora(base, dora), 1).eval()
    E.seeded_head(model.classifier, seed=3, gain=3.0)
    g = torch.Generator().manual_seed(1234)
    strain = torch.randn(n_windows, 2, 2048, generator=g).numpy()
    t0 = time.perf_counter()
    feats = torch.from_numpy(L.logmel_reference(strain, path="torch"))
    with torch.no_grad():
        out = model(feats[:, 0], feats[:, 1])
    dt = time.perf_counter() - t0
    return n_windows / dt, dt, float(out.mean())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = args.ref_windows
    for _ in range(max(args.warmup, 0) and 1):
        reference_windows_per_s(args.model, 2, threads)
    times = []
    for _ in range(args.steps):
        wps, dt, _ = reference_windows_per_s(args.model, per_step, threads)
        times.append(dt)
    tot = sum(times)
    value = args.steps * per_step * SEC_PER_WINDOW / tot
    line = {
        "impl": "reference", "metric": "strain-seconds searched/sec", "value": value,
        "unit": "strain-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Signal_vs_Noise two-detector classifier, whisper-{args.model} encoder + DoRA(q,k,v) "
                               f"+ 4-layer head, log-mel front end, 1 s windows @2048 Hz, hop 204",
                   "windows_per_step": per_step, "detectors": 2},
        "cpu_baseline": {"value": value, "unit": "strain-s/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} windows x 2 detectors per step (bounded sample of the 1024-window batch): "
                                   "scipy resample + HF WhisperFeatureExtractor + HF WhisperEncoder fp32 + unmerged DoRA + head"},
        "e2e": {"value": value, "unit": "strain-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from gw_whisper_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, D = args.batch, 2
    model, base = build_model(args.model, args.chunk)
    cfg = base.config
    d, L, ffn = cfg.d_model, cfg.encoder_layers, cfg.encoder_ffn_dim

    # synthetic whitened strain: i.i.d. N(0,1), one shard per rank (seed 1234 + rank)
    g = torch.Generator().manual_seed(1234 + rank)
    host_strain = torch.randn(B, D, 2048, generator=g).pin_memory()
    dev_strain = host_strain.to(dev)
    host_out = torch.empty(B, 1).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    thr = 0.0
    trig_idx = torch.zeros(B, dtype=torch.long, device=dev)
    trig_sc = torch.zeros(B, dtype=torch.float32, device=dev)
    trig_cnt = torch.zeros(1, dtype=torch.int32, device=dev)

    def step_device():
        flush.zero_()
        out = model.forward_strain(dev_strain)
        trig_cnt.zero_()
        _lib.check(lib.gww_threshold_compact(out.data_ptr(), 1, B, thr, rank * B, trig_idx.data_ptr(),
                                             trig_sc.data_ptr(), trig_cnt.data_ptr(), B, _lib.stream_ptr()))
        if world > 1:   # all-gather of the (padded) per-rank trigger lists: the path's only collective
            cnts = [torch.empty_like(trig_cnt) for _ in range(world)]
            dist.all_gather(cnts, trig_cnt)
            scs = [torch.empty_like(trig_sc) for _ in range(world)]
            dist.all_gather(scs, trig_sc)
        return out

    def step_e2e():
        flush.zero_()
        s = host_strain.to(dev, non_blocking=True)
        out = model.forward_strain(s)
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.gww_launch_count()
        h0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        timed.host_issue_ms = (time.perf_counter() - h0) * 1e3 / steps
        torch.cuda.synchronize()
        timed.launches = lib.gww_launch_count() - l0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, warm)
    launches_per_run = timed.launches
    host_issue_ms = timed.host_issue_ms
    ms_e2e = timed(step_e2e, args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None

    # live per-kernel-class timing (CUDA events on the launching stream) over K more steps
    nk = lib.gww_profile_num_kinds()
    ms_k = (C.c_double * nk)()
    cnt_k = (C.c_long * nk)()
    lib.gww_profile_begin()
    for _ in range(args.steps):
        step_device()
    torch.cuda.synchronize()
    _lib.check(lib.gww_profile_end(ms_k, cnt_k))
    names = [lib.gww_profile_kind_name(i).decode() for i in range(nk)]
    pruned = bool(lib.gww_set_last_layer_pruning(1))      # query (and keep the default: on)
    lib.gww_set_last_layer_pruning(int(pruned))
    fl_alg = flops_per_detwin(d, L, ffn)
    fl = flops_per_detwin(d, L, ffn, pruned=pruned)       # executed
    n_dw = B * D
    kernels = {}
    for i, nm in enumerate(names):
        if cnt_k[i] == 0:
            continue
        ent = {"ms_per_step": ms_k[i] / args.steps, "launches_per_step": cnt_k[i] // args.steps}
        if nm in fl:
            ent["tflops"] = fl[nm] * n_dw * args.steps / (ms_k[i] * 1e-3) / 1e12
        kernels[nm] = ent
    gemm_names = [n for n in names if n.startswith("gemm_")]
    gemm_ms = sum(ms_k[names.index(n)] for n in gemm_names)
    gemm_launches = sum(cnt_k[names.index(n)] for n in gemm_names)
    gemm_flops = sum(fl[n] for n in gemm_names) * n_dw * args.steps
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_src = "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst)"
    peak_tf = 1590.0
    hbm_peak = 6650.0
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", peak_tf)))
        hbm_peak = float(peaks.get("hbm_gbs", hbm_peak))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    ncu_traffic = {}
    tr_path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(tr_path):
        tr_path = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tr_path):
        ncu_traffic = json.load(open(tr_path))

    def traffic_of(key):
        """dram read+write bytes per launch from the committed ncu --set full capture (same model,
        256 det-windows per launch); None when the bench runs another geometry."""
        ent = ncu_traffic.get(key)
        if not ent or args.model != "base":
            return None
        # the capture was taken at ent["det_windows"] per launch; every kernel here streams its operands, so
        # DRAM bytes scale with the det-windows of a launch (average over the step's chunks, the last is ragged)
        n_chunks = -(-n_dw // args.chunk)
        return ent["dram_bytes_per_launch"] * (n_dw / n_chunks) / ent["det_windows"]

    traffic_src = (f"{ncu_traffic.get('source', 'profiles/' + os.path.basename(tr_path))}; dram read+write bytes per launch "
                   "of that capture scaled to this run's det-windows per launch (not measured in this run)") if ncu_traffic else None
    roofline_gemm = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, all encoder GEMMs)", "achieved": achieved,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": {k: traffic_of(k) for k in ("gemm_qkv", "gemm_out_proj", "gemm_fc1", "gemm_fc2")},
                     "traffic_source": traffic_src,
                     "peak_source": peak_src, "flops_per_launch": gemm_flops / max(gemm_launches, 1),
                     "avg_launch_ms": gemm_ms / max(gemm_launches, 1)}
    # the single kernel with the largest share of the step is the fused attention
    ia = names.index("attention")
    att_ms, att_n = ms_k[ia], cnt_k[ia]
    att_flops = fl["attention"] * n_dw * args.steps
    att_tf = att_flops / (att_ms * 1e-3) / 1e12 if att_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "attention_persist_kernel (tcgen05 QK^T / PV, softmax on MUFU)",
                "achieved": att_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": att_tf / peak_tf,
                "traffic": traffic_of("attention_persist_kernel"), "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": args.chunk * 1500 * (3 * d + d) * 2,
                "peak_source": peak_src, "flops_per_launch": att_flops / max(att_n, 1),
                "avg_launch_ms": att_ms / max(att_n, 1), "share_of_step": att_ms / max(sum(ms_k), 1e-9),
                "note": "head_dim 64: one exp per 128 MAC, so the softmax (MUFU.EX2 16/clk/SM) needs >= 2x the tensor time "
                        "of a tile: the tensor-pipe ceiling of this kernel is ~50% of peak (profiles/r2_ubench_softmax_mix2.txt: "
                        "12.6 exp/clk/SM is the instruction-mix ceiling at two softmax warps per sub-partition, 15.1 with 2 of 8 "
                        "pairs on the FMA-pipe polynomial, which the kernel uses; r2 ncu: XU 64 %, tensor 42 % active, "
                        "profiles/r2_ncu_summary.csv)"}
    if "logmel" in kernels:
        lm_bytes = (2048 * 4 + 3002 * 80 * 2) * n_dw   # fused path writes bf16 time-major features
        kernels["logmel"]["gbs"] = lm_bytes / (kernels["logmel"]["ms_per_step"] * 1e-3) / 1e9
        kernels["logmel"]["hbm_frac"] = kernels["logmel"]["gbs"] / hbm_peak
    total_flops = fl_alg["total"] * n_dw
    value = world * B * SEC_PER_WINDOW * args.steps / (ms_dev * 1e-3)
    value_full = None
    if pruned:   # transparency: the same step with the full 1500-token final layer
        lib.gww_set_last_layer_pruning(0)
        ms_full = timed(step_device, max(1, min(args.steps, 2)), 1)
        lib.gww_set_last_layer_pruning(1)
        value_full = world * B * SEC_PER_WINDOW * max(1, min(args.steps, 2)) / (ms_full * 1e-3)
    e2e_value = world * B * SEC_PER_WINDOW * args.steps / (ms_e2e * 1e-3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        wps, dt, _ = reference_windows_per_s(args.model, args.ref_windows, threads)
        cpu_baseline = {"value": wps * SEC_PER_WINDOW, "unit": "strain-s/s", "cores": threads, "kind": "port",
                        "sample": f"{args.ref_windows} windows x 2 detectors in {dt:.1f} s: scipy resample + HF "
                                  "WhisperFeatureExtractor + HF WhisperEncoder fp32 + unmerged DoRA + head"}
    extras = {}
    if not args.no_extra:
        del model, dev_strain, flush
        torch.cuda.empty_cache()
        extras["mlgwsc"] = mlgwsc_record(args, dev, rank, world, lib)
        if rank == 0 and world == 1:
            torch.cuda.empty_cache()
            extras["glitch_small"] = glitch_record(args, dev, lib)
    operand = lib.gww_operand_dtype().decode()
    if rank == 0:
        line = {
            "metric": "strain-seconds searched/sec", "value": value, "unit": "strain-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": operand, "data": "synthetic",
            "config": {"workload": f"Signal_vs_Noise two-detector classifier, whisper-{args.model} encoder + DoRA(q,k,v) "
                                   "+ 4-layer head, log-mel front end, 1 s windows @2048 Hz, hop 204",
                       "operands": f"{operand} tensor-core operands (tcgen05 kind::f16), fp32 accumulation / residual stream / "
                                   "softmax, f64 log-mel front end",
                       "windows_per_step_per_gpu": B, "detectors": D, "det_windows_per_chunk": args.chunk,
                       "parallelism": f"time-shard dp{world}",
                       "final_layer": ("last token only: all tokens' K/V, one query row (exact; the reference consumes "
                                       "last_hidden_state[:, -1, :] only)" if pruned else "full 1500 tokens"),
                       "l2": "256 MiB buffer written between timed steps (L2 flush); activations per chunk >> 126 MB L2"},
            "windows_per_s": world * B * args.steps / (ms_dev * 1e-3),
            "model_tflops": total_flops * world * args.steps / (ms_dev * 1e-3) / 1e12,
            "executed_tflops": fl["total"] * n_dw * world * args.steps / (ms_dev * 1e-3) / 1e12,
            "value_full_final_layer": value_full,
            "e2e": {"value": e2e_value, "unit": "strain-s/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host_strain.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": int(launches_per_run), "host_issue_ms_per_step": host_issue_ms,
            "roofline": roofline, "roofline_gemm": roofline_gemm, "kernels": kernels, "cpu_baseline": cpu_baseline, "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_single(args):
    """One sub-record alone (ncu / tuning runs): same code path as inside the full line."""
    import torch
    import torch.distributed as dist
    from gw_whisper_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    rec = mlgwsc_record(args, dev, rank, world, lib) if args.workload == "mlgwsc" else glitch_record(args, dev, lib)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="base", choices=["tiny", "base", "small"])
    ap.add_argument("--batch", type=int, default=1024, help="windows per step per GPU")
    ap.add_argument("--chunk", type=int, default=296,
                    help="det-windows per encoder pass (default 2 x 148 SMs: whole waves of the one-CTA-per-det-window "
                         "front end and 96 attention work items per SM; measured 217.8 vs 222.5 ms per step for 256)")
    ap.add_argument("--ref-windows", type=int, default=40, help="windows per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="all", choices=["all", "mlgwsc", "glitch"],
                    help="all = headline line with the sub-records; mlgwsc / glitch = that sub-record alone as the line "
                         "(profiling runs)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the mlgwsc (configs[3], sharded over the ranks) and glitch_small (configs[2]) sub-records")
    ap.add_argument("--mlgwsc-scale", type=float, default=1.0, help="scale of the one-hour MLGWSC-1 stream (tests: 0.02)")
    ap.add_argument("--mlgwsc-ref-windows", type=int, default=96, help="windows of the MLGWSC-1 CPU baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "all":
        run_single(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
