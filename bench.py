#!/usr/bin/env python
"""bench.py -- audio-seconds generated per wall-second (RTF^-1) of the HiFi-GAN
generator hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode tf32|fp16|bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: reference path on host cores

Headline line (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] -- a step is one
generator forward over 16 utterances x 172 frames (2 s each at 22.05 kHz, hop 256) in tf32 mode: 10-bit-mantissa
operands, fp32 accumulate, residual stream of >= 22 mantissa bits.  For the default configuration the library
holds every MMA operand as fp16 (tf32's mantissa, rounded to nearest) and the residual stream as an fp16 pair
hi + lo ("split plan", hfg_tf32_plan), so the MMAs are tcgen05 kind::f16; every rank
runs its own batch (utterance sharding, no data-path collective): "scaling": "weak".

  value  device-timed (CUDA events around each step, L2 flushed between steps, mel already resident in
         HBM), whole-job: sum of audio-seconds over ranks / max over ranks of the timed duration.
  e2e    the same metric through the public call a user with host data makes -- HiFiGANGenerator.generate_stream:
         page-locked host mel -> H2D -> kernels -> D2H into a page-locked host wav for every step, two
         submissions in flight so the copies overlap the neighbours' kernels (`e2e_synchronous`: one blocking
         HiFiGANGenerator(mel_cpu) call per step).
  roofline      dominant kernel group, from per-launch CUDA events inside the library (hfg_set_profiling),
                against the BURST peak (isolated sub-millisecond launches) with the sustained figure beside it.
  cpu_baseline  the oracle's ATen restatement of the reference (oracle/torch_port.py; the same conv kernels
                the reference dispatches to on CPU) timed on this box's host cores on the FULL batch.
  quality       every tensor-core mode against the ORACLE output of the same batch (max-abs, log-mel L1).

Sub-records in the same JSON line (BASELINE.json configs[2], [3], [0]):
  config3   256 x 172 frames, bf16 (and fp16), utterances strong-sharded shard_bounds(256, N, rank):
            device-timed and end to end through host buffers, final waveform gather timed separately.
  config4   one 5168-frame (60 s) mel, time-sharded over the N ranks with a receptive-field halo
            (sharding.generate_time_sharded), asserted bit-equal to the unchunked run on rank 0.
  config1   1 x 256 frames: single-utterance latency on the GPU (device path and host-buffer graph replay).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 22050
WORKLOAD = dict(batch=16, frames=172)          # BASELINE.json configs[1]
CONFIG3 = dict(batch=256, frames=172)          # BASELINE.json configs[2]
CONFIG4 = dict(batch=1, frames=5168)           # BASELINE.json configs[3]
CONFIG1 = dict(batch=1, frames=256)            # BASELINE.json configs[0]
METRIC = "audio-sec generated per sec (RTF^-1)"
UNIT = "audio-s/s"
DTYPE = {"fp32": "f32", "tf32": "tf32", "bf16": "bf16", "fp16": "f16"}


def audio_seconds(batch, frames, hop=256):
    return batch * frames * hop / SAMPLE_RATE


def config_dict(world):
    """Identical in both arms (GPU and --impl reference), so the driver compares like with like."""
    B, T = WORKLOAD["batch"], WORKLOAD["frames"]
    return {"workload": f"HiFi-GAN generator (default YAML config, random-init), batch {B} x {T} frames "
                        f"({audio_seconds(1, T):.2f} s utterances) per step and per GPU",
            "parallelism": f"utterance-sharded x{world}, no data-path collective",
            "l2": "256 MiB flush between timed steps (GPU arm)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the bench is under load."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, windows):
        """windows: [(name, t_begin, t_end)] of device-timed regions.  Samples inside any of them are
        summarised; the first window (the headline's timed region) is also summarised alone.  A region shorter
        than the sampling period may hold no sample: then everything sampled under load is used and said so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def summarise(picked):
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for l in picked:
                f = [x.strip() for x in l.split(",")]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            sm.sort()
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "samples": len(sm), "reasons": sorted(reasons)}

        inside = [l for (t, l) in self.lines if any(a <= t <= b for (_, a, b) in windows)]
        window = "timed regions: " + ", ".join(n for n, _, _ in windows)
        if not inside:
            window = "bench under load (warm-up .. end of the timed regions); timed regions shorter than the sampling period"
            inside = [l for (_, l) in self.lines]
        out = summarise(inside)
        out["window"] = window
        if windows:
            n0, a0, b0 = windows[0]
            head = [l for (t, l) in self.lines if a0 <= t <= b0]
            if head:
                out["headline_region"] = summarise(head)
        return out


def cpu_reference_run(batch, frames, repeats, threads=None, keep_output=False):
    """Time the reference path on host cores (oracle ATen restatement)."""
    import torch
    import oracle
    from tts_sambert_hifigan_b200 import synth
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.DEFAULT_CONFIG
    sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()}
    mel = torch.from_numpy(synth.make_mel(1, batch, cfg["n_mels"], frames))
    times, out = [], None
    with torch.no_grad():
        oracle.forward_torch(cfg, sd, mel[:1, :, : min(frames, 32)])       # warm the thread pool
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = oracle.forward_torch(cfg, sd, mel)
            times.append(time.perf_counter() - t0)
    return times, cores, (out if keep_output else None)


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the
    box's host cores.  Rank 0 only; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    B, frames = WORKLOAD["batch"], WORKLOAD["frames"]
    # the full batch per step, unless this box is so slow that K + W steps would not finish in a few minutes
    probe, cores, _ = cpu_reference_run(B, frames, 1)
    sample_batch = B
    while sample_batch > 1 and probe[0] * sample_batch / B * (args.warmup + args.steps) > 240.0:
        sample_batch //= 2
    times, cores, _ = cpu_reference_run(sample_batch, frames, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    val = audio_seconds(sample_batch, frames) * len(timed) / total
    sample = (f"the full batch: {sample_batch} x {frames} frames per step" if sample_batch == B else
              f"{sample_batch} of {B} utterances per step (the full batch would take {probe[0]:.1f} s per step here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample}, {len(timed)} steps, oracle/torch_port.py (same ATen conv "
                                   "kernels as the reference; pinned to the live reference by tests/golden)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("HFG_BENCH_MODE", "tf32"), choices=["fp32", "tf32", "bf16", "fp16"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch"])
    ap.add_argument("--frames", type=int, default=WORKLOAD["frames"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-quality", action="store_true", help="skip the other-mode / output-quality passes")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 3 / 4 / 1 sub-records")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import tts_sambert_hifigan_b200 as pkg
    from tts_sambert_hifigan_b200 import sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    cfg = synth.DEFAULT_CONFIG
    B, T = args.batch, args.frames
    steps = max(1, args.steps)
    warmup = max(3, args.warmup)
    sd_np = synth.make_weights(cfg, 0)
    # ONE module: the library packs every precision at commit time and the arithmetic mode is a per-call
    # argument, so switching `gen.mode` costs nothing
    gen = pkg.HiFiGANGenerator(**cfg, mode=args.mode).to(dev)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()})
    # every rank gets its own utterances (seed by rank): utterance sharding.  Rank 0's batch is the one the
    # CPU arm / oracle runs (seed 1).
    mel_host = torch.from_numpy(synth.make_mel(1 + rank, B, cfg["n_mels"], T)).pin_memory()   # page-locked input
    mel = mel_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def device_timed(fn, n, warm):
        """n steps of fn(), each bracketed by CUDA events on the current stream, L2 flushed in between (untimed).
        Returns (sorted per-step ms, wall-clock window)."""
        for _ in range(warm):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        barrier(); torch.cuda.synchronize()
        t_begin = time.time()
        for s in range(n):
            flush.fill_(s & 0xFF)
            ev[s][0].record()
            fn()
            ev[s][1].record()
        torch.cuda.synchronize(); barrier()
        return sorted(a.elapsed_time(b) for a, b in ev), (t_begin, time.time())

    def host_timed(fn, n, warm=2):
        for _ in range(warm):
            fn()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = time.perf_counter() - t0
        barrier()
        return dt

    def stream_timed(mel_cpu, n, warm=4):
        """n batches from page-locked host memory through gen.generate_stream (two submissions in flight: every
        step still copies its own mel H2D and its own waveform D2H, the copies ride under the neighbours' kernels)."""
        import itertools
        for _ in gen.generate_stream(itertools.repeat(mel_cpu, warm)):
            pass
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for w in gen.generate_stream(itertools.repeat(mel_cpu, n)):
            keep["wav_host"] = w
        dt = time.perf_counter() - t0
        barrier()
        return dt

    sampler = ClockSampler(local_rank)
    sampler.start()
    windows = []
    sub = {}
    with torch.no_grad():
        # ---------------- headline: device-timed region ----------------
        keep = {}

        def step():
            keep["wav"] = gen(mel)
        step_ms, win = device_timed(step, steps, warmup)
        windows.append(("headline", *win))
        launches_per_step = gen.last_launch_count
        wav = keep["wav"]
        dev_ms = sum(step_ms)

        # ---------------- config 3 / config 4 device-timed regions (sampler still running) ----------------
        c3, c4 = {}, {}
        if not args.no_configs:
            a3, b3 = sharding.shard_bounds(CONFIG3["batch"], world, rank)
            mel3_all = synth.make_mel(21, CONFIG3["batch"], cfg["n_mels"], CONFIG3["frames"])
            mel3_host = torch.from_numpy(np.ascontiguousarray(mel3_all[a3:b3])).pin_memory()
            mel3 = mel3_host.to(dev)
            n3 = max(5, min(steps, 20))
            for m in ("bf16", "fp16"):
                gen.mode = m
                ms3, win3 = device_timed(lambda: keep.__setitem__("c3", gen(mel3)), n3, 3)
                windows.append((f"config3 {m}", *win3))
                c3[m] = dict(ms=ms3, wav_head=keep["c3"][:4].clone() if rank == 0 else None, steps=n3)
            mel4_np = synth.make_mel(9, 1, cfg["n_mels"], CONFIG4["frames"])
            mel4 = torch.from_numpy(mel4_np).to(dev)
            n4 = max(5, min(steps, 20))
            for m in (args.mode, "bf16"):
                gen.mode = m
                ms4, win4 = device_timed(
                    lambda: keep.__setitem__("c4", sharding.generate_time_sharded(gen, mel4, hop=256, gather=False)
                                             if world > 1 else gen(mel4)), n4, 3)
                windows.append((f"config4 {m}", *win4))
                c4[m] = dict(ms=ms4, local=keep["c4"], steps=n4)
            gen.mode = args.mode
        clocks = sampler.stop(windows)

        # ---------------- end-to-end regions (host buffers) ----------------
        # nvidia-smi polling takes driver locks that stall the synchronous host calls of these loops by more
        # than a millisecond per step, so the sampler covers the device-timed regions and is stopped here.
        e2e_steps = args.e2e_steps or steps
        e2e_sync_s = host_timed(lambda: keep.__setitem__("wav_host", gen(mel_host)), e2e_steps)
        e2e_s = stream_timed(mel_host, e2e_steps)
        wav_host = keep["wav_host"]
        if c3:
            for m in c3:
                gen.mode = m
                c3[m]["e2e_sync_s"] = host_timed(lambda: gen(mel3_host), c3[m]["steps"])
                c3[m]["e2e_s"] = stream_timed(mel3_host, c3[m]["steps"], warm=3)
            # optional final gather of the waveforms (the only collective of the path), timed separately
            gen.mode = "bf16"
            if world > 1:
                local = gen(mel3)
                sizes = [sharding.shard_bounds(CONFIG3["batch"], world, r)[1] - sharding.shard_bounds(CONFIG3["batch"], world, r)[0]
                         for r in range(world)]
                for _ in range(2):
                    sharding._gather_var(local, sizes, 0)
                torch.cuda.synchronize(); barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(5):
                    sharding._gather_var(local, sizes, 0)
                g1.record()
                torch.cuda.synchronize()
                c3["gather_ms"] = g0.elapsed_time(g1) / 5
                del local
            gen.mode = args.mode

        # ---------------- config 4 exactness: chunks equal the unchunked run ----------------
        c4_check = {}
        if c4:
            for m in c4:
                gen.mode = m
                local = c4[m]["local"]
                if world > 1:
                    sizes = [(sharding.shard_bounds(CONFIG4["frames"], world, r)[1] -
                              sharding.shard_bounds(CONFIG4["frames"], world, r)[0]) * 256 for r in range(world)]
                    whole = sharding._gather_var(local, sizes, 2)
                else:                                     # one GPU: 8 chunks in this process, as config 4 describes
                    whole = sharding.generate_chunked(gen, mel4, 8, hop=256)
                full = gen(mel4)                          # unchunked, every rank (cheap) -- compared on rank 0
                torch.cuda.synchronize()
                c4_check[m] = {"max_abs_vs_unchunked": float((whole - full).abs().max()),
                               "bit_equal": bool(torch.equal(whole, full))}
                c4[m]["full"] = full if rank == 0 else None
            gen.mode = args.mode

        # ---------------- config 1: single-utterance latency ----------------
        c1 = {}
        if not args.no_configs and rank == 0:
            mel1_host = torch.from_numpy(synth.make_mel(1, 1, cfg["n_mels"], CONFIG1["frames"])).pin_memory()
            mel1 = mel1_host.to(dev)
            for m in ("tf32", "fp16", "bf16"):
                gen.mode = m
                for _ in range(5):
                    gen(mel1)
                lat = []
                for _ in range(30):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(); gen(mel1); e1.record()
                    torch.cuda.synchronize()
                    lat.append(e0.elapsed_time(e1))
                for _ in range(4):
                    gen(mel1_host)
                host = []
                for _ in range(30):
                    t0 = time.perf_counter(); gen(mel1_host); host.append(1e3 * (time.perf_counter() - t0))
                lat.sort(); host.sort()
                c1[m] = {"device_ms_median": lat[len(lat) // 2], "device_ms_best": lat[0],
                         "host_buffers_ms_median": host[len(host) // 2], "host_buffers_ms_best": host[0]}
            gen.mode = args.mode
        barrier()

        # ---------------- other arithmetic modes (untimed for `value`) ----------------
        outs, other = {args.mode: wav}, {}
        if rank == 0 and not args.no_quality:
            for m in ("fp32", "tf32", "fp16", "bf16"):
                if m == args.mode:
                    continue
                gen.mode = m
                if m == "fp32":
                    outs[m] = gen(mel)
                    continue
                n2 = max(3, steps // 2)
                for _ in range(3):
                    keep["o"] = gen(mel)
                ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n2)]
                for s2 in range(n2):                      # rank-0-only path: no barriers in here
                    flush.fill_(s2 & 0xFF)
                    ev2[s2][0].record()
                    keep["o"] = gen(mel)
                    ev2[s2][1].record()
                torch.cuda.synchronize()
                ms2 = [a.elapsed_time(b) for a, b in ev2]
                outs[m] = keep["o"]
                avg = sum(ms2) / len(ms2)
                other[m] = {"ms_per_step": avg, "value_per_gpu": audio_seconds(B, T) / (avg / 1e3),
                            "tflops_per_gpu": synth.flops_per_frame(cfg) * B * T / (avg / 1e3) / 1e12}
            gen.mode = args.mode
            torch.cuda.synchronize()
        barrier()

        # ---------------- per-kernel profile (separate, untimed pass) ----------------
        h = gen._handle_for(dev)
        tf32_split = h.tf32_plan_is_split()     # tf32 mode on fp16 operand planes + fp16 hi/lo residual stream
        prof_by_mode = {}
        for m in ([args.mode] if args.no_quality else [args.mode] + [x for x in ("tf32", "fp16", "bf16") if x != args.mode]):
            if m != args.mode and rank != 0:
                continue
            gen.mode = m
            h.set_profiling(True)
            for _ in range(3):
                flush.fill_(1)
                gen(mel)
                torch.cuda.synchronize()
                prof_by_mode[m] = h.get_profile()
            h.set_profiling(False)
        gen.mode = args.mode
        ws_bytes = {m: h.workspace_bytes(B, T, pkg._capi.MODES[m]) for m in ("fp32", "tf32", "fp16", "bf16")}

        # ---------------- measured tensor peaks for the roofline (cuBLAS as the yardstick, after the timed regions) ----
        tf32_peak = None
        if rank == 0:
            try:
                n = 8192
                a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
                old = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = True
                for _ in range(3):
                    a @ b
                best = 1e9
                for _ in range(10):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                torch.backends.cuda.matmul.allow_tf32 = old
                tf32_peak = 2.0 * n ** 3 / (best / 1e3) / 1e12
                del a, b
            except Exception:
                tf32_peak = None

    dev_ms_max, e2e_s_max, e2e_sync_s_max = max_over_ranks(dev_ms, e2e_s, e2e_sync_s)
    # config 3 / 4: max over ranks of the summed step times
    if c3:
        for m in ("bf16", "fp16"):
            c3[m]["ms_max"], c3[m]["e2e_max"], c3[m]["e2e_sync_max"] = max_over_ranks(sum(c3[m]["ms"]), c3[m]["e2e_s"], c3[m]["e2e_sync_s"])
    if c4:
        for m in c4:
            c4[m]["ms_max"], = max_over_ranks(sum(c4[m]["ms"]))

    if rank == 0:
        peaks = load_peaks()
        audio_step_all = audio_seconds(B, T) * world
        value = audio_step_all * steps / (dev_ms_max / 1e3)
        e2e_val = audio_step_all * e2e_steps / e2e_s_max
        flops_step = synth.flops_per_frame(cfg) * B * T

        # ---------------- CPU arm + oracle outputs (rank 0 host cores) ----------------
        from tts_sambert_hifigan_b200 import metrics
        cpu_baseline, quality = None, {}
        ref = None
        if not args.no_cpu_baseline:
            times, cores, ref = cpu_reference_run(B, T, 3, keep_output=True)      # rank 0's own batch (mel seed 1)
            best = min(times)
            cpu_baseline = {
                "value": audio_seconds(B, T) / best, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"the full batch ({B} x {T} frames), best of 3, oracle/torch_port.py (the ATen conv "
                          "kernels the reference dispatches to; pinned to the live reference by tests/golden)"}
        if ref is not None and not args.no_quality:
            ref_dev = ref.to(dev)
            for m, o in outs.items():
                q = {"max_abs_vs_oracle": float((o - ref_dev).abs().max()), "ref_peak": float(ref_dev.abs().max())}
                try:
                    q["log_mel_l1_vs_oracle"] = metrics.log_mel_l1(ref_dev, o)
                except Exception as e:  # torchaudio missing on the box
                    q["log_mel_l1_vs_oracle"] = None
                    q["log_mel_note"] = f"unavailable: {type(e).__name__}"
                quality[m] = q

        # ---------------- roofline of the dominant kernel group ----------------
        def roofline_of(mode):
            prof = prof_by_mode[mode]
            step_ms_prof = sum(p["ms"] for p in prof)
            mrf = [p for p in prof if p["kernel"].startswith("mrf")]
            burst = peaks["bf16_tflops"]
            sustained = peaks["bf16_tflops_sustained"]
            note = "bf16/fp16 dense, burst (a kernel timed alone); MEASURED_PEAKS.json"
            if mode == "tf32" and tf32_split:
                note += ("; tf32 mode on the split plan: every MMA of a fused pair is tcgen05 kind::f16 on fp16 operands, "
                         "so the bf16/fp16 peak applies")
            if mode == "tf32" and not tf32_split:
                if tf32_peak:
                    burst, sustained = tf32_peak, tf32_peak * peaks["bf16_tflops_sustained"] / peaks["bf16_tflops"]
                    note = ("tf32 dense burst measured in this run (torch.matmul 8192^3, allow_tf32, best of 10); "
                            "sustained scaled like the bf16 pair of MEASURED_PEAKS.json")
                else:
                    burst, sustained = burst / 2, sustained / 2
                    note = "tf32 dense taken as measured bf16 / 2 (no tf32 peak could be measured)"
            if mode == "fp32":
                dom = max(prof, key=lambda p: p["ms"])
                ffma = 148 * 128 * 2 * 1.965e9 / 1e12
                ach = dom["flops"] / (dom["ms"] / 1e3) / 1e12
                return {"bound": "tensor", "kernel": f"conv_tile_fp32 {dom['kernel']} (fp32 FFMA kernels: no tensor cores in this mode)",
                        "achieved": ach, "peak": ffma, "unit": "TFLOP/s", "frac": ach / ffma, "traffic": None,
                        "peak_source": "nominal fp32 FFMA rate 148 SM x 128 lanes x 2 x 1.965 GHz (not a tensor peak)"}
            dom = max(mrf, key=lambda p: p["ms"])
            if mode == "tf32" and not tf32_split and not dom["kernel"].endswith(":conv"):
                # a fused pair in tf32 mode runs conv1 as kind::tf32 and conv2 on the fp16 copy of the on-chip
                # intermediate (kind::f16): half of its FLOPs at each rate -> the peak of the launch is the
                # harmonic mean of the two measured peaks
                f16_b, f16_s = peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]
                burst, sustained = 2.0 / (1.0 / burst + 1.0 / f16_b), 2.0 / (1.0 / sustained + 1.0 / f16_s)
                note += ("; fused pair = conv1 at the tf32 rate + conv2 at the fp16 rate (fp16 intermediate), equal "
                         "FLOPs each: peak = harmonic mean of the tf32 and the bf16/fp16 burst peaks")
            achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12
            mrf_ms = sum(p["ms"] for p in mrf)
            mrf_tflops = sum(p["flops"] for p in mrf) / (mrf_ms / 1e3) / 1e12
            traffic, traffic_note = None, "no ncu capture committed for this kernel"
            for tname in ("r2_traffic.json", "r1_traffic.json"):
                tpath = os.path.join(ROOT, "profiles", tname)
                if os.path.exists(tpath):
                    with open(tpath) as f:
                        ent = json.load(f).get(f"{mode}:{dom['kernel']}")
                    if ent:
                        traffic, traffic_note = ent["dram_bytes_per_launch"], ent["source"]
                        break
            unfused = dom["kernel"].endswith(":conv")
            return {
                "bound": "tensor",
                "kernel": "%s %s (%d launches per step)" % ("tc_conv_kernel" if unfused else "tc_pair_kernel",
                                                            dom["kernel"], dom["launches"]),
                "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                "frac": achieved / burst, "frac_of_sustained_peak": achieved / sustained, "peak_sustained": sustained,
                "traffic": traffic, "traffic_note": traffic_note,
                "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"],
                "algorithmic_flops_per_launch": dom["flops"] / dom["launches"],
                "ms_per_launch": dom["ms"] / dom["launches"],
                "peak_source": peaks["source"] + "; " + note,
                "share_of_step": dom["ms"] / step_ms_prof if step_ms_prof else None,
                "serial_step_ms": step_ms_prof,
                "all_mrf_launches": {"launches": sum(p["launches"] for p in mrf), "tflops": mrf_tflops,
                                     "frac": mrf_tflops / burst, "share_of_step": mrf_ms / step_ms_prof},
                "per_stage": [{"kernel": p["kernel"], "launches": p["launches"], "ms": round(p["ms"], 4),
                               "tflops": round(p["flops"] / (p["ms"] / 1e3) / 1e12, 2) if p["ms"] > 0 else None,
                               "gbs": round(p["bytes"] / (p["ms"] / 1e3) / 1e9, 1) if p["ms"] > 0 else None}
                              for p in prof],
            }

        roofline = roofline_of(args.mode)
        roofline_other = {m: roofline_of(m) for m in prof_by_mode if m != args.mode}

        # ---------------- sub-records ----------------
        if c3:
            audio3 = audio_seconds(CONFIG3["batch"], CONFIG3["frames"])
            rec3 = {"workload": f"{CONFIG3['batch']} x {CONFIG3['frames']} frames, utterances strong-sharded "
                                f"shard_bounds({CONFIG3['batch']}, {world}, rank): {b3 - a3} on rank 0",
                    "scaling": "strong", "n_gpus": world, "audio_seconds_per_step": audio3}
            oracle_head = None
            if not args.no_cpu_baseline and not args.no_quality:
                import oracle
                sd_t = {k: torch.from_numpy(v) for k, v in sd_np.items()}
                with torch.no_grad():
                    oracle_head = oracle.forward_torch(cfg, sd_t, torch.from_numpy(mel3_all[:4])).to(dev)
            for m in ("bf16", "fp16"):
                r = c3[m]
                n3 = r["steps"]
                rec = {"dtype": DTYPE[m], "steps": n3, "ms_per_step": r["ms_max"] / n3,
                       "ms_per_step_best_rank0": r["ms"][0],
                       "value": audio3 * n3 / (r["ms_max"] / 1e3), "unit": UNIT,
                       "tflops_total": synth.flops_per_frame(cfg) * CONFIG3["batch"] * CONFIG3["frames"] * n3 / (r["ms_max"] / 1e3) / 1e12,
                       "e2e_synchronous": {"value": audio3 * n3 / r["e2e_sync_max"], "unit": UNIT},
                       "e2e": {"value": audio3 * n3 / r["e2e_max"], "unit": UNIT,
                               "h2d_bytes_per_step": CONFIG3["batch"] * cfg["n_mels"] * CONFIG3["frames"] * 4,
                               "d2h_bytes_per_step": CONFIG3["batch"] * CONFIG3["frames"] * 256 * 4}}
                if oracle_head is not None:
                    o = r["wav_head"]
                    rec["parity_vs_oracle"] = {"sample": "first 4 utterances of the batch, oracle on host cores",
                                               "max_abs": float((o - oracle_head).abs().max()),
                                               "ref_peak": float(oracle_head.abs().max())}
                    try:
                        rec["parity_vs_oracle"]["log_mel_l1"] = metrics.log_mel_l1(oracle_head, o)
                    except Exception:
                        rec["parity_vs_oracle"]["log_mel_l1"] = None
                rec3[m] = rec
            if "gather_ms" in c3:
                rec3["final_gather_ms"] = c3["gather_ms"]
                rec3["final_gather_note"] = "all_gather of the bf16-mode waveforms (NCCL), timed separately; not in value / e2e"
            sub["config3"] = rec3
        if c4:
            audio4 = audio_seconds(1, CONFIG4["frames"])
            rec4 = {"workload": f"1 x {CONFIG4['frames']} frames (60 s), "
                                + (f"time-sharded over {world} ranks" if world > 1 else "unchunked on one GPU (timed); 8 chunks in-process for the exactness check")
                                + f", halo {gen.receptive_radius + 1} frames (receptive radius {gen.receptive_radius})",
                    "scaling": "strong", "n_gpus": world, "audio_seconds_per_step": audio4}
            oracle4 = None
            if not args.no_cpu_baseline and not args.no_quality:
                import oracle
                sd_t = {k: torch.from_numpy(v) for k, v in sd_np.items()}
                with torch.no_grad():
                    oracle4 = oracle.forward_torch(cfg, sd_t, torch.from_numpy(mel4_np)).to(dev)
            for m in c4:
                r = c4[m]
                n4 = r["steps"]
                rec = {"dtype": DTYPE[m], "steps": n4, "ms_per_step": r["ms_max"] / n4,
                       "value": audio4 * n4 / (r["ms_max"] / 1e3), "unit": UNIT,
                       "chunks_equal_unchunked": c4_check[m]}
                if oracle4 is not None:
                    rec["parity_vs_oracle"] = {"max_abs": float((r["full"] - oracle4).abs().max()),
                                               "ref_peak": float(oracle4.abs().max())}
                    try:
                        rec["parity_vs_oracle"]["log_mel_l1"] = metrics.log_mel_l1(oracle4, r["full"])
                    except Exception:
                        rec["parity_vs_oracle"]["log_mel_l1"] = None
                rec4[m] = rec
            sub["config4"] = rec4
        if c1:
            sub["config1"] = {"workload": "1 x 256 frames (2.97 s): single-utterance latency",
                              "cpu_reference_note": "0.47 s on 8 host cores (SURVEY.md section 6)", **c1}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": dev_ms_max / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("f16 operands (tf32's 10-bit mantissa, round-to-nearest), f32 accumulate, residual stream as an "
                      "f16 pair hi+lo (22-bit mantissa)") if (args.mode == "tf32" and tf32_split) else DTYPE[args.mode],
            "data": "synthetic",
            "config": config_dict(world),
            "mode": args.mode,
            "tf32_plan": ("split: MMAs read fp16 planes (tcgen05 kind::f16), residual stream stored as fp16 hi + lo"
                          if tf32_split else "fp32 planes, tcgen05 kind::tf32"),
            "timing_notes": {
                "streams": "timed steps: the 3 resblocks of each MRF on 3 streams (fork/join events); the "
                           "per-kernel roofline pass serialises them so every launch is timed alone",
                "e2e_timer": "host perf_counter around the whole loop of steps"},
            "ms_per_step_median": step_ms[len(step_ms) // 2], "ms_per_step_best": step_ms[0],
            "tflops_per_gpu": flops_step * steps / (dev_ms_max / 1e3) / 1e12,
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": int(mel_host.numel() * 4),
                    "d2h_bytes_per_step": int(wav_host.numel() * 4),
                    "steps": e2e_steps,
                    "api": "HiFiGANGenerator.generate_stream(host mels) -> host waveforms: hfg_forward_host_submit / _wait, "
                           "two submissions in flight; every step copies its mel H2D from page-locked memory and its "
                           "waveform D2H into page-locked memory"},
            "e2e_synchronous": {"value": audio_step_all * e2e_steps / e2e_sync_s_max, "unit": UNIT,
                                "api": "HiFiGANGenerator(mel_cpu): hfg_forward_host_ex, one blocking call per step"},
            "gpu_launches": int(launches_per_step * steps),
            "quality": quality,
            "other_modes": other,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_other_modes": roofline_other,
            "tf32_peak_measured_tflops": tf32_peak,
            "workspace_bytes": ws_bytes,
            "workspace_bytes_per_mel_frame": {m: v / (B * T) for m, v in ws_bytes.items()},
            **sub,
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
