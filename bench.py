#!/usr/bin/env python
"""bench.py -- audio-seconds generated per wall-second (RTF^-1) of the HiFi-GAN
generator hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode tf32|bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: reference path on host cores

A step is one generator forward over one batch of synthetic mels
(workload = BASELINE.json configs[1]: 16 utterances x 172 frames = 2 s each at
22.05 kHz, hop 256).  Every rank runs its own batch (utterance sharding, no
data-path collective): "scaling": "weak".

  value  device-timed (CUDA events around each step, L2 flushed between steps,
         mel already resident in HBM), whole-job: sum of audio-seconds over ranks
         / max over ranks of the timed duration.
  e2e    the same metric through the public call a user makes --
         HiFiGANGenerator(mel_cpu) -> hfg_forward_host_ex: page-locked host mel ->
         H2D -> kernels -> D2H into a page-locked host wav, every step.
  roofline      dominant kernel class (the MRF convolutions), from per-launch
                CUDA events inside the library (hfg_set_profiling).
  cpu_baseline  the oracle's ATen restatement of the reference (oracle/torch_port.py;
                the same conv kernels the reference dispatches to on CPU) timed
                on this box's host cores on a bounded sample.
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
METRIC = "audio-sec generated per sec (RTF^-1)"
UNIT = "audio-s/s"


def audio_seconds(batch, frames, hop=256):
    return batch * frames * hop / SAMPLE_RATE


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
    """nvidia-smi clocks / throttle reasons during the timed region."""
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
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples taken inside [t_begin, t_end] (the timed region).  A region shorter
        than nvidia-smi's sampling period may hold none: then every sample taken while the bench was
        under load (warm-up .. end of the timed region) is used and the window is named in the result."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        window = "timed region"
        picked = [l for (t, l) in self.lines if t_begin is None or (t_begin <= t <= t_end)]
        if not picked:
            window = "bench under load (warm-up .. timed region); timed region shorter than the sampling period"
            picked = [l for (_, l) in self.lines]
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
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


def cpu_reference_run(batch, frames, repeats, threads=None):
    """Time the reference path on host cores (oracle ATen restatement)."""
    import torch
    import oracle
    from tts_sambert_hifigan_b200 import synth
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.DEFAULT_CONFIG
    sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()}
    mel = torch.from_numpy(synth.make_mel(1, batch, cfg["n_mels"], frames))
    times = []
    with torch.no_grad():
        oracle.forward_torch(cfg, sd, mel[:1, :, : min(frames, 32)])       # warm the thread pool
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.forward_torch(cfg, sd, mel)
            times.append(time.perf_counter() - t0)
    return times, cores


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the
    box's host cores.  Rank 0 only; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sample_batch = 4                                   # bounded sample of the 16-utterance batch
    frames = WORKLOAD["frames"]
    times, cores = cpu_reference_run(sample_batch, frames, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    val = audio_seconds(sample_batch, frames) * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"HiFi-GAN generator, batch {WORKLOAD['batch']} x {frames} frames (2 s utterances), "
                               "reference path on host CPU",
                   "sample": f"{sample_batch} of {WORKLOAD['batch']} utterances per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_batch} x {frames} frames per step, {len(timed)} steps, "
                                   "oracle/torch_port.py (same ATen conv kernels as the reference)"},
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
    ap.add_argument("--mode", default=os.environ.get("HFG_BENCH_MODE", "tf32"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch"])
    ap.add_argument("--frames", type=int, default=WORKLOAD["frames"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-quality", action="store_true", help="skip the other-mode / output-quality passes")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import tts_sambert_hifigan_b200 as pkg
    from tts_sambert_hifigan_b200 import synth

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

    cfg = synth.DEFAULT_CONFIG
    B, T = args.batch, args.frames
    steps = max(1, args.steps)
    warmup = max(3, args.warmup)
    gen = pkg.HiFiGANGenerator(**cfg, mode=args.mode).to(dev)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()})
    # every rank gets its own utterances (seed by rank): utterance sharding
    mel_host = torch.from_numpy(synth.make_mel(1 + rank, B, cfg["n_mels"], T)).pin_memory()   # page-locked input
    mel = mel_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    sampler = ClockSampler(local_rank)
    sampler.start()
    with torch.no_grad():
        for _ in range(warmup):
            wav = gen(mel)
        torch.cuda.synchronize()
        launches_per_step = gen.last_launch_count

        # ---------------- device-timed region ----------------
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier(); torch.cuda.synchronize()
        t_begin = time.time()
        for s in range(steps):
            flush.fill_(s & 0xFF)                    # evict L2 between timed iterations (untimed)
            ev[s][0].record()
            wav = gen(mel)
            ev[s][1].record()
        torch.cuda.synchronize(); barrier()
        t_end = time.time()
        step_ms = sorted(a.elapsed_time(b) for a, b in ev)
        dev_ms = sum(step_ms)

        # ---------------- end-to-end region (host buffers) ----------------
        # nvidia-smi polling takes driver locks that stall the synchronous host calls of this loop by
        # more than a millisecond per step (tools/e2e_breakdown.py vs this loop with the sampler on),
        # so the sampler covers warm-up + the device-timed region and is stopped here.
        clocks = sampler.stop(t_begin, t_end)
        e2e_steps = args.e2e_steps or steps
        for _ in range(2):
            gen(mel_host)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wav_host = gen(mel_host)                 # synchronous: returns a host tensor
        e2e_s = time.perf_counter() - t0
        barrier()

        # ---------------- other arithmetic modes + output quality (untimed for `value`) ----------------
        # the same batch in the strict fp32 mode is the on-device stand-in for the reference output
        # (fp32 mode matches the reference to 3e-8: tests/test_parity_gpu.py, profiles/)
        quality, other = {}, {}
        if rank == 0 and not args.no_quality:
            from tts_sambert_hifigan_b200 import metrics
            sd = gen.state_dict()
            outs = {args.mode: wav}
            for m in ("fp32", "tf32", "bf16"):
                if m == args.mode:
                    continue
                g2 = pkg.HiFiGANGenerator(**cfg, mode=m).to(dev)
                g2.load_state_dict(sd)
                for _ in range(3):
                    outs[m] = g2(mel)
                if m != "fp32":
                    n2 = max(3, steps // 2)
                    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n2)]
                    for s2 in range(n2):
                        flush.fill_(s2 & 0xFF)
                        ev2[s2][0].record()
                        outs[m] = g2(mel)
                        ev2[s2][1].record()
                    torch.cuda.synchronize()
                    ms2 = sum(a.elapsed_time(b) for a, b in ev2) / n2
                    other[m] = {"ms_per_step": ms2, "value_per_gpu": audio_seconds(B, T) / (ms2 / 1e3),
                                "tflops_per_gpu": synth.flops_per_frame(cfg) * B * T / (ms2 / 1e3) / 1e12}
                del g2
            torch.cuda.synchronize()
            ref32 = outs["fp32"]
            for m in ("tf32", "bf16"):
                q = {"max_abs_vs_fp32_mode": float((outs[m] - ref32).abs().max()),
                     "ref_peak": float(ref32.abs().max())}
                try:
                    q["log_mel_l1_vs_fp32_mode"] = metrics.log_mel_l1(ref32, outs[m])
                except Exception as e:  # torchaudio missing on the box
                    q["log_mel_l1_vs_fp32_mode"] = None
                    q["log_mel_note"] = f"unavailable: {type(e).__name__}"
                quality[m] = q
        barrier()

        # ---------------- per-kernel profile (separate, untimed pass) ----------------
        h = gen._handle_for(dev)
        h.set_profiling(True)
        prof_runs = []
        for _ in range(3):
            flush.fill_(1)
            gen(mel)
            torch.cuda.synchronize()
            prof_runs.append(h.get_profile())
        h.set_profiling(False)

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])

    if rank == 0:
        peaks = load_peaks()
        audio_step_all = audio_seconds(B, T) * world
        value = audio_step_all * steps / (dev_ms_max / 1e3)
        e2e_val = audio_step_all * e2e_steps / e2e_s_max
        flops_step = synth.flops_per_frame(cfg) * B * T
        # dominant kernel = the (stage, kernel-size) group of fused ResBlock launches with the largest
        # share of the step (3 launches: dilations 1, 3, 5).  All MRF launches together are also reported.
        prof = prof_runs[-1]
        mrf = [p for p in prof if p["kernel"].startswith("mrf")]
        step_ms_prof = sum(p["ms"] for p in prof)
        dom = max(mrf, key=lambda p: p["ms"])
        achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12
        mrf_ms = sum(p["ms"] for p in mrf)
        mrf_tflops = sum(p["flops"] for p in mrf) / (mrf_ms / 1e3) / 1e12
        tensor_peak = peaks["bf16_tflops_sustained"]
        peak_note = "bf16 dense, sustained"
        if args.mode == "tf32":
            tensor_peak = tensor_peak / 2.0
            peak_note = "tf32 dense = measured bf16 sustained / 2 (no tf32 peak is measured)"
        elif args.mode == "fp32":
            peak_note = "bf16 dense sustained (this mode runs fp32 FFMA kernels, not tensor cores)"
        traffic, traffic_note = None, "no ncu capture committed for this kernel"
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            ent = tj.get(f"{args.mode}:{dom['kernel']}")
            if ent:
                traffic, traffic_note = ent["dram_bytes_per_launch"], ent["source"]
        roofline = {
            "bound": "tensor",
            "kernel": "tc_pair_kernel %s (%d launches per step)" % (dom["kernel"], dom["launches"]),
            "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
            "frac": achieved / tensor_peak, "traffic": traffic, "traffic_note": traffic_note,
            "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"],
            "peak_source": peaks["source"] + "; " + peak_note,
            "share_of_step": dom["ms"] / step_ms_prof if step_ms_prof else None,
            "serial_step_ms": step_ms_prof,
            "all_mrf_launches": {"launches": sum(p["launches"] for p in mrf), "tflops": mrf_tflops,
                                 "frac": mrf_tflops / tensor_peak, "share_of_step": mrf_ms / step_ms_prof},
            "per_stage": [{"kernel": p["kernel"], "launches": p["launches"], "ms": round(p["ms"], 4),
                           "tflops": round(p["flops"] / (p["ms"] / 1e3) / 1e12, 2) if p["ms"] > 0 else None,
                           "gbs": round(p["bytes"] / (p["ms"] / 1e3) / 1e9, 1) if p["ms"] > 0 else None}
                          for p in prof],
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": dev_ms_max / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.mode], "data": "synthetic",
            "config": {"workload": f"HiFi-GAN generator (default YAML config, random-init), batch {B} x {T} frames "
                                   f"({audio_seconds(1, T):.2f} s utterances) per GPU, mode {args.mode}",
                       "parallelism": f"utterance-sharded x{world}, no data-path collective",
                       "l2": "256 MiB flush between timed steps",
                       "streams": "timed steps: the 3 resblocks of each MRF on 3 streams (fork/join events); the "
                                  "per-kernel roofline pass serialises them so every launch is timed alone",
                       "e2e_timer": "host perf_counter around synchronous calls"},
            "ms_per_step_median": step_ms[len(step_ms) // 2], "ms_per_step_best": step_ms[0],
            "tflops_per_gpu": flops_step * steps / (dev_ms_max / 1e3) / 1e12,
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": int(mel_host.numel() * 4),
                    "d2h_bytes_per_step": int(wav_host.numel() * 4),
                    "steps": e2e_steps},
            "gpu_launches": int(launches_per_step * steps),
            "quality": quality,
            "other_modes": other,
            "clocks": clocks,
            "roofline": roofline,
        }
        if not args.no_cpu_baseline:
            sample_b = 4
            times, cores = cpu_reference_run(sample_b, T, 3)
            best = min(times)
            line["cpu_baseline"] = {
                "value": audio_seconds(sample_b, T) / best, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{sample_b} of {B} utterances x {T} frames, best of 3, oracle/torch_port.py "
                          "(the ATen conv kernels the reference dispatches to)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
