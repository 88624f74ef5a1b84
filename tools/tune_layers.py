"""Per-layer micro-benchmark of the tensor-core conv kernel (hfg_bench_layer).

    python tools/tune_layers.py [--mode bf16] [--batch 16] [--frames 172]

Prints ms and TFLOP/s for conv1 of every (stage, resblock, pair) at the bench
workload's stage lengths, so fixed per-launch cost and per-tap cost separate."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import _capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="bf16")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=172)
ap.add_argument("--stages", default="0,1,2,3")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--which", default="0")
ap.add_argument("--resblocks", default="0,1,2")
ap.add_argument("--pairs", default="0,1,2")
a = ap.parse_args()
cfg = synth.DEFAULT_CONFIG
gen = pkg.HiFiGANGenerator(**cfg, mode=a.mode).to("cuda:0")
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()})
h = gen._handle_for(torch.device("cuda:0"))
mode = _capi.MODES[a.mode]
rows = a.frames
C = cfg["upsample_initial_channel"]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("HFG_TC"))
print(f"# mode={a.mode} batch={a.batch} frames={a.frames} {tag}")
for i, u in enumerate(cfg["upsample_rates"]):
    rows *= u
    C //= 2
    if str(i) not in a.stages.split(","):
        continue
    for j, k in enumerate(cfg["resblock_kernel_sizes"]):
        if str(j) not in a.resblocks.split(","):
            continue
        for l, d in enumerate(cfg["resblock_dilation_sizes"][j]):
            if str(l) not in a.pairs.split(","):
                continue
            for which in [int(w) for w in a.which.split(",")]:
                ms = h.bench_layer(i, j, l, which, a.batch, rows, mode, a.iters)
                fl = 2.0 * C * C * k * a.batch * rows * (2 if which == 2 else 1)
                print(f"stage{i} C={C:3d} rows={a.batch * rows:7d} k={k:2d} d={d} {['conv1', 'conv2', 'pair '][which]}: "
                      f"{ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")
