"""One warm-up forward + one forward of the bench workload, for ncu captures:

    ncu --set full -k regex:tc_conv_kernel -s <77 + idx> -c 1 ... python tools/profile_step.py --mode bf16

tc_conv_kernel launch order inside one forward (default config):
  0 conv_pre | 1 ups0 | 2..19 mrf0 | 20 ups1 | 21..38 mrf1 | 39 ups2 | 40..57 mrf2 | 58 ups3 | 59..76 mrf3
  inside an MRF: resblock j (k = 3, 7, 11) x pair l (d = 1, 3, 5) x (conv1, conv2) -> offset 6 j + 2 l + c
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="bf16")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=172)
ap.add_argument("--forwards", type=int, default=2)
a = ap.parse_args()
cfg = synth.DEFAULT_CONFIG
gen = pkg.HiFiGANGenerator(**cfg, mode=a.mode).to("cuda:0")
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()})
mel = torch.from_numpy(synth.make_mel(1, a.batch, 80, a.frames)).to("cuda:0")
with torch.no_grad():
    for _ in range(a.forwards):
        wav = gen(mel)
torch.cuda.synchronize()
print("ok", tuple(wav.shape), gen.last_launch_count)
