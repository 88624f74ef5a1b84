"""One warm-up forward + one forward of the bench workload, for ncu captures.

Launch order inside one forward (default config; ncu serialises the three streams in enqueue order):
  zero_pads, pack_mel, conv_pre, then per stage i: ups{i}, and the MRF as
  k3 pair0, k3 pair1, k7 pair0, k7 pair1, k11 pair0, k11 pair1, k3 pair2, k7 pair2, k11 pair2
  (pair l = dilation 1, 3, 5; in tf32 mode stage 0 is unfused: two tc_conv_kernel launches per pair),
  finally conv_post.  53 launches in tf32, 44 in bf16.

    # mrf1.k11 pair 0 (dominant kernel), second forward: 9 fused launches per stage
    ncu --set full -k regex:tc_pair_kernel -s 31 -c 1 ... python tools/profile_step.py --mode tf32   # 27 + 4
    ncu --set full -k regex:tc_pair_kernel -s 49 -c 1 ... python tools/profile_step.py --mode bf16   # 36 + 9 + 4
    tools/ncu_forward.sh tf32 53        # every kernel of the second forward, light metric set
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
