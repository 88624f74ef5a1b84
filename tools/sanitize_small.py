"""Tiny forward in every mode for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import synth
cfg = synth.DEFAULT_CONFIG
sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()}
mel = torch.from_numpy(synth.make_mel(1, 2, 80, 9)).cuda()
for mode in sys.argv[1:] or ["fp32", "tf32", "bf16"]:
    gen = pkg.HiFiGANGenerator(**cfg, mode=mode).cuda()
    gen.load_state_dict(sd)
    with torch.no_grad():
        w = gen(mel)
    torch.cuda.synchronize()
    print(mode, tuple(w.shape), float(w.abs().max()))
