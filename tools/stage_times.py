"""Stage-level device times of the bench workload with the resblocks running concurrently
(hfg_set_profiling(2)) next to the serial per-launch sums (hfg_set_profiling(1))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import synth

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = synth.DEFAULT_CONFIG
gen = pkg.HiFiGANGenerator(**cfg, mode=mode).to("cuda:0")
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()})
mel = torch.from_numpy(synth.make_mel(1, batch, 80, 172)).to("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
h = gen._handle_for(torch.device("cuda", 0))
with torch.no_grad():
    for _ in range(3):
        gen(mel)
    out = {}
    for level in (2, 1):
        h.set_profiling(level)
        acc = {}
        for _ in range(5):
            flush.fill_(1)
            gen(mel); torch.cuda.synchronize()
            for p in h.get_profile():
                key = p["kernel"] if level == 2 else p["kernel"].split(".")[0]
                acc[key] = acc.get(key, 0.0) + p["ms"] / 5
        h.set_profiling(0)
        out[level] = acc
keys = list(out[2].keys())
ser = out[1]
ser_head = sum(v for k, v in ser.items() if k in ("zero_pads", "pack_mel", "conv_pre"))
print(f"# {mode} batch {batch}: stage  concurrent_ms  serial_sum_ms")
for k in keys:
    s = ser_head if k == "head" else ser.get("conv_post" if k == "tail" else k, float("nan"))
    print(f"{k:6s} {out[2][k]:8.3f} {s:8.3f}")
print(f"total  {sum(out[2].values()):8.3f} {sum(ser.values()):8.3f}")
