"""Print a SHA-256 of the generator output for a fixed small input (tests compare kernel variants selected
through HFG_TC_* environment knobs, which are read once per process)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import synth

mode = sys.argv[1]
cfg = synth.DEFAULT_CONFIG
gen = pkg.HiFiGANGenerator(**cfg, mode=mode).to("cuda:0")
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 7).items()})
mel = torch.from_numpy(synth.make_mel(11, 3, 80, 97)).to("cuda:0")
with torch.no_grad():
    wav = gen(mel)
torch.cuda.synchronize()
print("HASH", hashlib.sha256(wav.cpu().numpy().tobytes()).hexdigest(), gen.last_launch_count)
if len(sys.argv) > 2 and sys.argv[2] == "--oracle":        # max-abs error against the oracle on the same input (tests only)
    import oracle
    ref = oracle.forward_torch(cfg, {k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 7).items()}, mel.cpu()).numpy()
    print("MAXABS", float(abs(wav.cpu().numpy() - ref).max()), float(abs(ref).max()))
