"""Where the end-to-end (host buffer) time goes: device-resident call, public host call, raw C-ABI host call."""
import os, sys, time, statistics as st
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import synth, _capi

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg = synth.DEFAULT_CONFIG
gen = pkg.HiFiGANGenerator(**cfg, mode=mode).to("cuda:0")
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()})
mel_h = torch.from_numpy(synth.make_mel(1, 16, 80, 172)).pin_memory()
mel_d = mel_h.cuda()

def timeit(fn, n):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return f"median {st.median(ts):.3f} mean {st.mean(ts):.3f} p90 {sorted(ts)[int(n*0.9)]:.3f} max {max(ts):.3f} ms"

def dev():
    gen(mel_d); torch.cuda.synchronize()
print(mode, "device-resident + sync :", timeit(dev, n))
print(mode, "public host call       :", timeit(lambda: gen(mel_h), n))
h = gen._handle_for(torch.device("cuda", 0))
wav = torch.empty((16, 1, 44032), dtype=torch.float32, pin_memory=True)
print(mode, "raw C-ABI, pinned bufs :", timeit(lambda: h.forward_host(mel_h.data_ptr(), 16, 172, wav.data_ptr(), _capi.MODES[mode], mel_pinned=True, wav_pinned=True), n))
mel_p = mel_h.clone()  # pageable
wav_p = torch.empty((16, 1, 44032), dtype=torch.float32)
print(mode, "raw C-ABI, pageable    :", timeit(lambda: h.forward_host(mel_p.data_ptr(), 16, 172, wav_p.data_ptr(), _capi.MODES[mode], mel_pinned=False, wav_pinned=False), n))
print(mode, "pinned alloc 2.8MB     :", timeit(lambda: torch.empty((16, 1, 44032), dtype=torch.float32, pin_memory=True), n))
