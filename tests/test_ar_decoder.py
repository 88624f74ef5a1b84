"""KV-cached autoregressive decoder (include/hfg_ard.h) against the unmodified
reference's O(T^2) loop (goldens: tests/golden/make_ar_decoder.py), and BASELINE config 5 at its stated batch of
64 end to end: acoustic fixture -> length regulator kernel -> KV-cached decoder -> generator."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from tts_sambert_hifigan_b200 import _capi, synth

# frame-for-frame equality within fp32 round-off: the decoder output has peak ~4.5 and every frame feeds back
# into the next 95, so summation-order differences (1e-7 relative per op) compound.  Measured: oracle vs
# reference 3e-6 (CPU), CUDA vs reference recorded in gpurun_out/parity_r2.jsonl; bound = ~10x that.
TOL_MEL = 5e-5


def _ref_sd():
    sd = {k: torch.from_numpy(v) for k, v in synth.make_ard_weights(synth.ARD_DEFAULT, 300).items()}
    import oracle.ar_decoder as oard
    sd["pos_encoding.pe"] = oard.positional_encoding(5000, 256).unsqueeze(0)
    return sd


def test_header_symbols_exported():
    header = open(os.path.join(ROOT, "include", "hfg_ard.h")).read()
    declared = set(re.findall(r"\b(hfg_ard_[a-z_]+)\s*\(", header))
    assert declared == set(_capi.ARD_SYMBOLS)
    lib = ctypes.CDLL(_capi.lib_path())
    for name in declared:
        assert hasattr(lib, name), name


def test_oracle_matches_reference_goldens():
    """The KV-cached restatement (oracle/ar_decoder.py) reproduces the live reference's frames: the decoder
    alone on 8 x 96 frames, and inside the acoustic model on the 64-utterance config-5 batch."""
    import oracle.ar_decoder as oard
    sd = _ref_sd()
    g = load_golden("ar_decoder_b8")
    hvar = torch.from_numpy(synth.normal(301, (8, 96, 256)))
    with torch.no_grad():
        mel = oard.decode(sd, hvar, 6, 8).numpy()
    err = float(np.abs(mel - g["mel_pred"]).max())
    print(f"oracle vs reference, 8 x 96 frames: max-abs {err:.3e} (peak {np.abs(g['mel_pred']).max():.2f})")
    assert err <= TOL_MEL
    fx = load_golden("config5_acoustic_b64")
    with torch.no_grad():
        mel = oard.decode(sd, torch.from_numpy(fx["hvar"]), 6, 8).numpy()
    err = float(np.abs(mel - fx["mel_pred"]).max())
    print(f"oracle vs reference acoustic model, 64 x {fx['hvar'].shape[1]} frames: max-abs {err:.3e}")
    assert err <= TOL_MEL


def test_mirror_keeps_the_reference_schema():
    import tts_sambert_hifigan_b200 as pkg
    dec = pkg.PNCAARDecoder(verbose=False)
    keys = [k for k in dec.state_dict() if k != "pos_encoding.pe"]
    assert [(k, tuple(dec.state_dict()[k].shape)) for k in keys] == synth.ard_weight_shapes(synth.ARD_DEFAULT)
    assert len(dec.state_dict()) == 115 and tuple(dec.state_dict()["pos_encoding.pe"].shape) == (1, 5000, 256)
    import oracle.ar_decoder as oard
    assert torch.equal(dec.state_dict()["pos_encoding.pe"][0], oard.positional_encoding(5000, 256))
    with pytest.raises(NotImplementedError):
        dec.train()(torch.zeros(1, 4, 256), mel_gt=torch.zeros(1, 4, 80))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec.eval()(torch.zeros(1, 4, 256))
    ref_dir = "/root/reference"
    if os.path.isdir(os.path.join(ref_dir, "models")):      # CPU box: the live class, key for key
        import contextlib, io, sys
        sys.path.insert(0, ref_dir)
        try:
            from models.ar_decoder import PNCAARDecoder as Ref
        finally:
            sys.path.remove(ref_dir)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = Ref().eval()
        assert list(ref.state_dict()) == list(dec.state_dict())
        new = pkg.PNCAARDecoder.from_reference(ref, verbose=False)
        for k, v in ref.state_dict().items():
            assert torch.equal(new.state_dict()[k], v), k


def _record(case, err, peak, mode="fp32", **kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_r2.jsonl"), "a") as f:
        f.write(json.dumps(dict(case=case, mode=mode, max_abs=err, ref_peak=peak, **kw)) + "\n")


def _cuda_decoder():
    import tts_sambert_hifigan_b200 as pkg
    dec = pkg.PNCAARDecoder(verbose=False).eval().to("cuda:0")
    sd = _ref_sd()
    dec.load_state_dict(sd)
    return dec


@pytest.mark.gpu
def test_cuda_decoder_matches_reference_frames(capsys):
    g = load_golden("ar_decoder_b8")
    dec = _cuda_decoder()
    hvar = torch.from_numpy(synth.normal(301, (8, 96, 256))).to("cuda:0")
    with torch.no_grad():
        mel = dec(hvar)
        mel2 = dec(hvar)
        short = dec(hvar, max_len=10)
    torch.cuda.synchronize()
    assert mel.shape == (8, 96, 80) and mel.dtype == torch.float32
    err = float((mel.cpu() - torch.from_numpy(g["mel_pred"])).abs().max())
    per_frame = (mel.cpu() - torch.from_numpy(g["mel_pred"])).abs().amax(dim=(0, 2))
    print(f"CUDA KV-cached decoder vs reference, 8 x 96 frames: max-abs {err:.3e}; frame 0 {float(per_frame[0]):.2e}, "
          f"frame 95 {float(per_frame[95]):.2e}; launches {dec.last_launch_count}")
    _record("ar_decoder_b8", err, float(np.abs(g["mel_pred"]).max()), launches=dec.last_launch_count)
    assert err <= TOL_MEL
    assert torch.equal(mel, mel2)                          # deterministic
    assert torch.equal(short, mel[:, :10])                 # a prefix is a prefix (causal)
    assert dec.last_launch_count > 10 * 70
    # the reference's progress prints, line for line (models/ar_decoder.py:187-236)
    capsys.readouterr()
    dec.verbose = True
    with torch.no_grad():
        dec(hvar[:2, :3].contiguous())
    out = capsys.readouterr().out.splitlines()
    assert out == ["[PNCAARDecoder] Inference mode - Input Hvar shape: torch.Size([2, 3, 256])",
                   "[PNCAARDecoder] Generating 3 frames autoregressively with chunk_size=1",
                   "[PNCAARDecoder] Initial mel_pred shape: torch.Size([2, 1, 80])",
                   "[PNCAARDecoder] Chunk 0: Generated 1 frames, current shape: torch.Size([2, 2, 80])",
                   "[PNCAARDecoder] Chunk 1: Generated 1 frames, current shape: torch.Size([2, 3, 80])",
                   "[PNCAARDecoder] Chunk 2: Generated 1 frames, current shape: torch.Size([2, 4, 80])",
                   "[PNCAARDecoder] Final output mel_pred shape: torch.Size([2, 3, 80])",
                   "[PNCAARDecoder] Total chunks generated: 3"]


@pytest.mark.gpu
def test_config5_batch64_end_to_end():
    """BASELINE.json configs[4] at its stated size: 64 synthetic phoneme sequences through the unmodified
    reference acoustic model (fixture) -> on the GPU: length regulator kernel (bit-exact frame indexing), KV-cached
    decoder (frames within round-off of the reference's), generator reading the decoder's [B, T, 80] layout."""
    import time
    import oracle
    import tts_sambert_hifigan_b200 as pkg
    fx = load_golden("config5_acoustic_b64")
    hvar, mel_ref, dur = fx["hvar"], fx["mel_pred"], fx["dur"]
    B, T = hvar.shape[:2]
    assert B == 64 and mel_ref.shape == (64, T, 80) and dur.min() >= 1 and T == int(dur.sum(axis=1).max())
    # (1) integer frame indexing: repeat_interleave + zero padding, bit for bit
    lr = pkg.LengthRegulator()
    hlr = lr(torch.from_numpy(fx["henc"]).cuda(), torch.from_numpy(dur).cuda())
    assert np.array_equal(hlr.cpu().numpy(), fx["hlr"])
    # (2) KV-cached decoder on the reference's Hvar
    dec = _cuda_decoder()
    x = torch.from_numpy(hvar).to("cuda:0")
    with torch.no_grad():
        dec(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mel = dec(x)
        torch.cuda.synchronize()
        t_dec = time.perf_counter() - t0
    err = float((mel.cpu() - torch.from_numpy(mel_ref)).abs().max())
    print(f"config5 B=64: decoder {T} frames in {1e3 * t_dec:.1f} ms ({dec.last_launch_count} launches), "
          f"mel_pred max-abs vs reference {err:.3e} (peak {np.abs(mel_ref).max():.2f})")
    _record("config5_b64_mel_pred", err, float(np.abs(mel_ref).max()), decode_ms=1e3 * t_dec)
    assert err <= TOL_MEL
    # (3) generator on the decoder's own output, frames-last; against the oracle generator on the REFERENCE mel
    cfg = synth.DEFAULT_CONFIG
    sd = synth.make_weights(cfg, 0)
    ref_wav = oracle.forward_torch(cfg, {k: torch.from_numpy(v) for k, v in sd.items()},
                                   torch.from_numpy(mel_ref).transpose(1, 2).contiguous()).numpy()
    for mode, tol in (("tf32", 3e-4), ("fp16", 2.5e-4), ("bf16", 1.8e-3)):
        gen = pkg.HiFiGANGenerator(**cfg, mode=mode).to("cuda:0")
        gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        with torch.no_grad():
            wav = gen.forward_frames_last(mel)
            lens = torch.from_numpy(dur.sum(axis=1).astype(np.int32))
            rag = gen.forward_ragged(mel.transpose(1, 2).contiguous(), lens)
        torch.cuda.synchronize()
        e = float(np.abs(wav.cpu().numpy() - ref_wav).max())
        print(f"config5 B=64 wav[{mode}] max-abs vs oracle on the reference mel {e:.3e} (peak {np.abs(ref_wav).max():.3f})")
        _record("config5_b64_wav", e, float(np.abs(ref_wav).max()), mode=mode)
        assert wav.shape == (64, 1, T * 256) and e <= tol
        for i in range(B):                                 # length-aware run: valid region identical
            n = int(lens[i]) * 256
            assert torch.equal(rag[i, :, :n], wav[i, :, :n])


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [
    dict(d_model=64, n_mels=20, n_layers=2, n_heads=4, d_ff=96, B=70, T=9, max_len=12),     # head_dim 16, two row tiles, max_len > frames
    dict(d_model=128, n_mels=80, n_layers=1, n_heads=2, d_ff=520, B=3, T=33, max_len=5),    # head_dim 64, split-K with a ragged last chunk, max_len < frames
    dict(d_model=128, n_mels=16, n_layers=1, n_heads=1, d_ff=64, B=2, T=4, max_len=4),      # head_dim 128
])
def test_cuda_decoder_other_geometries_match_oracle(geom):
    """Geometries the reference never instantiates (its defaults are fixed in models/acoustic_model.py:112-114) against
    the KV-cached oracle, which is itself pinned to the reference at the default geometry."""
    import oracle.ar_decoder as oard
    import tts_sambert_hifigan_b200 as pkg
    cfg = {k: geom[k] for k in ("d_model", "n_mels", "n_layers", "n_heads", "d_ff")}
    sd = {k: torch.from_numpy(v) for k, v in synth.make_ard_weights(cfg, 410).items()}
    sd["pos_encoding.pe"] = oard.positional_encoding(5000, cfg["d_model"]).unsqueeze(0)
    hvar = torch.from_numpy(synth.normal(411, (geom["B"], geom["T"], cfg["d_model"])))
    with torch.no_grad():
        want = oard.decode(sd, hvar, cfg["n_layers"], cfg["n_heads"], max_len=geom["max_len"])
    dec = pkg.PNCAARDecoder(**cfg, verbose=False).eval().to("cuda:0")
    dec.load_state_dict(sd)
    with torch.no_grad():
        got = dec(hvar.to("cuda:0"), max_len=geom["max_len"])
    torch.cuda.synchronize()
    assert got.shape == want.shape == (geom["B"], geom["max_len"], cfg["n_mels"])
    err = float((got.cpu() - want).abs().max())
    print(f"decoder {cfg}: max-abs vs oracle {err:.3e} (peak {float(want.abs().max()):.2f})")
    assert err <= TOL_MEL


@pytest.mark.gpu
def test_cuda_decoder_rejects_bad_arguments():
    import tts_sambert_hifigan_b200 as pkg
    from tts_sambert_hifigan_b200 import ar_decoder
    with pytest.raises(_capi.HfgError) as e:               # head_dim 24 is not a supported tile
        ar_decoder._Handle(96, 80, 1, 4, 64, 5000)
    assert e.value.code == _capi.ERR_UNSUPPORTED
    dec = pkg.PNCAARDecoder(d_model=64, n_mels=8, n_layers=1, n_heads=2, d_ff=32, verbose=False).eval().to("cuda:0")
    with pytest.raises(RuntimeError):
        dec(torch.zeros(1, 4, 65, device="cuda:0"))
    with pytest.raises(_capi.HfgError) as e:               # beyond the positional-encoding table (reference: 5000 rows)
        dec(torch.zeros(1, 4, 64, device="cuda:0"), max_len=5001)
    assert e.value.code == _capi.ERR_INVALID
